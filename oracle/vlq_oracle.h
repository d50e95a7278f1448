/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's VLQ hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product path (vector_line_quantization_b200/) never includes, links or calls anything in oracle/.
 *
 * PARITY PINNING: the reference holds no golden vectors and no CPU implementation of VLQ.
 *  - The stock pieces restated here (coarse assignment, k-means, PQ training/encoding, heap top-k, shard merge and
 *    the lambda==0 slice of the ADC scan == IndexIVFPQ precomputed-table search) ARE pinned against the reference
 *    itself, compiled unmodified into oracle/_ref/libfaiss_ref.so (tests/test_oracle_vs_reference.py) and against
 *    fixtures generated from it (tests/golden/make_golden.py).
 *  - The VLQ-only arithmetic (graph, line stage, lambda quantiser, residual, line selection, lambda-dependent scan)
 *    exists in the reference only as Pascal-era CUDA that cannot be compiled for sm_100 or run here:
 *    for those functions parity is UNPINNED by the reference; it is pinned only by this restatement, whose outputs
 *    are frozen as fixtures, and by mathematical identities checked in tests/ (distance == ||q - recon||^2 - ||q||^2).
 *
 * All arithmetic fp32 (compiled with -ffp-contract=off) unless a function says otherwise; "argmin" means lowest
 * index on an exact tie. All matrices row-major, compact.
 */
#ifndef VLQ_ORACLE_H
#define VLQ_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int vlqo_num_threads(void);

/* glibc random_r with the reference's 8-byte state (utils.cpp:135-160) and its Fisher-Yates (utils.cpp:307-317) */
void vlqo_rand_perm(int* perm, long n, long seed);

/* Coarse distances in GEMM form, D = ||c||^2 - 2 x.c (+ ||x||^2 when add_xnorm), and the k smallest per row,
 * ascending, ties by lowest index.  gpu/impl/Distance.cu:579-710 + L2Select.cu:25-165 (+BroadcastSum.cu:677-698);
 * CPU twin utils.cpp:833-901.  outD/outI are [n][k]. */
void vlqo_l2_topk(const float* x, long n, int d, const float* cent, long C, int k, int add_xnorm, float* outD,
                  int* outI);

/* Clustering::train (Clustering.cpp:66-206) + km_update_centroids (utils.cpp:1369-1449) with the flat assigner
 * above (add_xnorm=1 as IndexFlatL2 does).  x is [n][d]; centroids_out [k][d]; obj_out (may be NULL) [niter]. */
void vlqo_kmeans(int d, int k, long n, const float* x, int niter, long seed, int max_points_per_centroid,
                 float* centroids_out, float* obj_out);

/* kNN graph of the centroids: E+1 nearest by exact-flag GEMM-form distance, rank 0 dropped blindly
 * (gpu/GpuIndexFlat.cu:869-893).  edge [C][E], edge_d2 [C][E]. */
void vlqo_knn_graph(const float* cent, long C, int d, int E, int* edge, float* edge_d2);

/* Line stage (gpu/GpuIndexFlat.cu:466-550, gpu/utils/triangle.cuh:54-87), intended semantics of SURVEY Q1:
 * argmin_e q_e over lines with 0<=lambda_e<=1, else global argmin; ties -> lowest e.
 * A[n] nearest centroid; out_list = A*E+e, out_lambda float (unclamped). */
void vlqo_line_stage(const float* x, long n, int d, const int* A, const float* cent, const int* edge,
                     const float* edge_d2, int E, int* out_list, float* out_lambda);

/* lambda -> uint8 = argmin_j (lambda - cb[j])^2, lowest j on ties (gpu/GpuIndexFlat.cu:579-596) */
void vlqo_lambda_quantize(const float* lambda, long n, const float* cb, int nL, uint8_t* out);

/* r = x - ((1-l)c_A + l c_s), l = cb[lamq]  (gpu/GpuIndexFlat.cu:1111-1122) */
void vlqo_residual(const float* x, long n, int d, const int* list, const uint8_t* lamq, const float* cb,
                   const float* cent, const int* edge, int E, float* r);

/* PQ encode: per sub-space argmin_j ||r_m - p_mj||^2 by direct differences, first minimum wins
 * (ProductQuantizer.cpp:311-336).  pq is (M, ksub, dsub); codes [n][M]. */
void vlqo_pq_encode(const float* r, long n, int d, const float* pq, int M, int ksub, uint8_t* codes);

/* Stable counting sort of entries into lists (what the reference's per-list push_back produces,
 * gpu/GpuIndexIVFPQ.cu:741-858): offsets [nlists+1], perm [n] = entry ordinals in list-major order. */
void vlqo_build_lists(const int* list, long n, long nlists, long* offsets, long* perm);

/* term2 table T2[c][m][j] = ||p_mj||^2 + 2 c_m . p_mj  (gpu/impl/IVFPQ.cu:599-684, BroadcastSum.cu:872-889) */
void vlqo_term2(const float* cent, long C, int d, const float* pq, int M, int ksub, float* T2);

/* Full search (SURVEY.md section 9 "Search"): coarse top-P without ||q||^2, line scoring + top-W, term3, the
 * lambda-dependent ADC over the first min(len,cap) entries of each selected list, top-k ascending padded with
 * (FLT_MAX,-1).  codes/lamq/ids are in list-major order described by offsets.  T2 may be NULL (computed inside).
 * Optional outputs (may be NULL): out_coarse [nq][P] ids, out_lines [nq][W] list ids (-1 padded),
 * out_nscanned [nq] number of entries scanned. */
void vlqo_search(const float* q, long nq, int d, const float* cent, long C, const int* edge, const float* edge_d2,
                 int E, const float* lambda_cb, int nL, const float* pq, int M, int ksub, const float* T2,
                 const long* offsets, const uint8_t* codes, const uint8_t* lamq, const long* ids, int P, int W, int k,
                 int cap, float* outD, long* outI, int* out_coarse, int* out_lines, long* out_nscanned);

/* The scan half of vlqo_search on a GIVEN line choice: lines [nq][W] list ids in rank order, -1 padded at the end
 * (term1 / term6 / term5 are recomputed from the codebooks).  Test infrastructure for implementations whose line
 * selection differs from the oracle's by a verified line-score near-tie. */
void vlqo_scan_lines(const float* q, long nq, int d, const float* cent, long C, const int* edge, const float* edge_d2,
                     int E, const float* lambda_cb, int nL, const float* pq, int M, int ksub, const float* T2,
                     const long* offsets, const uint8_t* codes, const uint8_t* lamq, const long* ids,
                     const int* lines, int W, int k, int cap, float* outD, long* outI);

/* Shard merge: k smallest of the R*k candidates per query, [rank][nq][k] in, ties by (rank,pos)
 * (gpu/GpuIndexIVFPQ.cu:1467-1518; CPU twin MetaIndexes.cpp:290-347). */
void vlqo_merge_topk(const float* D, const long* I, int R, long nq, int k, float* outD, long* outI);

/* float64 decode check value: ||q - ((1-l)c + l s) - p(code)||^2 - ||q||^2 for one entry (identity used in tests) */
double vlqo_decode_distance(const float* q, int d, const float* c, const float* s, float l, const float* pq, int M,
                            int ksub, const uint8_t* code);

/* ---- the inverted multi-index coarse quantizer of BASELINE configs[4].
 * MultiIndexQuantizer::search (IndexPQ.cpp:813-855): per sub-space tables ||x_m - c_mj||^2
 * (ProductQuantizer.cpp:410-462), then the k cells with the smallest table sums, ascending; cell label =
 * sum_m j_m * ksub^m.  k == 1: per-sub-space arg-min (first minimum).  k > 1: best-first walk of the M-dimensional grid
 * of the sorted tables (the multi-sequence algorithm; reference MinSumK, IndexPQ.cpp:636-778).  cent = (M, ksub, dsub).
 * Order among cells with EQUAL sums is unspecified in the reference (heap order); here: by position in the sorted grid. */
void vlqo_imi_search(const float* x, long n, int d, const float* cent, int M, int ksub, int k, float* D, long* I);

#ifdef __cplusplus
}
#endif
#endif
