// Inverted multi-index (IMI) pieces for the IMI-PQ baseline index of BASELINE configs[4]:
//   coarse quantizer = MultiIndexQuantizer(d, 2, nbits): two codebooks of K = 2^nbits centroids over the two halves of
//   the vector, cell (i1, i2) has label i1 | i2 << nbits (IndexPQ.cpp:780-857); the index is an IVFPQ over the K^2 cells
//   with residual PQ codes (tests/sift1b_imi_pq.cpp:216-236, IndexIVFPQ.cpp:645-687 use_precomputed_table = 2).
//
//   vlq_imi_top_cells   the nprobe cells with the smallest d1[i1] + d2[i2] per query -- the reference walks the two
//                       sorted tables with a heap (MinSumK, IndexPQ.cpp:637-778).  Here: a cell of rank r can only
//                       combine the a-th smallest of d1 with the b-th smallest of d2 where (a+1)(b+1) <= nprobe, so the
//                       ~nprobe ln(nprobe) sums under that hyperbola are formed from the two sorted top-nprobe prefixes
//                       (vlq_select_rows) and one block select keeps the nprobe smallest, ties to the lowest label.
//   vlq_imi_encode      residual to the cell centroid, PQ code by direct differences (first minimum wins,
//                       ProductQuantizer.cpp:311-336) and the per-entry scalar kappa = ||p||^2 + 2 c.p that lets the
//                       VLQ scan kernels serve as the IVFPQ scan: dist = ||q - c||^2 + kappa - 2 q.p = ||q - c - p||^2
//                       (the reference splits 2 c.p into two per-half tables, IndexIVFPQ.cpp:645-687).
//   vlq_copy_columns    contiguous copy of a column block (the two halves of the query / database rows).
// The scan itself is vlq_scan_topk with a one-level lambda codebook {0}, term1 = the cell distance, edge_d2 = NULL.
#include <cfloat>

#include "topk.cuh"

namespace vlq {

__global__ void copy_columns_kernel(const float* __restrict__ src, int64_t n, int64_t ld, int col0, int ncols,
                                    float* __restrict__ dst) {
  const int64_t total = n * ncols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ncols;
    const int c = (int)(i - r * ncols);
    dst[i] = src[r * ld + col0 + c];
  }
}

constexpr int IMI_THREADS = 256;

// v1/i1, v2/i2: [nq][L] ascending prefixes of the two half-distance tables (values, centroid ids)
__global__ void __launch_bounds__(IMI_THREADS)
imi_top_cells_kernel(const float* __restrict__ v1, const int* __restrict__ i1, const float* __restrict__ v2,
                     const int* __restrict__ i2, int L, int nbits, int nprobe, int cap, int* __restrict__ out_cell,
                     float* __restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char smem[];
  BlockSelect<IMI_THREADS> sel;
  sel.init(smem, nprobe, cap, 1);
  const int64_t q = blockIdx.x;
  const float* a = v1 + q * L;
  const float* b = v2 + q * L;
  const int* ia = i1 + q * L;
  const int* ib = i2 + q * L;
  // rows a = 0 .. L-1 of the hyperbola; row a holds b = 0 .. min(L, nprobe / (a + 1)) - 1.  Threads walk the flattened
  // (a, b) enumeration in batches of IMI_THREADS; the row of a position is found by walking (rows shrink monotonically).
  int row = 0, row_start = 0;
  int row_len = min(L, nprobe);
  int64_t total = 0;
  for (int r = 0; r < L; r++) {
    const int len = min(L, nprobe / (r + 1));
    if (len == 0) break;
    total += len;
  }
  for (int64_t base = 0; base < total; base += IMI_THREADS) {
    const int64_t pos = base + threadIdx.x;
    bool valid = pos < total;
    float s = 0.f;
    uint32_t label = 0;
    if (valid) {
      while (pos >= (int64_t)row_start + row_len) {  // advance to the row of this position
        row_start += row_len;
        row++;
        row_len = min(L, nprobe / (row + 1));
      }
      const int col = (int)(pos - row_start);
      const int ca = ia[row], cb = ib[col];
      valid = ca >= 0 && cb >= 0;
      if (valid) {
        s = a[row] + b[col];
        label = (uint32_t)ca | ((uint32_t)cb << nbits);
      }
    }
    const bool any = sel.offer_f(valid, s, label);
    sel.end_batch(any);
  }
  sel.finish();
  for (int i = threadIdx.x; i < nprobe; i += IMI_THREADS) {
    const uint64_t key = sel.keys[i];
    out_cell[q * nprobe + i] = key != kKeyInf ? (int)key_payload(key) : -1;
    out_dist[q * nprobe + i] = key != kKeyInf ? key_val(key) : FLT_MAX;
  }
}

// one warp per vector; lane l owns the codewords l, l + 32, ... of every sub-quantizer
__global__ void __launch_bounds__(256)
imi_encode_kernel(const float* __restrict__ x, int64_t n, int d, const int* __restrict__ a1, const int* __restrict__ a2,
                  const float* __restrict__ cb1, const float* __restrict__ cb2, int nbits, const float* __restrict__ pq,
                  int M, int ksub, int dsub, int* __restrict__ out_cell, uint8_t* __restrict__ out_codes,
                  float* __restrict__ out_kappa) {
  extern __shared__ float sm[];  // per warp: residual r[d], centroid c[d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* r = sm + (size_t)warp * 2 * d;
  float* c = r + d;
  const int h = d / 2;
  for (int64_t v = blockIdx.x * 8 + warp; v < n; v += (int64_t)gridDim.x * 8) {
    const int i1 = a1[v], i2 = a2[v];
    const bool ok = i1 >= 0 && i2 >= 0;  // invalid (NaN) rows get no cell, like the VLQ encoder
    if (lane == 0) out_cell[v] = ok ? (i1 | (i2 << nbits)) : -1;
    if (!ok) {
      for (int m = lane; m < M; m += 32) out_codes[v * M + m] = 0;
      if (lane == 0) out_kappa[v] = 0.f;
      continue;
    }
    for (int t = lane; t < d; t += 32) {
      const float cv = t < h ? cb1[(size_t)i1 * h + t] : cb2[(size_t)i2 * h + (t - h)];
      c[t] = cv;
      r[t] = x[v * d + t] - cv;
    }
    __syncwarp();
    float kappa = 0.f;
    for (int m = 0; m < M; m++) {
      const float* rm = r + m * dsub;
      const float* cm = c + m * dsub;
      float best = FLT_MAX, bk = 0.f;
      int bj = 0x7fffffff;
      for (int j = lane; j < ksub; j += 32) {
        const float* p = pq + ((size_t)m * ksub + j) * dsub;
        float dis = 0.f, pn = 0.f, cp = 0.f;
        for (int t = 0; t < dsub; t++) {
          const float pv = p[t];
          const float df = rm[t] - pv;
          dis = fmaf(df, df, dis);
          pn = fmaf(pv, pv, pn);
          cp = fmaf(cm[t], pv, cp);
        }
        if (dis < best) {  // j ascending within the lane: first minimum wins
          best = dis;
          bj = j;
          bk = pn + 2.f * cp;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {  // arg-min over the lanes, ties to the lowest codeword
        const float ob = __shfl_xor_sync(kFull, best, o);
        const int oj = __shfl_xor_sync(kFull, bj, o);
        const float ok_ = __shfl_xor_sync(kFull, bk, o);
        if (ob < best || (ob == best && oj < bj)) {
          best = ob;
          bj = oj;
          bk = ok_;
        }
      }
      if (lane == 0) out_codes[v * M + m] = (uint8_t)bj;
      kappa += bk;
    }
    if (lane == 0) out_kappa[v] = kappa;
    __syncwarp();
  }
}

}  // namespace vlq

extern "C" {

int vlq_copy_columns(const float* src, int64_t n, int64_t ld, int col0, int ncols, float* dst, vlq_stream_t stream) {
  using namespace vlq;
  if (n < 0 || ncols <= 0 || col0 < 0 || ld < col0 + ncols) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!src || !dst) return VLQ_EINVAL;
  VLQ_LAUNCH(copy_columns_kernel, 148 * 8, 256, 0, as_stream(stream), src, n, ld, col0, ncols, dst);
  return last_error();
}

int vlq_imi_top_cells(const float* v1, const int* i1, const float* v2, const int* i2, int64_t nq, int L, int nbits,
                      int nprobe, int* out_cell, float* out_dist, vlq_stream_t stream) {
  using namespace vlq;
  if (nq < 0 || L <= 0 || nbits <= 0 || nbits > 15 || nprobe <= 0 || nprobe > VLQ_MAX_K) return VLQ_EINVAL;
  if (nq == 0) return VLQ_OK;
  if (!v1 || !i1 || !v2 || !i2 || !out_cell || !out_dist) return VLQ_EINVAL;
  const int cap = select_capacity(nprobe, IMI_THREADS, 1, (long long)nprobe * 16);
  const size_t smem = select_smem_bytes(cap);
  VLQ_CUDA_TRY(cudaFuncSetAttribute(imi_top_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VLQ_LAUNCH(imi_top_cells_kernel, (unsigned)nq, IMI_THREADS, smem, as_stream(stream), v1, i1, v2, i2, L, nbits, nprobe,
             cap, out_cell, out_dist);
  return last_error();
}

int vlq_imi_encode(const float* x, int64_t n, int d, const int* a1, const int* a2, const float* cb1, const float* cb2,
                   int nbits, const float* pq, int M, int* out_cell, uint8_t* out_codes, float* out_kappa,
                   vlq_stream_t stream) {
  using namespace vlq;
  if (n < 0 || d <= 0 || d % 2 != 0 || M <= 0 || d % M != 0 || (d / 2) % (d / M) != 0 || nbits <= 0 || nbits > 15)
    return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!x || !a1 || !a2 || !cb1 || !cb2 || !pq || !out_cell || !out_codes || !out_kappa) return VLQ_EINVAL;
  const size_t smem = sizeof(float) * 8 * 2 * d;
  const unsigned grid = (unsigned)(div_up(n, 8) < 148 * 8 ? div_up(n, 8) : 148 * 8);
  VLQ_LAUNCH(imi_encode_kernel, grid, 256, smem, as_stream(stream), x, n, d, a1, a2, cb1, cb2, nbits, pq, M, 256, d / M,
             out_cell, out_codes, out_kappa);
  return last_error();
}

}  // extern "C"
