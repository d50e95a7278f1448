/* TEST INFRASTRUCTURE ONLY -- see vlq_oracle.h for scope, provenance and the parity-pinning statement.
 * CPU restatement of the reference's VLQ hot path; every function cites the reference file:line it follows.
 */
#define _GNU_SOURCE
#include "vlq_oracle.h"

#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int vlqo_num_threads(void) { return omp_get_max_threads(); }

/* ---------------------------------------------------------------------------------------------------------------
 * fp32 reductions.  Eight strided partial sums combined in a fixed tree: deterministic, vectorisable, and no less
 * legitimate than the (unspecified) summation order of the reference's BLAS / warp reductions.
 * ------------------------------------------------------------------------------------------------------------- */
static inline float dot8(const float* a, const float* b, int d) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int j = 0;
  for (; j + 8 <= d; j += 8) {
#pragma omp simd
    for (int l = 0; l < 8; l++) acc[l] += a[j + l] * b[j + l];
  }
  for (int l = 0; j < d; j++, l++) acc[l] += a[j] * b[j];
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

static inline float l2sqr8(const float* a, const float* b, int d) {
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int j = 0;
  for (; j + 8 <= d; j += 8) {
#pragma omp simd
    for (int l = 0; l < 8; l++) {
      float t = a[j + l] - b[j + l];
      acc[l] += t * t;
    }
  }
  for (int l = 0; j < d; j++, l++) {
    float t = a[j] - b[j];
    acc[l] += t * t;
  }
  return ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
}

/* ---------------------------------------------------------------------------------------------------------------
 * Exact top-k of a stream, ascending by (value, index).  Replaces Heap.h:89-143,296-323 / Select.cuh:77-277:
 * both yield the k smallest values ascending; tie order is unspecified there, fixed here to lowest index first.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  float v;
  long i;
} cand_t;

static inline int cand_less(float av, long ai, float bv, long bi) { return av < bv || (av == bv && ai < bi); }

/* sorted-insertion buffer of capacity k (k <= 1024 everywhere on this path) */
static inline void topk_push(cand_t* buf, int* cnt, int k, float v, long i) {
  int n = *cnt;
  if (n == k) {
    if (!cand_less(v, i, buf[n - 1].v, buf[n - 1].i)) return;
    n--;
  }
  int p = n;
  while (p > 0 && cand_less(v, i, buf[p - 1].v, buf[p - 1].i)) {
    buf[p] = buf[p - 1];
    p--;
  }
  buf[p].v = v;
  buf[p].i = i;
  *cnt = n + 1;
}

/* ---------------------------------------------------------------------------------------------------------------
 * RNG: glibc random_r on an 8-byte state, exactly as utils.cpp:135-160 constructs it.
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct {
  char state[8];
  struct random_data data;
} rng_t;

static void rng_init(rng_t* r, long seed) {
  memset(&r->data, 0, sizeof(r->data));
  initstate_r((unsigned)seed, r->state, sizeof(r->state), &r->data);
}
static int rng_int(rng_t* r) {
  int32_t a;
  random_r(&r->data, &a);
  return a;
}
static float rng_float(rng_t* r) { return rng_int(r) / (float)(1L << 31); } /* utils.cpp:207-210 */

void vlqo_rand_perm(int* perm, long n, long seed) { /* utils.cpp:307-317 */
  for (long i = 0; i < n; i++) perm[i] = (int)i;
  rng_t rng;
  rng_init(&rng, seed);
  for (long i = 0; i + 1 < n; i++) {
    int i2 = (int)(i + rng_int(&rng) % (int)(n - i));
    int t = perm[i];
    perm[i] = perm[i2];
    perm[i2] = t;
  }
}

/* ---------------------------------------------------------------------------------------------------------------
 * Coarse assignment / top-k (a2, a11).
 * ------------------------------------------------------------------------------------------------------------- */
void vlqo_l2_topk(const float* x, long n, int d, const float* cent, long C, int k, int add_xnorm, float* outD,
                  int* outI) {
  float* cn = (float*)malloc(sizeof(float) * C);
#pragma omp parallel for
  for (long j = 0; j < C; j++) cn[j] = dot8(cent + j * d, cent + j * d, d);
#pragma omp parallel
  {
    cand_t* buf = (cand_t*)malloc(sizeof(cand_t) * (k + 1));
#pragma omp for schedule(dynamic, 16)
    for (long i = 0; i < n; i++) {
      const float* xi = x + i * d;
      float xn = add_xnorm ? dot8(xi, xi, d) : 0.f;
      int cnt = 0;
      for (long j = 0; j < C; j++) {
        float ip = dot8(xi, cent + j * d, d);
        float dis = cn[j] - 2.f * ip; /* l2SelectMin*: ||c||^2 + (-2 q.c)  (L2Select.cu:66-90,146-164) */
        topk_push(buf, &cnt, k, dis, j);
      }
      for (int r = 0; r < k; r++) {
        if (r < cnt) {
          outD[i * k + r] = buf[r].v + xn; /* sumAlongRows: + ||q||^2 on the winners only (BroadcastSum.cu:677-698) */
          outI[i * k + r] = (int)buf[r].i;
        } else {
          outD[i * k + r] = FLT_MAX;
          outI[i * k + r] = -1;
        }
      }
    }
    free(buf);
  }
  free(cn);
}

/* ---------------------------------------------------------------------------------------------------------------
 * k-means (a3): Clustering.cpp:66-206 + utils.cpp:1369-1449.
 * ------------------------------------------------------------------------------------------------------------- */
static int km_update(const float* x, float* centroids, const int* assign, int d, long k, long n) {
  long* hassign = (long*)calloc(k, sizeof(long));
  memset(centroids, 0, sizeof(float) * d * k);
  /* the reference partitions centroid ranges over threads but accumulates each centroid in row order */
  for (long i = 0; i < n; i++) {
    long ci = assign[i];
    float* c = centroids + ci * d;
    hassign[ci]++;
    for (int j = 0; j < d; j++) c[j] += x[i * d + j];
  }
  for (long ci = 0; ci < k; ci++) {
    float ni = (float)hassign[ci];
    if (ni != 0)
      for (int j = 0; j < d; j++) centroids[ci * d + j] /= ni;
  }
  int nsplit = 0;
  rng_t rng;
  rng_init(&rng, 1234);
  const float EPS = 1 / 1024.f;
  for (long ci = 0; ci < k; ci++) {
    if (hassign[ci] == 0) {
      long cj;
      for (cj = 0; 1; cj = (cj + 1) % k) {
        float p = (hassign[cj] - 1.0) / (float)(n - k);
        float r = rng_float(&rng);
        if (r < p) break;
      }
      memcpy(centroids + ci * d, centroids + cj * d, sizeof(float) * d);
      for (int j = 0; j < d; j++) {
        if (j % 2 == 0) {
          centroids[ci * d + j] *= 1 + EPS;
          centroids[cj * d + j] *= 1 - EPS;
        } else {
          centroids[ci * d + j] *= 1 - EPS;
          centroids[cj * d + j] *= 1 + EPS;
        }
      }
      hassign[ci] = hassign[cj] / 2;
      hassign[cj] -= hassign[ci];
      nsplit++;
    }
  }
  free(hassign);
  return nsplit;
}

void vlqo_kmeans(int d, int k, long n, const float* x_in, int niter, long seed, int max_points_per_centroid,
                 float* centroids, float* obj_out) {
  const float* x = x_in;
  float* x_new = NULL;
  if (n > (long)k * max_points_per_centroid) { /* Clustering.cpp:81-93 */
    int* perm = (int*)malloc(sizeof(int) * n);
    vlqo_rand_perm(perm, n, seed);
    n = (long)k * max_points_per_centroid;
    x_new = (float*)malloc(sizeof(float) * n * d);
    for (long i = 0; i < n; i++) memcpy(x_new + i * d, x_in + (long)perm[i] * d, sizeof(float) * d);
    x = x_new;
    free(perm);
  }
  {
    int* perm = (int*)malloc(sizeof(int) * n); /* Clustering.cpp:132-141 */
    vlqo_rand_perm(perm, n, seed + 1);
    for (int i = 0; i < k; i++) memcpy(centroids + (long)i * d, x + (long)perm[i] * d, sizeof(float) * d);
    free(perm);
  }
  int* assign = (int*)malloc(sizeof(int) * n);
  float* dis = (float*)malloc(sizeof(float) * n);
  for (int it = 0; it < niter; it++) { /* Clustering.cpp:161-193 */
    vlqo_l2_topk(x, n, d, centroids, k, 1, 1, dis, assign);
    float err = 0;
    for (long j = 0; j < n; j++) err += dis[j];
    if (obj_out) obj_out[it] = err;
    km_update(x, centroids, assign, d, k, n);
  }
  free(assign);
  free(dis);
  free(x_new);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Centroid kNN graph (a4): gpu/GpuIndexFlat.cu:869-893.
 * ------------------------------------------------------------------------------------------------------------- */
void vlqo_knn_graph(const float* cent, long C, int d, int E, int* edge, float* edge_d2) {
  int k = E + 1;
  float* D = (float*)malloc(sizeof(float) * C * k);
  int* I = (int*)malloc(sizeof(int) * C * k);
  vlqo_l2_topk(cent, C, d, cent, C, k, 1, D, I);
  for (long i = 0; i < C; i++)
    for (int e = 0; e < E; e++) { /* drop rank 0 (assumed to be the centroid itself) */
      edge[i * E + e] = I[i * k + 1 + e];
      edge_d2[i * E + e] = D[i * k + 1 + e];
    }
  free(D);
  free(I);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Line stage (a5): gpu/GpuIndexFlat.cu:466-550, triangle.cuh:54-87 (project, dist2).
 * ------------------------------------------------------------------------------------------------------------- */
void vlqo_line_stage(const float* x, long n, int d, const int* A, const float* cent, const int* edge,
                     const float* edge_d2, int E, int* out_list, float* out_lambda) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) {
    const float* xi = x + i * d;
    long a_id = A[i];
    float b = l2sqr8(xi, cent + a_id * d, d); /* exact differences, GpuIndexFlat.cu:499-515 */
    int best_valid = -1, best_any = -1;
    float qv = 0, qa = 0, lv = 0, la = 0;
    for (int e = 0; e < E; e++) {
      long s = edge[a_id * E + e];
      float a = l2sqr8(xi, cent + s * d, d); /* GpuIndexFlat.cu:470-494 */
      float c2 = edge_d2[a_id * E + e];
      float lam = -0.5f * (a - b - c2) / c2;                 /* project(), triangle.cuh:86 */
      float q2 = b + lam * lam * c2 + lam * (a - b - c2);    /* dist2(),   triangle.cuh:59 */
      if (best_any < 0 || q2 < qa) {
        best_any = e;
        qa = q2;
        la = lam;
      }
      if (lam >= 0.f && lam <= 1.f && (best_valid < 0 || q2 < qv)) { /* GpuIndexFlat.cu:536-545 */
        best_valid = e;
        qv = q2;
        lv = lam;
      }
    }
    int e = best_valid >= 0 ? best_valid : best_any;
    out_list[i] = (int)(a_id * E + e); /* GpuIndexFlat.cu:548 */
    out_lambda[i] = best_valid >= 0 ? lv : la;
  }
}

void vlqo_lambda_quantize(const float* lambda, long n, const float* cb, int nL, uint8_t* out) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) { /* GpuIndexFlat.cu:579-596 */
    float best = 0;
    int bj = -1;
    for (int j = 0; j < nL; j++) {
      float t = lambda[i] - cb[j];
      float dd = t * t;
      if (bj < 0 || dd < best) {
        best = dd;
        bj = j;
      }
    }
    out[i] = (uint8_t)bj;
  }
}

void vlqo_residual(const float* x, long n, int d, const int* list, const uint8_t* lamq, const float* cb,
                   const float* cent, const int* edge, int E, float* r) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) { /* GpuIndexFlat.cu:1111-1122 */
    long a_id = list[i] / E;
    int e = list[i] % E;
    long s = edge[a_id * E + e];
    float l = cb[lamq[i]];
    for (int j = 0; j < d; j++)
      r[i * d + j] = x[i * d + j] - ((1.f - l) * cent[a_id * d + j] + l * cent[s * d + j]);
  }
}

void vlqo_pq_encode(const float* r, long n, int d, const float* pq, int M, int ksub, uint8_t* codes) {
  int dsub = d / M;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; i++) { /* ProductQuantizer.cpp:311-336 */
    for (int m = 0; m < M; m++) {
      const float* xs = r + i * d + m * dsub;
      float mind = 0;
      int best = -1;
      for (int j = 0; j < ksub; j++) {
        const float* p = pq + ((long)m * ksub + j) * dsub;
        float dis = 0;
        for (int t = 0; t < dsub; t++) {
          float df = xs[t] - p[t];
          dis += df * df;
        }
        if (best < 0 || dis < mind) {
          mind = dis;
          best = j;
        }
      }
      codes[i * M + m] = (uint8_t)best;
    }
  }
}

void vlqo_build_lists(const int* list, long n, long nlists, long* offsets, long* perm) {
  memset(offsets, 0, sizeof(long) * (nlists + 1));
  for (long i = 0; i < n; i++) offsets[list[i] + 1]++;
  for (long l = 0; l < nlists; l++) offsets[l + 1] += offsets[l];
  long* cur = (long*)malloc(sizeof(long) * nlists);
  memcpy(cur, offsets, sizeof(long) * nlists);
  for (long i = 0; i < n; i++) perm[cur[list[i]]++] = i; /* insertion order within a list */
  free(cur);
}

void vlqo_term2(const float* cent, long C, int d, const float* pq, int M, int ksub, float* T2) {
  int dsub = d / M;
  float* pn = (float*)malloc(sizeof(float) * M * ksub);
  for (int m = 0; m < M; m++)
    for (int j = 0; j < ksub; j++) {
      const float* p = pq + ((long)m * ksub + j) * dsub;
      float s = 0;
      for (int t = 0; t < dsub; t++) s += p[t] * p[t];
      pn[m * ksub + j] = s;
    }
#pragma omp parallel for schedule(static)
  for (long c = 0; c < C; c++)
    for (int m = 0; m < M; m++)
      for (int j = 0; j < ksub; j++) {
        const float* p = pq + ((long)m * ksub + j) * dsub;
        const float* cm = cent + c * d + m * dsub;
        float ip = 0;
        for (int t = 0; t < dsub; t++) ip += cm[t] * p[t];
        /* 2.0f * GEMM, then += norms (IVFPQ.cu:642-668, BroadcastSum.cu:872-889) */
        T2[(c * M + m) * ksub + j] = 2.f * ip + pn[m * ksub + j];
      }
  free(pn);
}

/* ---------------------------------------------------------------------------------------------------------------
 * Search (a11-a15).
 * ------------------------------------------------------------------------------------------------------------- */
/* in_lines == NULL: the whole query path.  in_lines != NULL ([nq][W] list ids, -1 padded at the end): the coarse top-P
 * and the line selection are skipped and exactly these lines are scanned, in this order (test infrastructure: lets a
 * test compare the scan of another implementation's line choice when that choice differs by a line-score near-tie). */
static void search_impl(const float* q, long nq, int d, const float* cent, long C, const int* edge, const float* edge_d2,
                        int E, const float* lambda_cb, int nL, const float* pq, int M, int ksub, const float* T2_in,
                        const long* offsets, const uint8_t* codes, const uint8_t* lamq, const long* ids, int P, int W, int k,
                        int cap, float* outD, long* outI, int* out_coarse, int* out_lines, long* out_nscanned,
                        const int* in_lines) {
  (void)nL;
  int dsub = d / M;
  if (P > C) P = (int)C;
  float* T2own = NULL;
  const float* T2 = T2_in;
  if (!T2) {
    T2own = (float*)malloc(sizeof(float) * C * M * ksub);
    vlqo_term2(cent, C, d, pq, M, ksub, T2own);
    T2 = T2own;
  }
  float* cn = (float*)malloc(sizeof(float) * C);
#pragma omp parallel for
  for (long j = 0; j < C; j++) cn[j] = dot8(cent + j * d, cent + j * d, d);

#pragma omp parallel
  {
    float* D = (float*)malloc(sizeof(float) * C);
    cand_t* cbuf = (cand_t*)malloc(sizeof(cand_t) * (P + 1));
    cand_t* lbuf = (cand_t*)malloc(sizeof(cand_t) * (W + 1));
    cand_t* kbuf = (cand_t*)malloc(sizeof(cand_t) * (k + 1));
    long* lc = (long*)malloc(sizeof(long) * W); /* centroid and edge of every selected line */
    int* le = (int*)malloc(sizeof(int) * W);
    float* T3 = (float*)malloc(sizeof(float) * M * ksub);
    float* t23 = (float*)malloc(sizeof(float) * M * ksub);
    float* t4 = (float*)malloc(sizeof(float) * M * ksub);
#pragma omp for schedule(dynamic, 4)
    for (long qi = 0; qi < nq; qi++) {
      const float* qv = q + qi * d;
      /* coarse: D = ||c||^2 - 2 q.c, NO ||q||^2 (Distance.cu:287-290,352-373; L2Select.cu:146-164) */
      int ccnt = 0;
      for (long j = 0; j < C; j++) {
        D[j] = cn[j] - 2.f * dot8(qv, cent + j * d, d);
        if (!in_lines) topk_push(cbuf, &ccnt, P, D[j], j);
      }
      if (out_coarse)
        for (int p = 0; p < P; p++) out_coarse[qi * P + p] = p < ccnt ? (int)cbuf[p].i : -1;
      /* line scoring over the P*E lines, flat index i = p*E + e (BroadcastSum.cu:505-520) */
      int lcnt = 0;
      for (int p = 0; !in_lines && p < ccnt; p++) {
        long c = cbuf[p].i;
        for (int e = 0; e < E; e++) {
          long s = edge[c * E + e];
          float a2 = D[s], b2 = D[c], c2 = edge_d2[c * E + e];
          float v = a2 - b2;
          v -= c2;
          float score = (v > 0) ? b2 : (b2 - 0.25f * v * v / c2);
          topk_push(lbuf, &lcnt, W, score, (long)p * E + e);
        }
      }
      if (in_lines) {
        for (int w = 0; w < W && in_lines[qi * W + w] >= 0; w++) {
          lc[w] = in_lines[qi * W + w] / E;
          le[w] = in_lines[qi * W + w] % E;
          lcnt = w + 1;
        }
      } else {
        for (int w = 0; w < lcnt; w++) {
          lc[w] = cbuf[lbuf[w].i / E].i;
          le[w] = (int)(lbuf[w].i % E);
        }
      }
      /* term3 = -2 q_m . p_mj (IVFPQ.cu:1409-1432) */
      for (int m = 0; m < M; m++)
        for (int j = 0; j < ksub; j++) {
          const float* p = pq + ((long)m * ksub + j) * dsub;
          float ip = 0;
          for (int t = 0; t < dsub; t++) ip += qv[m * dsub + t] * p[t];
          T3[m * ksub + j] = -2.f * ip;
        }
      int kcnt = 0;
      long pos = 0; /* position in the query's concatenated candidate stream (IVFUtils.cu:72-90,133-169) */
      for (int w = 0; w < W; w++) {
        if (w >= lcnt) {
          if (out_lines) out_lines[qi * W + w] = -1;
          continue;
        }
        int e = le[w];
        long c = lc[w];
        long s = edge[c * E + e];
        long list = c * E + e; /* BroadcastSum.cu:552 */
        if (out_lines) out_lines[qi * W + w] = (int)list;
        float term1 = D[c];          /* b2            */
        float term6 = D[s] - D[c];   /* a2 - b2       */
        float term5 = edge_d2[c * E + e];
        const float* T2c = T2 + c * M * ksub;
        const float* T2s = T2 + s * M * ksub;
        long len = offsets[list + 1] - offsets[list];
        long limit = len < cap ? len : cap; /* IVFUtils.cu:87, PQScanMultiPassPrecomputed.cu:728 */
        if (limit == 0) continue;
        for (int t = 0; t < M * ksub; t++) {
          t23[t] = T2c[t] + T3[t];  /* loadPrecomputedTerm       (PQScan...cu:753-757) */
          t4[t] = T2s[t] - T2c[t];  /* loadPrecomputedTermGraph  (PQScan...cu:758-761) */
        }
        for (long r = 0; r < limit; r++) {
          long ent = offsets[list] + r;
          float la = lambda_cb[lamq[ent]];
          float dist = term1 + la * term6 + (la * la - la) * term5; /* PQScan...cu:779-780 */
          float tmp = 0;
          for (int m = 0; m < M; m++) {
            int code = codes[ent * M + m];
            dist += t23[m * ksub + code];
            tmp += t4[m * ksub + code];
          }
          float outv = dist + la * tmp; /* PQScan...cu:811 */
          /* candidates keyed by stream position so ties resolve like the reference's two-pass select; the id is
           * recovered afterwards (IVFUtilsSelect2.cu:446-497) */
          topk_push(kbuf, &kcnt, k, outv, pos + r);
        }
        pos += limit;
      }
      if (out_nscanned) out_nscanned[qi] = pos;
      /* map stream positions back to entries */
      {
        long pos2 = 0;
        /* small k: resolve each winner by walking the selected lists again */
        for (int r = 0; r < k; r++) {
          outD[qi * k + r] = FLT_MAX;
          outI[qi * k + r] = -1;
        }
        for (int w = 0; w < lcnt; w++) {
          long list = lc[w] * E + le[w];
          long len = offsets[list + 1] - offsets[list];
          long limit = len < cap ? len : cap;
          for (int r = 0; r < kcnt; r++)
            if (kbuf[r].i >= pos2 && kbuf[r].i < pos2 + limit) {
              outD[qi * k + r] = kbuf[r].v;
              outI[qi * k + r] = ids[offsets[list] + (kbuf[r].i - pos2)];
            }
          pos2 += limit;
        }
      }
    }
    free(D);
    free(cbuf);
    free(lbuf);
    free(kbuf);
    free(lc);
    free(le);
    free(T3);
    free(t23);
    free(t4);
  }
  free(cn);
  free(T2own);
}

void vlqo_search(const float* q, long nq, int d, const float* cent, long C, const int* edge, const float* edge_d2,
                 int E, const float* lambda_cb, int nL, const float* pq, int M, int ksub, const float* T2_in,
                 const long* offsets, const uint8_t* codes, const uint8_t* lamq, const long* ids, int P, int W, int k,
                 int cap, float* outD, long* outI, int* out_coarse, int* out_lines, long* out_nscanned) {
  search_impl(q, nq, d, cent, C, edge, edge_d2, E, lambda_cb, nL, pq, M, ksub, T2_in, offsets, codes, lamq, ids, P, W, k,
              cap, outD, outI, out_coarse, out_lines, out_nscanned, NULL);
}

void vlqo_scan_lines(const float* q, long nq, int d, const float* cent, long C, const int* edge, const float* edge_d2,
                     int E, const float* lambda_cb, int nL, const float* pq, int M, int ksub, const float* T2_in,
                     const long* offsets, const uint8_t* codes, const uint8_t* lamq, const long* ids,
                     const int* lines, int W, int k, int cap, float* outD, long* outI) {
  search_impl(q, nq, d, cent, C, edge, edge_d2, E, lambda_cb, nL, pq, M, ksub, T2_in, offsets, codes, lamq, ids, 1, W, k,
              cap, outD, outI, NULL, NULL, NULL, lines);
}

void vlqo_merge_topk(const float* D, const long* I, int R, long nq, int k, float* outD, long* outI) {
#pragma omp parallel
  {
    cand_t* buf = (cand_t*)malloc(sizeof(cand_t) * (k + 1));
#pragma omp for
    for (long qi = 0; qi < nq; qi++) { /* mergekernel, gpu/GpuIndexIVFPQ.cu:1491-1515: flat index i = rank*k + pos */
      int cnt = 0;
      for (int r = 0; r < R; r++)
        for (int p = 0; p < k; p++) topk_push(buf, &cnt, k, D[((long)r * nq + qi) * k + p], (long)r * k + p);
      for (int t = 0; t < k; t++) {
        long fi = buf[t].i;
        outD[qi * k + t] = buf[t].v;
        outI[qi * k + t] = I[((fi / k) * nq + qi) * k + (fi % k)];
      }
    }
    free(buf);
  }
}

double vlqo_decode_distance(const float* q, int d, const float* c, const float* s, float l, const float* pq, int M,
                            int ksub, const uint8_t* code) {
  int dsub = d / M;
  double acc = 0, qn = 0;
  for (int j = 0; j < d; j++) {
    int m = j / dsub, t = j % dsub;
    double anchor = (1.0 - (double)l) * c[j] + (double)l * s[j];
    double y = anchor + pq[((long)m * ksub + code[m]) * dsub + t];
    double df = (double)q[j] - y;
    acc += df * df;
    qn += (double)q[j] * q[j];
  }
  return acc - qn;
}

/* ------------------------------------------------------------------------------------------------------------------
 * IMI coarse quantizer (next row f3): see vlq_oracle.h.  Follows IndexPQ.cpp:813-855 / 636-778.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
  float sum;
  long pos; /* sum_m t_m * ksub^m, t_m = rank in the sorted table of sub-space m */
} imi_node_t;

static inline int imi_less(const imi_node_t* a, const imi_node_t* b) {
  return a->sum < b->sum || (a->sum == b->sum && a->pos < b->pos);
}
static void imi_heap_push(imi_node_t* h, int* n, imi_node_t v) {
  int i = (*n)++;
  while (i > 0) {
    int p = (i - 1) / 2;
    if (!imi_less(&v, &h[p])) break;
    h[i] = h[p];
    i = p;
  }
  h[i] = v;
}
static imi_node_t imi_heap_pop(imi_node_t* h, int* n) {
  imi_node_t top = h[0], v = h[--(*n)];
  int i = 0;
  for (;;) {
    int c = 2 * i + 1;
    if (c >= *n) break;
    if (c + 1 < *n && imi_less(&h[c + 1], &h[c])) c++;
    if (!imi_less(&h[c], &v)) break;
    h[i] = h[c];
    i = c;
  }
  h[i] = v;
  return top;
}
/* open-addressing set of visited grid positions */
static int imi_seen_add(long* set, int cap, long pos) {
  unsigned long hsh = (unsigned long)pos * 0x9E3779B97F4A7C15ul;
  int i = (int)(hsh % (unsigned long)cap);
  while (set[i] != -1) {
    if (set[i] == pos) return 0;
    i = i + 1 == cap ? 0 : i + 1;
  }
  set[i] = pos;
  return 1;
}
typedef struct {
  float v;
  int j;
} imi_ent_t;
static int imi_ent_cmp(const void* a, const void* b) {
  const imi_ent_t *x = (const imi_ent_t*)a, *y = (const imi_ent_t*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return x->j - y->j;
}

void vlqo_imi_search(const float* x, long n, int d, const float* cent, int M, int ksub, int k, float* D, long* I) {
  const int dsub = d / M;
#pragma omp parallel
  {
    float* tab = (float*)malloc(sizeof(float) * (size_t)M * ksub);
    imi_ent_t* srt = (imi_ent_t*)malloc(sizeof(imi_ent_t) * (size_t)M * ksub);
    const int hcap = k * M + M + 1, scap = 4 * (k * M + M) + 7;
    imi_node_t* heap = (imi_node_t*)malloc(sizeof(imi_node_t) * hcap);
    long* seen = (long*)malloc(sizeof(long) * scap);
#pragma omp for
    for (long i = 0; i < n; i++) {
      const float* xi = x + i * d;
      for (int m = 0; m < M; m++) /* ProductQuantizer::compute_distance_table, direct differences */
        for (int j = 0; j < ksub; j++) tab[(size_t)m * ksub + j] = l2sqr8(xi + m * dsub, cent + ((size_t)m * ksub + j) * dsub, dsub);
      if (k == 1) { /* IndexPQ.cpp:823-846 */
        float dis = 0.f;
        long label = 0, w = 1;
        for (int m = 0; m < M; m++) {
          float vmin = HUGE_VALF;
          long lmin = -1;
          for (int j = 0; j < ksub; j++)
            if (tab[(size_t)m * ksub + j] < vmin) {
              vmin = tab[(size_t)m * ksub + j];
              lmin = j;
            }
          dis += vmin;
          label += lmin * w;
          w *= ksub;
        }
        D[i] = dis;
        I[i] = label;
        continue;
      }
      for (int m = 0; m < M; m++) {
        for (int j = 0; j < ksub; j++) {
          srt[(size_t)m * ksub + j].v = tab[(size_t)m * ksub + j];
          srt[(size_t)m * ksub + j].j = j;
        }
        qsort(srt + (size_t)m * ksub, ksub, sizeof(imi_ent_t), imi_ent_cmp);
      }
      int hn = 0;
      for (int t = 0; t < scap; t++) seen[t] = -1;
      imi_node_t first = {0.f, 0};
      for (int m = 0; m < M; m++) first.sum += srt[(size_t)m * ksub].v;
      imi_heap_push(heap, &hn, first);
      imi_seen_add(seen, scap, 0);
      for (int r = 0; r < k; r++) {
        if (hn == 0) { /* fewer than k cells exist */
          D[i * k + r] = FLT_MAX;
          I[i * k + r] = -1;
          continue;
        }
        imi_node_t cur = imi_heap_pop(heap, &hn);
        long label = 0, w = 1, pp = cur.pos;
        for (int m = 0; m < M; m++) { /* grid position -> cell label through the sort permutations */
          int t = (int)(pp % ksub);
          pp /= ksub;
          label += (long)srt[(size_t)m * ksub + t].j * w;
          w *= ksub;
        }
        D[i * k + r] = cur.sum;
        I[i * k + r] = label;
        w = 1;
        pp = cur.pos;
        for (int m = 0; m < M; m++) { /* followers: one step along every axis (MinSumK::enqueue_follower) */
          int t = (int)(pp % ksub);
          pp /= ksub;
          if (t + 1 < ksub) {
            long npos = cur.pos + w;
            if (imi_seen_add(seen, scap, npos)) {
              imi_node_t nx;
              nx.pos = npos;
              nx.sum = cur.sum + (srt[(size_t)m * ksub + t + 1].v - srt[(size_t)m * ksub + t].v); /* sum + get_diff */
              imi_heap_push(heap, &hn, nx);
            }
          }
          w *= ksub;
        }
      }
    }
    free(tab);
    free(srt);
    free(heap);
    free(seen);
  }
}
