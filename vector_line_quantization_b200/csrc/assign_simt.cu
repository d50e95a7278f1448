// Exact fp32 coarse assignment on CUDA cores.
//
//   D[i][j] = ||c_j||^2 - 2 x_i . c_j          (GEMM form, as the reference: gpu/impl/Distance.cu:352-373,679-685)
//
// This is the exact path: it serves (1) dimensions the tensor-core kernel does not cover (d % 64 != 0, e.g. the 1-D
// lambda k-means and the PQ sub-space k-means), (2) the exact re-evaluation of rows the tcgen05 kernel flags as
// ambiguous, and (3) the first-correct implementation the tensor path is checked against.
//
// Kernel: 128x128 output tile per CTA, BK=16, 256 threads, 8x8 register micro-tile (split 4+4 so shared-memory
// reads are conflict free), double-buffered shared memory.  Two epilogues: store D, or fused per-row arg-min
// (the CTA then walks all centroid tiles of its 128 rows and D is never materialised).
#include "common.cuh"
#include "topk.cuh"

namespace vlq {

std::atomic<uint64_t> g_launch_count{0};

// ------------------------------------------------------------------------------------------------- row norms (a1)
__global__ void row_norms_kernel(const float* __restrict__ x, int64_t n, int d, float* __restrict__ out) {
  int64_t row = (int64_t)blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
  int lane = threadIdx.x % kWarp;
  if (row >= n) return;
  const float* xr = x + row * d;
  float acc = 0.f;
  for (int j = lane; j < d; j += kWarp) {
    float v = xr[j];
    acc = fmaf(v, v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

// ------------------------------------------------------------------------------------------------- fp32 GEMM tiles
constexpr int BM = 128, BN = 128, BK = 16, GEMM_THREADS = 256;
constexpr int LDS_PAD = 4;  // keeps float4 alignment of rows, skews banks for the transposed stores

struct TileLoader {
  // loads a (128 rows x 16 k) tile of a row-major [rows][d] matrix into registers (2 float4 per thread)
  const float* base;
  int64_t rows;
  int d;
  bool vec;
  __device__ __forceinline__ void load(int64_t row0, int k0, float4 (&r)[2]) const {
#pragma unroll
    for (int t = 0; t < 2; t++) {
      int idx = threadIdx.x + t * GEMM_THREADS;  // 0..511
      int rr = idx >> 2;                         // 0..127
      int kk = (idx & 3) * 4;                    // 0,4,8,12
      int64_t row = row0 + rr;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < rows) {
        const float* p = base + row * d + k0 + kk;
        if (vec && k0 + kk + 3 < d) {
          v = *reinterpret_cast<const float4*>(p);
        } else {
          if (k0 + kk + 0 < d) v.x = p[0];
          if (k0 + kk + 1 < d) v.y = p[1];
          if (k0 + kk + 2 < d) v.z = p[2];
          if (k0 + kk + 3 < d) v.w = p[3];
        }
      }
      r[t] = v;
    }
  }
  __device__ __forceinline__ static void store(float (*s)[BM + LDS_PAD], const float4 (&r)[2]) {
#pragma unroll
    for (int t = 0; t < 2; t++) {
      int idx = threadIdx.x + t * GEMM_THREADS;
      int rr = idx >> 2;
      int kk = (idx & 3) * 4;
      s[kk + 0][rr] = r[t].x;
      s[kk + 1][rr] = r[t].y;
      s[kk + 2][rr] = r[t].z;
      s[kk + 3][rr] = r[t].w;
    }
  }
};

// acc[i][j] += sum_k A[row_i][k] * B[col_j][k] for one 128x128 tile over the whole K range
__device__ __forceinline__ void gemm_tile(const TileLoader& la, int64_t row0, const TileLoader& lb, int64_t col0,
                                          int d, float (*As)[BK][BM + LDS_PAD], float (*Bs)[BK][BN + LDS_PAD],
                                          float (&acc)[8][8]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float4 ra[2], rb[2];
  const int nk = (d + BK - 1) / BK;
  la.load(row0, 0, ra);
  lb.load(col0, 0, rb);
  __syncthreads();  // previous tile's readers are done with buffer 0
  TileLoader::store(As[0], ra);
  TileLoader::store(Bs[0], rb);
  __syncthreads();
  for (int kt = 0; kt < nk; kt++) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      la.load(row0, (kt + 1) * BK, ra);
      lb.load(col0, (kt + 1) * BK, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      TileLoader::store(As[cur ^ 1], ra);
      TileLoader::store(Bs[cur ^ 1], rb);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ int tile_row(int ty, int i) { return (i < 4 ? 0 : 64) + ty * 4 + (i & 3); }
__device__ __forceinline__ int tile_col(int tx, int j) { return (j < 4 ? 0 : 64) + tx * 4 + (j & 3); }

// epilogue 1: store D = cnorm - 2 acc
__global__ void __launch_bounds__(GEMM_THREADS, 2)
l2_dist_store_kernel(const float* __restrict__ x, int64_t n, int d, const float* __restrict__ cent,
                     const float* __restrict__ cnorm, int C, float* __restrict__ D, int64_t ldD) {
  __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(cent)) % 16 == 0);
  TileLoader la{x, n, d, vec}, lb{cent, (int64_t)C, d, vec};
  const int64_t row0 = (int64_t)blockIdx.y * BM;
  const int64_t col0 = (int64_t)blockIdx.x * BN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
  gemm_tile(la, row0, lb, col0, d, As, Bs, acc);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    int64_t row = row0 + tile_row(ty, i);
    if (row >= n) continue;
#pragma unroll
    for (int jj = 0; jj < 2; jj++) {
      int64_t col = col0 + tile_col(tx, jj * 4);
      float v[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int64_t cj = col + j;
        v[j] = cj < C ? fmaf(-2.f, acc[i][jj * 4 + j], cnorm[cj]) : 0.f;
      }
      float* out = D + row * ldD + col;
      if (col + 3 < C && (ldD % 4 == 0) && (reinterpret_cast<uintptr_t>(D) % 16 == 0)) {
        *reinterpret_cast<float4*>(out) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (col + j < C) out[j] = v[j];
      }
    }
  }
}

// epilogue 2: fused arg-min over all centroid tiles of a 128-row block
__global__ void __launch_bounds__(GEMM_THREADS, 2)
l2_argmin_kernel(const float* __restrict__ x, int64_t n, int d, const float* __restrict__ cent,
                 const float* __restrict__ cnorm, int C, int add_xnorm, int* __restrict__ out_ids,
                 float* __restrict__ out_dist) {
  __shared__ __align__(16) float As[2][BK][BM + LDS_PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + LDS_PAD];
  const bool vec = (d % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(cent)) % 16 == 0);
  TileLoader la{x, n, d, vec}, lb{cent, (int64_t)C, d, vec};
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  uint64_t best[8];
#pragma unroll
  for (int i = 0; i < 8; i++) best[i] = kKeyInf;

  for (int64_t col0 = 0; col0 < C; col0 += BN) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 8; j++) acc[i][j] = 0.f;
    gemm_tile(la, row0, lb, col0, d, As, Bs, acc);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int64_t cj = col0 + tile_col(tx, j);
      if (cj < C) {
        float cn = cnorm[cj];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          uint64_t key = make_key(fmaf(-2.f, acc[i][j], cn), (uint32_t)cj);
          best[i] = key < best[i] ? key : best[i];
        }
      }
    }
  }
  // reduce across the 16 threads (same ty, tx = 0..15) that share these rows: they sit in one half-warp
#pragma unroll
  for (int i = 0; i < 8; i++) {
    uint64_t b = best[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      uint64_t t = __shfl_xor_sync(kFull, b, o);
      b = t < b ? t : b;
    }
    int64_t row = row0 + tile_row(ty, i);
    if (tx == 0 && row < n) {
      // a row containing NaN / Inf has no nearest centroid: id -1, the encoder then skips the vector like the reference
      // does for invalid input (gpu/GpuIndexIVFPQ.cu:751-755); same behaviour as the tcgen05 path.  Decided from the
      // row itself (the 64-bit key minimum is not a reliable NaN carrier).
      const float* xr = x + row * d;
      float xn = 0.f;
      for (int t = 0; t < d; t++) xn = fmaf(xr[t], xr[t], xn);
      const bool valid = xn < __int_as_float(0x7f800000);  // false for NaN and Inf
      out_ids[row] = (valid && b != kKeyInf) ? (int)key_payload(b) : -1;
      if (out_dist) out_dist[row] = valid ? key_val(b) + (add_xnorm ? xn : 0.f) : 3.402823466e+38f;
    }
  }
}

// ------------------------------------------------------------------------------------------------- row select (a15)
constexpr int SEL_THREADS = 256;
constexpr int SEL_BATCH = 4;
__global__ void __launch_bounds__(SEL_THREADS)
select_rows_kernel(const float* __restrict__ D, int cols, int64_t ldD, int k, int cap, const float* __restrict__ row_add,
                   float* __restrict__ out_val, int* __restrict__ out_idx) {
  extern __shared__ __align__(16) unsigned char smem[];
  BlockSelect<SEL_THREADS> sel;
  sel.init(smem, k, cap, SEL_BATCH);
  const int64_t row = blockIdx.x;
  const float* dr = D + row * ldD;
  for (int base = 0; base < cols; base += SEL_BATCH * SEL_THREADS) {
    float v[SEL_BATCH];
#pragma unroll
    for (int b = 0; b < SEL_BATCH; b++) {  // all loads of the batch in flight before the first compare
      const int j = base + b * SEL_THREADS + threadIdx.x;
      v[b] = j < cols ? dr[j] : 0.f;
    }
    bool any = false;
#pragma unroll
    for (int b = 0; b < SEL_BATCH; b++) {
      const int j = base + b * SEL_THREADS + threadIdx.x;
      any |= sel.offer_f(j < cols, v[b], (uint32_t)j);
    }
    sel.end_batch(any);
  }
  sel.finish();
  const float add = row_add ? row_add[row] : 0.f;
  for (int i = threadIdx.x; i < k; i += SEL_THREADS) {
    uint64_t key = sel.keys[i];
    bool ok = key != kKeyInf;
    out_val[row * k + i] = ok ? key_val(key) + add : 3.402823466e+38f;
    out_idx[row * k + i] = ok ? (int)key_payload(key) : -1;
  }
}

// ------------------------------------------------------------------------------------------------- small utilities
__global__ void gather_rows_kernel(const float* __restrict__ src, int d, const int64_t* __restrict__ rows, int64_t n,
                                   float* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
  if (i >= n) return;
  const float* s = src + rows[i] * d;
  float* o = dst + i * d;
  for (int j = threadIdx.x % kWarp; j < d; j += kWarp) o[j] = s[j];
}
__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, int64_t count, float* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < count; i += stride) dst[i] = (float)src[i];
}
__global__ void shift_ids_kernel(int64_t* __restrict__ ids, int64_t n, int64_t shift) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (ids[i] >= 0) ids[i] += shift;
}
__global__ void iota_i64_kernel(int64_t* dst, int64_t n, int64_t start) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = start + i;
}

}  // namespace vlq

using namespace vlq;

extern "C" {

uint64_t vlq_launch_count(void) { return g_launch_count.load(); }

int vlq_row_norms(const float* x, int64_t n, int d, float* out, vlq_stream_t stream) {
  if (n < 0 || d <= 0) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!x || !out) return VLQ_EINVAL;
  const int warps = 8;
  VLQ_LAUNCH(row_norms_kernel, (unsigned)div_up(n, warps), warps * kWarp, 0, as_stream(stream), x, n, d, out);
  return last_error();
}

int vlq_l2_assign(const float* x, int64_t n, int d, const float* cent, const float* cnorm, int C, int add_xnorm,
                  int* out_ids, float* out_dist, vlq_stream_t stream) {
  if (n < 0 || d <= 0 || C <= 0) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!x || !cent || !cnorm || !out_ids) return VLQ_EINVAL;
  VLQ_LAUNCH(l2_argmin_kernel, (unsigned)div_up(n, BM), GEMM_THREADS, 0, as_stream(stream), x, n, d, cent, cnorm, C,
             add_xnorm, out_ids, out_dist);
  return last_error();
}

int vlq_l2_distances(const float* x, int64_t n, int d, const float* cent, const float* cnorm, int C, float* D,
                     int64_t ldD, vlq_stream_t stream) {
  if (n < 0 || d <= 0 || C <= 0 || ldD < C) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!x || !cent || !cnorm || !D) return VLQ_EINVAL;
  if (div_up(n, BM) > 65535) return VLQ_EINVAL;  // callers tile queries (GpuIndex.cu:109-147 pages at 32Ki)
  dim3 grid((unsigned)div_up(C, BN), (unsigned)div_up(n, BM));
  VLQ_LAUNCH(l2_dist_store_kernel, grid, GEMM_THREADS, 0, as_stream(stream), x, n, d, cent, cnorm, C, D, ldD);
  return last_error();
}

int vlq_select_rows(const float* D, int64_t n, int cols, int64_t ldD, int k, const float* row_add, float* out_val,
                    int* out_idx, vlq_stream_t stream) {
  if (n < 0 || cols <= 0 || k <= 0 || k > VLQ_MAX_K || ldD < cols) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!D || !out_val || !out_idx) return VLQ_EINVAL;
  const int cap = select_capacity(k, SEL_THREADS, SEL_BATCH, cols);
  const size_t smem = select_smem_bytes(cap);
  VLQ_LAUNCH(select_rows_kernel, (unsigned)n, SEL_THREADS, smem, as_stream(stream), D, cols, ldD, k, cap, row_add,
             out_val, out_idx);
  return last_error();
}

int vlq_gather_rows(const float* src, int d, const int64_t* rows, int64_t n, float* dst, vlq_stream_t stream) {
  if (d <= 0 || n < 0) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!src || !rows || !dst) return VLQ_EINVAL;
  VLQ_LAUNCH(gather_rows_kernel, (unsigned)div_up(n, 8), 256, 0, as_stream(stream), src, d, rows, n, dst);
  return last_error();
}
int vlq_u8_to_f32(const uint8_t* src, int64_t count, float* dst, vlq_stream_t stream) {
  if (count < 0) return VLQ_EINVAL;
  if (count == 0) return VLQ_OK;
  if (!src || !dst) return VLQ_EINVAL;
  VLQ_LAUNCH(u8_to_f32_kernel, 148 * 8, 256, 0, as_stream(stream), src, count, dst);
  return last_error();
}
int vlq_shift_ids(int64_t* ids, int64_t n, int64_t shift, vlq_stream_t stream) {
  if (n < 0) return VLQ_EINVAL;
  if (n == 0 || shift == 0) return VLQ_OK;
  if (!ids) return VLQ_EINVAL;
  VLQ_LAUNCH(shift_ids_kernel, 148 * 4, 256, 0, as_stream(stream), ids, n, shift);
  return last_error();
}
int vlq_iota_i64(int64_t* dst, int64_t n, int64_t start, vlq_stream_t stream) {
  if (n < 0) return VLQ_EINVAL;
  if (n == 0) return VLQ_OK;
  if (!dst) return VLQ_EINVAL;
  VLQ_LAUNCH(iota_i64_kernel, 148 * 4, 256, 0, as_stream(stream), dst, n, start);
  return last_error();
}

}  // extern "C"
