// Shared device/host helpers for the VLQ sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/vlq_b200.h"

namespace vlq {

extern std::atomic<uint64_t> g_launch_count;

inline cudaStream_t as_stream(vlq_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Every launch goes through this macro so that vlq_launch_count() is exact.
#define VLQ_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);            \
    ::vlq::g_launch_count.fetch_add(1, std::memory_order_relaxed);         \
  } while (0)

#define VLQ_CUDA_TRY(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

inline int last_error() {
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? VLQ_OK : (int)e;
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int64_t div_up(int64_t a, int64_t b) { return (a + b - 1) / b; }

// float -> uint32 whose unsigned order equals the float order (handles negatives; NaN sorts last-ish)
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ uint64_t make_key(float v, uint32_t payload) {
  return (static_cast<uint64_t>(f2ord(v)) << 32) | payload;
}
__device__ __forceinline__ float key_val(uint64_t k) { return ord2f(static_cast<uint32_t>(k >> 32)); }
__device__ __forceinline__ uint32_t key_payload(uint64_t k) { return static_cast<uint32_t>(k); }

constexpr uint64_t kKeyInf = 0xffffffffffffffffull;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(kFull, v, o);
    v = t < v ? t : v;
  }
  return v;
}

// streaming 16-byte load that does not pollute L1 (codes are read once)
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_nc_v2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

}  // namespace vlq
