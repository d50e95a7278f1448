// Exact block-wide top-k (k <= 1024) on 64-bit keys = (order-preserving float bits << 32) | payload.
//
// Replaces the reference's BlockSelect/WarpSelect (gpu/utils/Select.cuh:77-277, MergeNetwork*.cuh) and the CPU heaps
// (Heap.h:89-323).  Design: threshold filter + shared-memory pending buffer + block bitonic sort on overflow.
// Because the key embeds the payload (column / stream position), ties resolve deterministically to the lowest
// payload, independent of thread scheduling -- the reference leaves tie order unspecified (TestGpuSelect.cu:84-112).
#pragma once
#include "common.cuh"

namespace vlq {

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// shared-memory footprint (bytes) of a BlockTopK for a given k and thread count
__host__ __device__ inline int topk_sort_size(int k, int threads) {
  return next_pow2(k + (2 * threads > k ? 2 * threads : k));
}
__host__ __device__ inline size_t topk_smem_bytes(int k, int threads) {
  return sizeof(uint64_t) * topk_sort_size(k, threads) + 16;
}

template <int THREADS>
struct BlockTopK {
  uint64_t* keys;  // [S]; [0,k) current best ascending, [k,S) pending
  int* cnt;        // pending count
  int S, k, PB;

  // smem must hold topk_smem_bytes(k, THREADS) bytes, 8-byte aligned
  __device__ void init(void* smem, int k_) {
    k = k_;
    S = topk_sort_size(k_, THREADS);
    PB = S - k;
    keys = reinterpret_cast<uint64_t*>(smem);
    cnt = reinterpret_cast<int*>(keys + S);
    for (int i = threadIdx.x; i < S; i += THREADS) keys[i] = kKeyInf;
    if (threadIdx.x == 0) *cnt = 0;
    __syncthreads();
  }

  __device__ __forceinline__ uint64_t threshold() const { return keys[k - 1]; }

  __device__ void sort_all() {
    for (int k2 = 2; k2 <= S; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (S >> 1); t += THREADS) {
          int i = 2 * t - (t & (j - 1));  // index with bit j cleared
          int l = i | j;
          bool up = (i & k2) == 0;
          uint64_t a = keys[i], b = keys[l];
          if ((a > b) == up) {
            keys[i] = b;
            keys[l] = a;
          }
        }
        __syncthreads();
      }
    }
  }

  __device__ void flush() {
    sort_all();
    for (int i = k + threadIdx.x; i < S; i += THREADS) keys[i] = kKeyInf;
    if (threadIdx.x == 0) *cnt = 0;
    __syncthreads();
  }

  // Collective: every thread of the block calls this the same number of times; at most one candidate per call.
  __device__ __forceinline__ void add(bool valid, uint64_t key) {
    if (valid && key < keys[k - 1]) {
      int slot = atomicAdd(cnt, 1);
      keys[k + slot] = key;  // slot < PB is guaranteed by the flush rule below
    }
    __syncthreads();
    int c = *cnt;
    __syncthreads();
    if (c > PB - THREADS) flush();
  }

  // Collective: after this, keys[0..k) hold the k smallest keys ascending (kKeyInf padded).
  __device__ void finish() {
    __syncthreads();
    if (*cnt > 0) {
      __syncthreads();
      flush();
    } else {
      __syncthreads();
    }
  }
};

}  // namespace vlq
