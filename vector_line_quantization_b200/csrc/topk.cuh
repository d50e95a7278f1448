// Exact block-wide k-selection (k <= 1024) on 64-bit keys = (order-preserving float bits << 32) | payload.
//
// Replaces the reference's BlockSelect / WarpSelect thread-queue + bitonic merge networks (gpu/utils/Select.cuh:77-277,
// MergeNetwork*.cuh) and the CPU heaps (Heap.h:89-323).  Design:
//   * candidates that beat the current threshold are appended to a shared-memory buffer (one atomicAdd each);
//     __syncthreads_count gives every thread the same conservative fill level with ONE barrier per batch;
//   * when the buffer could overflow it is compacted to the k smallest keys by an 8-pass byte-wise RADIX SELECT
//     (256-bin shared histogram per pass, warp-shuffle prefix scan to find the digit of the k-th key) -- O(n) work
//     instead of the O(n log^2 n) sorting network a flush used to cost; the k-th key becomes the new threshold;
//   * finish(): one last radix select, then a bitonic sort of just the k survivors.
// Keys are unique (the payload is the column / stream position), so ties resolve deterministically to the lowest
// payload, independent of thread scheduling; the reference leaves tie order unspecified (TestGpuSelect.cu:84-112).
#pragma once
#include "common.cuh"

namespace vlq {

__host__ __device__ inline int next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

constexpr int kSelMaxItems = 16;  // buffer capacity <= 16 * THREADS

// buffer capacity for a selection of k out of (at most) `total` candidates, appended in batches of `batch` per thread
__host__ __device__ inline int select_capacity(int k, int threads, int batch, long long total) {
  long long want = total < 2048 ? total : 2048;  // small buffer: more CTAs per SM, and the first compaction sets a
                                                 // threshold early so that most later candidates are never stored
  int lo = k + batch * threads;  // one full batch must always fit behind k survivors
  if (want < lo) want = lo;
  int cap = next_pow2((int)want);
  if (cap > kSelMaxItems * threads) cap = kSelMaxItems * threads;
  return cap;
}
__host__ __device__ inline size_t select_smem_bytes(int cap) { return sizeof(uint64_t) * cap + sizeof(int) * (256 + 8); }  // keys | hist[256] (8-byte aligned) | meta[8]

// Bitonic sort of 32*R keys by ONE warp: lane l holds keys l, l+32, ... in registers; partners at distance < 32 are
// exchanged with shuffles, larger distances are register-to-register.  No block barriers.
template <int R>
__device__ __forceinline__ void warp_bitonic_sort(uint64_t* keys) {
  const int lane = threadIdx.x & 31;
  uint64_t v[R];
#pragma unroll
  for (int r = 0; r < R; r++) v[r] = keys[lane + 32 * r];
#pragma unroll
  for (int k2 = 2; k2 <= 32 * R; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
#pragma unroll
        for (int r = 0; r < R; r++) {
          if ((r & (j >> 5)) == 0) {
            const bool up = ((lane + 32 * r) & k2) == 0;
            const uint64_t a = v[r], b = v[r | (j >> 5)];
            if ((a > b) == up) {
              v[r] = b;
              v[r | (j >> 5)] = a;
            }
          }
        }
      } else {
        const bool lower = (lane & j) == 0;
#pragma unroll
        for (int r = 0; r < R; r++) {
          const uint64_t pth = __shfl_xor_sync(kFull, v[r], j);
          const bool up = ((lane + 32 * r) & k2) == 0;
          const bool take_min = lower == up;
          v[r] = take_min ? (pth < v[r] ? pth : v[r]) : (pth > v[r] ? pth : v[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; r++) keys[lane + 32 * r] = v[r];
}

// BAR = 0: the whole CTA takes part (__syncthreads).  BAR > 0: the selection is run by the first THREADS threads of the
// CTA only, synchronised through named barrier BAR (the streaming scan keeps its TMA producer warp out of it).
template <int THREADS, int BAR = 0>
struct BlockSelect {
  static __device__ __forceinline__ void bsync() {
    if constexpr (BAR == 0) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"n"(BAR), "n"(THREADS) : "memory");
  }
  static __device__ __forceinline__ int bsync_count(bool p) {
    if constexpr (BAR == 0) {
      return __syncthreads_count(p);
    } else {
      int r;
      asm volatile(
          "{\n\t"
          ".reg .pred q;\n\t"
          "setp.ne.b32 q, %3, 0;\n\t"
          "bar.red.popc.u32 %0, %1, %2, q;\n\t"
          "}"
          : "=r"(r)
          : "n"(BAR), "n"(THREADS), "r"((int)p)
          : "memory");
      return r;
    }
  }
  static __device__ __forceinline__ bool bsync_or(bool p) { return bsync_count(p) != 0; }

  uint64_t* keys;  // [cap]
  int* hist;       // [256]
  int* meta;       // [0] append cursor, [1] chosen digit, [2] remaining rank
  int cap, k, batch;
  int fill;        // conservative fill level, identical in every thread
  uint64_t thr;    // only keys < thr can still enter the result
  float thr_f;     // float view of thr: one FSETP rejects almost every candidate before a key is even built

  // smem: select_smem_bytes(cap) bytes, 8-byte aligned.  cap >= k + batch*THREADS (see select_capacity).
  __device__ void init(void* smem, int k_, int cap_, int batch_) {
    k = k_;
    cap = cap_;
    batch = batch_;
    keys = reinterpret_cast<uint64_t*>(smem);
    hist = reinterpret_cast<int*>(keys + cap);
    meta = hist + 256;
    fill = 0;
    thr = kKeyInf;
    thr_f = __int_as_float(0x7f800000);
    if (threadIdx.x == 0) meta[0] = 0;
    bsync();
  }

  // adopt a buffer (same layout as init) that other code has filled: meta[0] holds the number of keys in keys[0..)
  __device__ void attach(void* smem, int k_, int cap_) {
    k = k_;
    cap = cap_;
    batch = 1;
    keys = reinterpret_cast<uint64_t*>(smem);
    hist = reinterpret_cast<int*>(keys + cap);
    meta = hist + 256;
    fill = cap;
    thr = kKeyInf;
    thr_f = __int_as_float(0x7f800000);
  }

  // non-collective part of a batch: call up to `batch` times per thread, then end_batch() once (collective)
  __device__ __forceinline__ bool offer(bool valid, uint64_t key) {
    if (valid && key < thr) {
      const int slot = atomicAdd(&meta[0], 1);
      keys[slot] = key;
      return true;
    }
    return false;
  }
  // same, from the raw value: the cheap float compare comes first (NaN values never enter)
  __device__ __forceinline__ bool offer_f(bool valid, float v, uint32_t payload) {
    if (valid && v <= thr_f) {
      const uint64_t key = make_key(v, payload);
      if (key < thr) {
        const int slot = atomicAdd(&meta[0], 1);
        keys[slot] = key;
        return true;
      }
    }
    return false;
  }
  // ---- direct placement: when ALL candidates are known up front and fit the buffer (total <= cap) every thread stores
  // its candidates at their own index -- no atomics, no per-batch barriers -- and one radix select in finish() does the
  // rest.  put() for every index in [0, total), then placed(total); NaN values must be offered as invalid.  Invalid
  // slots hold kKeyInf; placed() reports whether there were any (they would defeat the byte-skipping of the radix
  // select), in which case the caller re-runs the selection through offer_f()/end_batch() after a fresh init().
  __device__ __forceinline__ bool put(int idx, bool valid, float v, uint32_t payload) {
    valid = valid && v == v;
    keys[idx] = valid ? make_key(v, payload) : kKeyInf;
    return !valid;
  }
  __device__ __forceinline__ bool placed(int total, bool any_invalid) {
    const bool bad = bsync_or(any_invalid);
    if (threadIdx.x == 0) meta[0] = total;
    fill = total;
    bsync();
    return !bad;
  }

  // collective; `any` = this thread appended at least one key in the batch
  __device__ __forceinline__ void end_batch(bool any) {
    fill += batch * bsync_count(any);  // barrier + identical conservative count in every thread
    if (fill > cap - batch * THREADS) compact();
  }

  // collective: keep the k smallest keys of keys[0..n) (unsorted) in keys[0..k), thr = k-th smallest
  __device__ void compact() {
    bsync();
    const int n = meta[0];
    if (n <= k) {  // nothing to drop; the conservative counter was too pessimistic
      fill = n;
      bsync();
      return;
    }
    // ---- radix select of the k-th smallest key, most significant byte first.
    // (a) the bytes above the first one in which min and max differ are skipped (clustered keys would serialise on one
    // shared-memory counter there); (b) as soon as the wanted key is alone in its bin it is looked up directly (the
    // payload bytes rarely need their own passes).  (__match_any_sync aggregation was measured: slower.)
    const uint64_t key0 = keys[0];
    unsigned dhi = 0, dlo = 0;  // OR of (key ^ key0): the highest set bit marks the most significant varying byte
    for (int i = threadIdx.x; i < n; i += THREADS) {
      const uint64_t x = keys[i] ^ key0;
      dhi |= (unsigned)(x >> 32);
      dlo |= (unsigned)x;
    }
    dhi = __reduce_or_sync(kFull, dhi);
    dlo = __reduce_or_sync(kFull, dlo);
    if (threadIdx.x < 2) hist[threadIdx.x] = 0;
    bsync();
    if ((threadIdx.x & 31) == 0) {
      if (dhi) atomicOr(reinterpret_cast<unsigned*>(&hist[0]), dhi);
      if (dlo) atomicOr(reinterpret_cast<unsigned*>(&hist[1]), dlo);
    }
    bsync();
    const uint64_t diff = ((uint64_t)(unsigned)hist[0] << 32) | (unsigned)hist[1];
    bsync();
    const int top = diff ? (63 - __clzll((long long)diff)) >> 3 : 0;  // most significant byte that varies
    uint64_t prefix = top == 7 ? 0 : (key0 >> ((top + 1) * 8));
    int need = k;  // rank (1-based) of the wanted key among the keys matching `prefix`
    uint64_t kth = 0;
    bool found = false;
#pragma unroll 1
    for (int pass = top; pass >= 0; pass--) {
      hist[threadIdx.x & 255] = 0;
      if (THREADS < 256)
        for (int j = threadIdx.x; j < 256; j += THREADS) hist[j] = 0;
      bsync();
      const int shift = pass * 8;
      for (int i = threadIdx.x; i < n; i += THREADS) {
        const uint64_t key = keys[i];
        const bool match = pass == 7 ? true : ((key >> (shift + 8)) == prefix);
        if (match) atomicAdd(&hist[(int)((key >> shift) & 255)], 1);
      }
      bsync();
      if (threadIdx.x < kWarp) {  // lane l owns bins [8l, 8l+8)
        const int lane = threadIdx.x;
        int c[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
          c[j] = hist[lane * 8 + j];
          s += c[j];
        }
        int inc = s;
#pragma unroll
        for (int o = 1; o < kWarp; o <<= 1) {
          int t = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += t;
        }
        int before = inc - s;  // keys in bins below this lane's range
        if (before < need && need <= inc) {  // the wanted key lies in this lane's bins
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (need <= before + c[j]) {
              meta[1] = lane * 8 + j;
              meta[2] = need - before;
              meta[3] = c[j];
              break;
            }
            before += c[j];
          }
        }
      }
      bsync();
      prefix = (prefix << 8) | (uint64_t)meta[1];
      need = meta[2];
      const int in_bin = meta[3];
      if (in_bin == 1 && pass > 0) {  // block-uniform: the wanted key is the only one with this prefix
        for (int i = threadIdx.x; i < n; i += THREADS) {
          const uint64_t key = keys[i];
          if ((key >> shift) == prefix) *reinterpret_cast<uint64_t*>(hist) = key;
        }
        bsync();
        kth = *reinterpret_cast<uint64_t*>(hist);
        found = true;
        bsync();
        break;
      }
    }
    if (!found) kth = prefix;
    // exactly k keys are <= kth (keys are unique)
    // ---- compaction through registers (in-place writes would race with other threads' reads)
    uint64_t mine[kSelMaxItems];
#pragma unroll
    for (int t = 0; t < kSelMaxItems; t++) {
      if (t * THREADS >= n) break;  // block-uniform
      const int i = threadIdx.x + t * THREADS;
      mine[t] = i < n ? keys[i] : kKeyInf;
    }
    bsync();
    if (threadIdx.x == 0) meta[0] = 0;
    bsync();
#pragma unroll
    for (int t = 0; t < kSelMaxItems; t++) {
      if (t * THREADS >= n) break;
      if (mine[t] <= kth) {
        const int slot = atomicAdd(&meta[0], 1);
        keys[slot] = mine[t];
      }
    }
    thr = kth;
    thr_f = key_val(kth);
    fill = k;
    bsync();
  }

  // collective: afterwards keys[0..k) hold the k smallest keys ascending, padded with kKeyInf
  __device__ void finish(bool sorted = true) {
    compact();
    const int n = meta[0] < k ? meta[0] : k;
    int S = next_pow2(k);  // S <= cap because cap >= k + batch*THREADS and cap is a power of two
    if (S < 32) S = 32;
    for (int i = n + threadIdx.x; i < S; i += THREADS) keys[i] = kKeyInf;
    bsync();
    if (!sorted) return;  // the caller only needs the SET of the k smallest keys
    if (S <= 64) {  // one warp sorts in registers (no barriers per step); measured slower than the block network for S >= 128
      if (threadIdx.x < kWarp) {
        if (S == 32) warp_bitonic_sort<1>(keys);
        else warp_bitonic_sort<2>(keys);
      }
      bsync();
      return;
    }
    for (int k2 = 2; k2 <= S; k2 <<= 1) {
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        for (int t = threadIdx.x; t < (S >> 1); t += THREADS) {
          const int i = 2 * t - (t & (j - 1));
          const int l = i | j;
          const bool up = (i & k2) == 0;
          const uint64_t a = keys[i], b = keys[l];
          if ((a > b) == up) {
            keys[i] = b;
            keys[l] = a;
          }
        }
        bsync();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------------
// Warp-scope twin of BlockSelect for k <= 128: one 256-key buffer + one 256-bin histogram per WARP in shared memory.
// Appends are ballot-compacted (no atomics), the compaction is the same byte-wise radix select with __syncwarp()
// instead of __syncthreads().  A warp that owns one of these never waits for another warp, which is what lets the
// 1B-scale scan keep loads of many lists in flight (the block-wide barrier per batch of BlockSelect exposes one DRAM
// round trip per batch).  The per-warp survivors are merged by a BlockSelect at the very end.
constexpr int kWarpSelCap = 256;
constexpr int kWarpSelMaxK = 128;
constexpr int kWarpSelSmemBytes = kWarpSelCap * 8 + 256 * 4;  // keys | hist

struct WarpSelect {
  uint64_t* keys;  // [256]
  int* hist;       // [256]
  int cnt;         // keys in the buffer (warp-uniform)
  int k;
  uint64_t thr;
  float thr_f;

  __device__ void init(void* smem, int k_) {
    keys = reinterpret_cast<uint64_t*>(smem);
    hist = reinterpret_cast<int*>(keys + kWarpSelCap);
    cnt = 0;
    k = k_;
    thr = kKeyInf;
    thr_f = __int_as_float(0x7f800000);
  }

  // warp-collective: every lane offers at most one candidate
  __device__ __forceinline__ void offer(bool valid, float v, uint32_t payload) {
    bool take = valid && v <= thr_f;
    uint64_t key = 0;
    if (take) {
      key = make_key(v, payload);
      take = key < thr;
    }
    const unsigned m = __ballot_sync(kFull, take);
    if (m) {
      const int lane = threadIdx.x & 31;
      if (take) keys[cnt + __popc(m & ((1u << lane) - 1))] = key;
      cnt += __popc(m);
      if (cnt > kWarpSelCap - 32) compact();
    }
  }

  // warp-collective: keep the k smallest keys (unsorted) in keys[0..k), thr = k-th smallest
  __device__ void compact() {
    __syncwarp();
    const int n = cnt;
    if (n <= k) return;
    const int lane = threadIdx.x & 31;
    constexpr int ITEMS = kWarpSelCap / 32;
    uint64_t mine[ITEMS];
    unsigned dhi = 0, dlo = 0;
    const uint64_t key0 = keys[0];
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
      const int i = lane + 32 * t;
      mine[t] = i < n ? keys[i] : kKeyInf;
      if (i < n) {
        const uint64_t x = mine[t] ^ key0;
        dhi |= (unsigned)(x >> 32);
        dlo |= (unsigned)x;
      }
    }
    dhi = __reduce_or_sync(kFull, dhi);
    dlo = __reduce_or_sync(kFull, dlo);
    const uint64_t diff = ((uint64_t)dhi << 32) | dlo;
    const int top = diff ? (63 - __clzll((long long)diff)) >> 3 : 0;
    uint64_t prefix = top == 7 ? 0 : (key0 >> ((top + 1) * 8));
    int need = k;
    uint64_t kth = 0;
    bool found = false;
#pragma unroll 1
    for (int pass = top; pass >= 0; pass--) {
#pragma unroll
      for (int j = 0; j < 8; j++) hist[lane * 8 + j] = 0;
      __syncwarp();
      const int shift = pass * 8;
#pragma unroll
      for (int t = 0; t < ITEMS; t++) {
        const bool in = lane + 32 * t < n;
        const bool match = in && (pass == 7 ? true : ((mine[t] >> (shift + 8)) == prefix));
        if (match) atomicAdd(&hist[(int)((mine[t] >> shift) & 255)], 1);
      }
      __syncwarp();
      int c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) {
        c[j] = hist[lane * 8 + j];
        sum += c[j];
      }
      int inc = sum;
#pragma unroll
      for (int o = 1; o < kWarp; o <<= 1) {
        const int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      int before = inc - sum;
      const bool here = before < need && need <= inc;  // exactly one lane
      int digit = 0, rem = 0, inbin = 0;
      if (here) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          if (need <= before + c[j]) {
            digit = lane * 8 + j;
            rem = need - before;
            inbin = c[j];
            break;
          }
          before += c[j];
        }
      }
      const int src = __ffs(__ballot_sync(kFull, here)) - 1;
      digit = __shfl_sync(kFull, digit, src);
      need = __shfl_sync(kFull, rem, src);
      inbin = __shfl_sync(kFull, inbin, src);
      prefix = (prefix << 8) | (uint64_t)digit;
      __syncwarp();
      if (inbin == 1 && pass > 0) {  // the wanted key is the only one with this prefix
        uint64_t cand = 0;
        bool hit = false;
#pragma unroll
        for (int t = 0; t < ITEMS; t++)
          if (lane + 32 * t < n && (mine[t] >> shift) == prefix) {
            cand = mine[t];
            hit = true;
          }
        const unsigned who = __ballot_sync(kFull, hit);
        kth = __shfl_sync(kFull, cand, __ffs(who) - 1);
        found = true;
        break;
      }
    }
    if (!found) kth = prefix;
    // compaction from the register copies (all lanes hold their keys already)
    int base = 0;
#pragma unroll
    for (int t = 0; t < ITEMS; t++) {
      const bool keep = lane + 32 * t < n && mine[t] <= kth;
      const unsigned m = __ballot_sync(kFull, keep);
      if (keep) keys[base + __popc(m & ((1u << lane) - 1))] = mine[t];
      base += __popc(m);
    }
    cnt = base;  // == k
    thr = kth;
    thr_f = key_val(kth);
    __syncwarp();
  }
};


// ---------------------------------------------------------------------------------------------------------------------
// ONE candidate buffer per CTA shared by autonomous warps (the long-list scans): slots are reserved with an atomicAdd,
// and the warp whose reservation crosses the capacity compacts the buffer alone (radix select of the k smallest) while
// the others keep scanning.  One threshold per query: ~k ln(n/k) insertions in total instead of that per warp.  The
// buffer has the BlockSelect layout (keys | hist[256] | meta[8]) so that BlockSelect::attach + finish ends the query.
constexpr int kSharedSelCap = 2048;
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// k smallest of keys[0..n) by ONE warp: byte-wise radix select (same scheme as WarpSelect::compact, keys streamed from
// shared memory), then an in-place stable partition.  Returns the k-th smallest key; keys[0..k) hold the survivors.
__device__ inline uint64_t warp_select_smem(uint64_t* keys, int n, int k, int* hist) {
  const int lane = threadIdx.x & 31;
  const uint64_t key0 = keys[0];
  unsigned dhi = 0, dlo = 0;
  for (int i = lane; i < n; i += 32) {
    const uint64_t x = keys[i] ^ key0;
    dhi |= (unsigned)(x >> 32);
    dlo |= (unsigned)x;
  }
  dhi = __reduce_or_sync(kFull, dhi);
  dlo = __reduce_or_sync(kFull, dlo);
  const uint64_t diff = ((uint64_t)dhi << 32) | dlo;
  const int top = diff ? (63 - __clzll((long long)diff)) >> 3 : 0;
  uint64_t prefix = top == 7 ? 0 : (key0 >> ((top + 1) * 8));
  int need = k;
  uint64_t kth = 0;
  bool found = false;
#pragma unroll 1
  for (int pass = top; pass >= 0; pass--) {
#pragma unroll
    for (int j = 0; j < 8; j++) hist[lane * 8 + j] = 0;
    __syncwarp();
    const int shift = pass * 8;
    for (int i = lane; i < n; i += 32) {
      const uint64_t key = keys[i];
      const bool match = pass == 7 ? true : ((key >> (shift + 8)) == prefix);
      if (match) atomicAdd(&hist[(int)((key >> shift) & 255)], 1);
    }
    __syncwarp();
    int c[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      c[j] = hist[lane * 8 + j];
      sum += c[j];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, inc, o);
      if (lane >= o) inc += t;
    }
    int before = inc - sum;
    const bool here = before < need && need <= inc;  // exactly one lane
    int digit = 0, rem = 0, inbin = 0;
    if (here) {
#pragma unroll
      for (int j = 0; j < 8; j++) {
        if (need <= before + c[j]) {
          digit = lane * 8 + j;
          rem = need - before;
          inbin = c[j];
          break;
        }
        before += c[j];
      }
    }
    const int src = __ffs(__ballot_sync(kFull, here)) - 1;
    digit = __shfl_sync(kFull, digit, src);
    need = __shfl_sync(kFull, rem, src);
    inbin = __shfl_sync(kFull, inbin, src);
    prefix = (prefix << 8) | (uint64_t)digit;
    __syncwarp();
    if (inbin == 1 && pass > 0) {  // the wanted key is the only one with this prefix
      uint64_t cand = 0;
      bool hit = false;
      for (int i = lane; i < n; i += 32) {
        const uint64_t key = keys[i];
        if ((key >> shift) == prefix) {
          cand = key;
          hit = true;
        }
      }
      const unsigned who = __ballot_sync(kFull, hit);
      kth = __shfl_sync(kFull, cand, __ffs(who) - 1);
      found = true;
      break;
    }
  }
  if (!found) kth = prefix;
  int base = 0;  // in-place partition: writes never pass the reads (base <= i0)
  for (int i0 = 0; i0 < n; i0 += 32) {
    const int i = i0 + lane;
    const uint64_t key = i < n ? keys[i] : kKeyInf;
    const bool keep = i < n && key <= kth;
    const unsigned m = __ballot_sync(kFull, keep);
    __syncwarp();
    if (keep) keys[base + __popc(m & ((1u << lane) - 1))] = key;
    base += __popc(m);
    __syncwarp();
  }
  return kth;
}

// meta words of the shared selection (BlockSelect uses [0..3]: [0] is the append cursor)
constexpr int META_COMMITTED = 4;
constexpr int META_THR = 5;

// warp-collective: lanes with take == true append their key
__device__ __forceinline__ void shared_offer(uint64_t* keys, int* hist, int* meta, int k, bool valid, float dist,
                                             uint32_t payload, int lane) {
  for (;;) {
    const float thr = __int_as_float(ld_volatile(&meta[META_THR]));
    const bool take = valid && dist <= thr;
    const unsigned m = __ballot_sync(kFull, take);
    if (!m) return;
    const int n = __popc(m);
    int base = 0;
    if (lane == 0) base = atomicAdd(&meta[0], n);
    base = __shfl_sync(kFull, base, 0);
    if (base + n <= kSharedSelCap) {
      if (take) keys[base + __popc(m & ((1u << lane) - 1))] = make_key(dist, payload);
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(&meta[META_COMMITTED], n);
      }
      return;
    }
    if (base <= kSharedSelCap) {
      // this warp's reservation crossed the capacity: slots [0, base) belong to other warps' appends.  Wait until they
      // are all written, keep the k smallest, publish the new threshold and reopen the buffer.
      while (ld_volatile(&meta[META_COMMITTED]) != base) __nanosleep(64);
      __threadfence_block();
      int kept = base;
      if (base > k) {
        const uint64_t kth = warp_select_smem(keys, base, k, hist);
        kept = k;
        if (lane == 0) *reinterpret_cast<volatile int*>(&meta[META_THR]) = __float_as_int(key_val(kth));
      }
      __syncwarp();
      if (lane == 0) {
        *reinterpret_cast<volatile int*>(&meta[META_COMMITTED]) = kept;
        __threadfence_block();
        atomicExch(&meta[0], kept);
      }
      __syncwarp();
    } else {
      while (ld_volatile(&meta[0]) > kSharedSelCap) __nanosleep(128);  // closed for compaction
    }
  }
}

}  // namespace vlq
