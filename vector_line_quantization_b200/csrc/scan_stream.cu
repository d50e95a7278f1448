// TMA-staged streaming ADC scan for long inverted lists (SURVEY.md 8a rows a14 + a15, north_star item 4).
//   replaces pqScanPrecomputedMultiPassGraph + pass1/pass2SelectLists
//   (gpu/impl/PQScanMultiPassPrecomputed.cu:675-881, IVFUtilsSelect1.cu:28-144, IVFUtilsSelect2.cu:398-569).
//
// One persistent CTA per SM; queries are pulled from a global counter.  Inside a CTA:
//
//   producer warp (the last warp, all 32 lanes)  walks the W selected lists of the current query as a stream of 64-entry
//     chunks and brings every chunk into a shared-memory ring with three cp.async.bulk copies (codes 16 B / entry, kappa
//     4 B, lambda byte 1 B; SASS UBLKCP) that complete on the slot's `full` mbarrier; it runs ahead of the consumers by
//     the depth of the ring (tens of KB per SM in flight), also across query boundaries, so the HBM stream never
//     drains while the consumers finish one query and set up the next.
//   consumer warps  take the chunks round-robin, wait on `full`, score two entries per lane against the bank-skewed
//     term-3 tables of the query (see below), release the slot (`empty` mbarrier) and offer the few candidates that beat
//     the running threshold to ONE shared candidate buffer: slots are reserved with an atomicAdd, and the warp whose
//     reservation crosses the capacity compacts the buffer alone (radix select of the k smallest) while the others keep
//     scanning.  One threshold per query: ~k ln(n/k) insertions in total instead of that per warp.
//   query boundary  markers travel through the same ring: the consumers meet on a named barrier, finish the selection
//     (BlockSelect on the shared buffer), write the k results, expand the next query's tables and go on.
//
// Tables.  T3[m][j] = -2 q_m . p_mj of the query is stored code-major with a 64-word row: word c of row `code` holds
// T3[c mod M][code] for c < 31 + M.  Lane l works on sub-quantizer (s + l) mod M at step s and reads word l + s of row
// `code`: the 32 lanes of a warp always hit 32 different banks whatever their codes are, "extract the code byte, scale
// it by the 256-byte row stride, add the lane's column" is ONE byte-permute and the step offset is an immediate --
// PRMT + LDS + FADD per lookup.  The lists store the code bytes of the entry at position pos rotated by pos mod M
// (scan.cuh), and chunks start at multiples of 64 inside a list, so byte s of the stored code is the byte lane l needs
// at step s.  The M terms are summed in a lane-dependent order: distances can differ from the block-synchronous scans
// in the last ulp.
//
// Streamed per entry: M + 1 + 4 bytes (codes, lambda byte, kappa); SURVEY 8d counts M + 1 of them as algorithmic.
#include <cfloat>
#include <cstdlib>

#include "scan.cuh"
#include "topk.cuh"

namespace vlq {
namespace stream {

constexpr int CH = 64;          // entries per chunk (two per consumer lane)
constexpr int SEL_CAP = 2048;   // shared candidate buffer (keys)
constexpr int ROW_WORDS = 64;   // table row stride in words
constexpr int BAR_CONS = 1;     // named barrier of the consumer warps

enum : int { KIND_DATA = 0, KIND_BOUNDARY = 1 };

struct __align__(16) ChunkMeta {  // 32 bytes, written by the producer before it arrives on `full`
  int kind;
  int n;     // DATA: entries in the chunk.                    BOUNDARY: query to finish (-1: none)
  int pay0;  // DATA: result payload of entry 0 (w << 20 | e0).  BOUNDARY: query to set up (-1: none, exit)
  int offs;  // DATA: byte offsets of entry 0 inside the three 16-byte-aligned copies: codes | kappa << 8 | lambda << 16
  float t1, t6, t5;
  int qcount;  // BOUNDARY: ordinal of the query to set up (parity of its table / descriptor buffers)
};

struct __align__(16) PDesc {  // producer-private line descriptor
  int64_t st;
  int len;
  float t1, t6, t5;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ int ld_volatile(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void cons_sync(int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(BAR_CONS), "r"(nthreads) : "memory");
}

// shared-memory carve-up (host and device agree through this struct)
struct Layout {
  int tables, t3stage, sel, lcb, start, cpre, pdesc, meta, bars, ring;
  int slot_bytes, codes_bytes, nslot, total;
};
__host__ __device__ inline Layout make_layout(int M, int W, int smem_limit) {
  Layout L;
  int off = 0;
  L.tables = off;  off += 256 * ROW_WORDS * 4;
  L.t3stage = off; off += 256 * M * 4;
  L.sel = off;     off += (int)select_smem_bytes(SEL_CAP) + 32;  // BlockSelect layout; meta[4..7]: committed, threshold
  off = (off + 15) & ~15;
  L.lcb = off;     off += 256 * 4;
  L.start = off;   off += 2 * W * 8;
  L.cpre = off;    off += ((W + 1) * 4 + 15) & ~15;
  L.pdesc = off;   off += W * (int)sizeof(PDesc);
  L.codes_bytes = CH * M + 16;
  L.slot_bytes = (L.codes_bytes + (CH * 4 + 16) + (CH + 16) + 127) & ~127;
  const int per_slot = L.slot_bytes + (int)sizeof(ChunkMeta) + 16;
  const int fixed = off + 64 /* t3 barrier + counters */ + 128 /* ring alignment */;
  L.nslot = (smem_limit - fixed) / per_slot;
  if (L.nslot > 256) L.nslot = 256;
  if (L.nslot < 0) L.nslot = 0;
  L.meta = off;    off += L.nslot * (int)sizeof(ChunkMeta);
  L.bars = off;    off += L.nslot * 16 + 64;
  off = (off + 127) & ~127;
  L.ring = off;    off += L.nslot * L.slot_bytes;
  L.total = off;
  return L;
}

// k smallest of keys[0..n) by ONE warp: byte-wise radix select (same scheme as WarpSelect::compact, keys streamed from
// shared memory), then an in-place stable partition.  Returns the k-th smallest key; keys[0..k) hold the survivors.
__device__ uint64_t warp_select_smem(uint64_t* keys, int n, int k, int* hist) {
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
    if (base + n <= SEL_CAP) {
      if (take) keys[base + __popc(m & ((1u << lane) - 1))] = make_key(dist, payload);
      __syncwarp();
      if (lane == 0) {
        __threadfence_block();
        atomicAdd(&meta[META_COMMITTED], n);
      }
      return;
    }
    if (base <= SEL_CAP) {
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
      while (ld_volatile(&meta[0]) > SEL_CAP) __nanosleep(128);  // closed for compaction
    }
  }
}

// PROBE: the consumers release every chunk without scoring it (results are meaningless): measures what the producer /
// TMA / HBM side can deliver for this access pattern, the ceiling of the real kernel (tools/bench_scan.py).
template <int MS, int NC, bool PROBE>
__global__ void __launch_bounds__((NC + 1) * 32, 1) scan_stream_kernel(ScanArgs a, int64_t nq, int* work_counter,
                                                                        Layout L) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int NCT = NC * 32;
  const int W = a.W;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned char* tbl = smem + L.tables;
  float* t3stage = reinterpret_cast<float*>(smem + L.t3stage);
  unsigned char* selmem = smem + L.sel;
  uint64_t* skeys = reinterpret_cast<uint64_t*>(selmem);
  int* shist = reinterpret_cast<int*>(skeys + SEL_CAP);
  int* smeta = shist + 256;
  float* lcb = reinterpret_cast<float*>(smem + L.lcb);
  int64_t* start = reinterpret_cast<int64_t*>(smem + L.start);  // [2][W]
  int* cpre = reinterpret_cast<int*>(smem + L.cpre);            // [W + 1] producer
  PDesc* pdesc = reinterpret_cast<PDesc*>(smem + L.pdesc);      // [W] producer
  ChunkMeta* cmeta = reinterpret_cast<ChunkMeta*>(smem + L.meta);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + L.nslot;
  uint64_t* t3bar = empty + L.nslot;
  int* flags = reinterpret_cast<int*>(t3bar + 1);  // [0] tables expanded so far
  unsigned char* ring = smem + L.ring;
  const int NSLOT = L.nslot;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(t3bar, 1);
    flags[0] = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 256; i += blockDim.x) lcb[i] = i < a.nL ? a.lambda_cb[i] : 0.f;
  __syncthreads();

  if (warp == NC) {
    // =========================================================================================== producer
    int slot0 = 0, use0 = 0;  // ring position of the next sequence number (warp-uniform)
    int qcount = 0;
    int prev_q = -1;
    constexpr uint32_t T3_BYTES = 256u * MS * 4u;
    for (;;) {
      int q = 0;
      if (lane == 0) q = atomicAdd(work_counter, 1);
      q = __shfl_sync(kFull, q, 0);
      const bool have = q < nq;
      // ---- boundary markers: one per consumer warp
      if (lane < NC) {
        int slot = slot0 + lane, use = use0;
        if (slot >= NSLOT) {
          slot -= NSLOT;
          use++;
        }
        mbar_wait(&empty[slot], (use & 1) ^ 1);
        ChunkMeta m;
        m.kind = KIND_BOUNDARY;
        m.n = prev_q;
        m.pay0 = have ? q : -1;
        m.offs = 0;
        m.t1 = m.t6 = m.t5 = 0.f;
        m.qcount = qcount;
        cmeta[slot] = m;
        mbar_arrive(&full[slot]);
      }
      __syncwarp();
      slot0 += NC;
      if (slot0 >= NSLOT) {
        slot0 -= NSLOT;
        use0++;
      }
      if (!have) break;
      // ---- the staging buffer and descriptor set of this parity are free once table qcount-1 has been expanded
      while (ld_volatile(&flags[0]) < qcount) __nanosleep(64);
      if (lane == 0) {
        mbar_arrive_expect_tx(t3bar, T3_BYTES);
        bulk_g2s(t3stage, a.t3 + (size_t)q * MS * 256, T3_BYTES, t3bar);
      }
      // ---- line descriptors + chunk prefix
      int64_t* st_q = start + (size_t)(qcount & 1) * W;
      int run = 0;
      for (int w0 = 0; w0 < W; w0 += 32) {
        const int w = w0 + lane;
        int nch = 0;
        if (w < W) {
          const int list = a.line_list[(int64_t)q * W + w];
          PDesc d;
          d.st = 0;
          d.len = 0;
          d.t5 = 0.f;
          if (list >= 0) {
            d.st = a.offsets[list];
            const int64_t l = a.offsets[list + 1] - d.st;
            d.len = (int)(l < a.cap ? l : a.cap);
            d.t5 = a.edge_d2[list];
          }
          d.t1 = a.term1[(int64_t)q * W + w];
          d.t6 = a.term6[(int64_t)q * W + w];
          pdesc[w] = d;
          st_q[w] = d.st;
          nch = (d.len + CH - 1) / CH;
        }
        int inc = nch;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFull, inc, o);
          if (lane >= o) inc += t;
        }
        if (w < W) cpre[w] = run + inc - nch;
        run += __shfl_sync(kFull, inc, 31);
      }
      if (lane == 0) cpre[W] = run;
      __syncwarp();
      const int total = run;
      // ---- chunk issue, 32 chunks per iteration (one per lane)
      for (int c0 = 0; c0 < total; c0 += 32) {
        const int c = c0 + lane;
        if (c < total) {
          int lo = 0, hi = W;  // largest w with cpre[w] <= c
          while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cpre[mid] <= c) lo = mid; else hi = mid;
          }
          const PDesc d = pdesc[lo];
          const int e0 = (c - cpre[lo]) * CH;
          const int n = min(CH, d.len - e0);
          const int64_t first = d.st + e0;
          int slot = slot0 + lane, use = use0;
          if (slot >= NSLOT) {
            slot -= NSLOT;
            use++;
          }
          // 16-byte aligned supersets of the three byte ranges of the chunk
          const int64_t cb0 = first * MS, kb0 = first * 4, lb0 = first;
          const int64_t ca = cb0 & ~int64_t(15), ka = kb0 & ~int64_t(15), la = lb0 & ~int64_t(15);
          const uint32_t cbytes = (uint32_t)(((cb0 + (int64_t)n * MS + 15) & ~int64_t(15)) - ca);
          const uint32_t kbytes = (uint32_t)(((kb0 + (int64_t)n * 4 + 15) & ~int64_t(15)) - ka);
          const uint32_t lbytes = (uint32_t)(((lb0 + n + 15) & ~int64_t(15)) - la);
          mbar_wait(&empty[slot], (use & 1) ^ 1);
          ChunkMeta m;
          m.kind = KIND_DATA;
          m.n = n;
          m.pay0 = (lo << 20) | e0;
          m.offs = (int)(cb0 - ca) | ((int)(kb0 - ka) << 8) | ((int)(lb0 - la) << 16);
          m.t1 = d.t1;
          m.t6 = d.t6;
          m.t5 = d.t5;
          m.qcount = qcount;
          cmeta[slot] = m;
          unsigned char* dst = ring + (size_t)slot * L.slot_bytes;
          mbar_arrive_expect_tx(&full[slot], cbytes + kbytes + lbytes);
          bulk_g2s(dst, a.codes + ca, cbytes, &full[slot]);
          bulk_g2s(dst + L.codes_bytes, reinterpret_cast<const unsigned char*>(a.kappa) + ka, kbytes, &full[slot]);
          bulk_g2s(dst + L.codes_bytes + CH * 4 + 16, a.lamq + la, lbytes, &full[slot]);
        }
        __syncwarp();
        const int cnt = min(32, total - c0);
        slot0 += cnt;
        if (slot0 >= NSLOT) {
          slot0 -= NSLOT;
          use0++;
        }
      }
      prev_q = q;
      qcount++;
    }
    return;
  }

  // ============================================================================================= consumers
  int slot = warp, use = 0;
  const uint32_t lofs = 4u * (uint32_t)lane;
  int cur_q = -1, cur_par = 0;
  for (;;) {
    mbar_wait(&full[slot], use & 1);
    const int4 m0 = reinterpret_cast<const int4*>(cmeta + slot)[0];
    const int4 m1 = reinterpret_cast<const int4*>(cmeta + slot)[1];
    if (PROBE && m0.x == KIND_DATA) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
    } else if (m0.x == KIND_DATA) {
      const unsigned char* sp = ring + (size_t)slot * L.slot_bytes;
      const int n = m0.y;
      const unsigned char* cp = sp + (m0.w & 0xff);
      const float* kp = reinterpret_cast<const float*>(sp + L.codes_bytes + ((m0.w >> 8) & 0xff));
      const unsigned char* lp = sp + L.codes_bytes + CH * 4 + 16 + ((m0.w >> 16) & 0xff);
      const float t1 = __int_as_float(m1.x), t6 = __int_as_float(m1.y), t5 = __int_as_float(m1.z);
      uint32_t cw[2][MS / 4];
      float kap[2];
      uint32_t lq[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const int i = u * 32 + lane;
        if constexpr (MS == 16) {
          const uint4 c = *reinterpret_cast<const uint4*>(cp + i * 16);
          cw[u][0] = c.x; cw[u][1] = c.y; cw[u][MS / 4 - 2] = c.z; cw[u][MS / 4 - 1] = c.w;
        } else {
          const uint2 c = *reinterpret_cast<const uint2*>(cp + i * 8);
          cw[u][0] = c.x; cw[u][MS / 4 - 1] = c.y;
        }
        kap[u] = kp[i];
        lq[u] = lp[i];
      }
      float dist[2];
#pragma unroll
      for (int u = 0; u < 2; u++) {
        const float la = lcb[lq[u]];
        const float base_d = t1 + la * t6 + (la * la - la) * t5;
        float acc = 0.f;
#pragma unroll
        for (int s = 0; s < MS; s++) {
          const uint32_t o = __byte_perm(cw[u][s >> 2], lofs, 0x5504 | ((s & 3) << 4));  // code << 8 | 4 * lane
          acc += *reinterpret_cast<const float*>(tbl + o + 4 * s);
        }
        dist[u] = (kap[u] + acc) + base_d;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);  // every lane has its entries in registers
      const float thr = __int_as_float(ld_volatile(&smeta[META_THR]));
      const bool v0 = lane < n, v1 = lane + 32 < n;
      const bool pass = (v0 && dist[0] <= thr) || (v1 && dist[1] <= thr);
      if (__any_sync(kFull, pass)) {  // rare once the threshold has tightened
        shared_offer(skeys, shist, smeta, a.k, v0, dist[0], (uint32_t)(m0.z + lane), lane);
        shared_offer(skeys, shist, smeta, a.k, v1, dist[1], (uint32_t)(m0.z + 32 + lane), lane);
      }
    } else {
      // ---- query boundary: finish m0.y, set up m0.z
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);
      const int fin_q = m0.y, next_q = m0.z, qcount = m1.w;
      if (fin_q >= 0) {
        cons_sync(NCT);  // every consumer has offered its last candidate of the query
        BlockSelect<NCT, BAR_CONS> sel;
        sel.attach(selmem, a.k, SEL_CAP);
        sel.finish();
        const int64_t* st_q = start + (size_t)cur_par * W;
        for (int i = threadIdx.x; i < a.k; i += NCT) {
          const uint64_t key = sel.keys[i];
          float dv = FLT_MAX;
          int64_t id = -1;
          if (key != kKeyInf) {
            const uint32_t pay = key_payload(key);
            dv = key_val(key);
            id = a.ids[st_q[pay >> 20] + (pay & 0xfffffu)];
          }
          a.outD[(int64_t)fin_q * a.k + i] = dv;
          a.outI[(int64_t)fin_q * a.k + i] = id;
        }
      }
      if (next_q < 0) break;
      cur_q = next_q;
      cur_par = qcount & 1;
      mbar_wait(t3bar, qcount & 1);
      if (fin_q >= 0) cons_sync(NCT);  // the result loop above has read the selection buffer
      {  // word c of row `code` = T3[c mod M][code] (staged code-major): whole float4s
        const float4* src4 = reinterpret_cast<const float4*>(t3stage);
        float4* dst4 = reinterpret_cast<float4*>(tbl);
        constexpr int Q4 = (31 + MS + 3) / 4, S4 = MS / 4;
        for (int i = threadIdx.x; i < 256 * Q4; i += NCT) {
          const int code = i / Q4, q4 = i % Q4;
          dst4[code * (ROW_WORDS / 4) + q4] = src4[code * S4 + (q4 & (S4 - 1))];
        }
      }
      if (threadIdx.x == 0) {
        smeta[0] = 0;
        smeta[META_COMMITTED] = 0;
        smeta[META_THR] = 0x7f800000;
      }
      cons_sync(NCT);
      if (threadIdx.x == 0) {
        __threadfence_block();
        *reinterpret_cast<volatile int*>(&flags[0]) = qcount + 1;  // stage + descriptor buffers of this parity are free
      }
    }
    slot += NC;
    if (slot >= NSLOT) {
      slot -= NSLOT;
      use++;
    }
  }
  (void)cur_q;
}

static int sm_count() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return sms;
}
static int smem_optin() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return v;
}

constexpr int NC_DEFAULT = 19;  // consumer warps (+ 1 producer warp = 640 threads)

template <int MS, int NC>
static int launch_t(const ScanArgs& a, int64_t nq, int* counter, cudaStream_t st) {
  const Layout L = make_layout(MS, a.W, smem_optin());
  if (L.nslot < 2 * NC + 32) return VLQ_EUNSUPPORTED;
  const unsigned grid = (unsigned)(nq < sm_count() ? nq : sm_count());
  static const bool probe = getenv("VLQ_SCAN_PROBE") != nullptr;  // measurement aid, see the kernel
  if (probe) {
    VLQ_CUDA_TRY(cudaFuncSetAttribute(scan_stream_kernel<MS, NC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    VLQ_LAUNCH((scan_stream_kernel<MS, NC, true>), grid, (NC + 1) * 32, (size_t)L.total, st, a, nq, counter, L);
    return last_error();
  }
  VLQ_CUDA_TRY(cudaFuncSetAttribute(scan_stream_kernel<MS, NC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  VLQ_LAUNCH((scan_stream_kernel<MS, NC, false>), grid, (NC + 1) * 32, (size_t)L.total, st, a, nq, counter, L);
  return last_error();
}

}  // namespace stream

bool scan_stream_supported(const ScanArgs& a) {
  if (!(a.M == 16 || a.M == 8) || a.ksub != 256 || !a.t3 || a.nL > 256 || a.k > VLQ_MAX_K || a.cap > (1 << 20)) return false;
  const uintptr_t al = reinterpret_cast<uintptr_t>(a.codes) | reinterpret_cast<uintptr_t>(a.kappa) |
                       reinterpret_cast<uintptr_t>(a.lamq) | reinterpret_cast<uintptr_t>(a.t3);
  if (al & 15) return false;
  const stream::Layout L = stream::make_layout(a.M, a.W, stream::smem_optin());
  return L.nslot >= 2 * stream::NC_DEFAULT + 32;
}

int launch_scan_stream(const ScanArgs& a, int64_t nq, int* counter, cudaStream_t st) {
  if (!scan_stream_supported(a)) return VLQ_EUNSUPPORTED;
  VLQ_CUDA_TRY(cudaMemsetAsync(counter, 0, sizeof(int), st));
  if (a.M == 16) return stream::launch_t<16, stream::NC_DEFAULT>(a, nq, counter, st);
  return stream::launch_t<8, stream::NC_DEFAULT>(a, nq, counter, st);
}

}  // namespace vlq
