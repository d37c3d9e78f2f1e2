// slk_kernels.cuh -- the two W-templated kernels (classify, build-side emit) and their launch descriptors.
// Each window width W = k-m+1 is instantiated in its own translation unit (slk_inst.cu with -DSLK_W=n) so the
// library builds in parallel; slacken_gpu.cu dispatches on the index's W at run time.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "../../include/slacken_gpu.h"
#include "slk_core.h"
#include "slk_group.h"

#define BUILD_WPT 96  // k-mer windows scanned per thread of the build-side emit kernel

// ASCII input: bases*/off* (+shift: offsets are absolute in the caller's buffer, the device buffer starts at `shift`).
// Packed input: bases* point at the uint64 code blocks, mask* at the uint32 masks, off* are block offsets, len* the
// read lengths in bases.
struct slk_classify_args {
  slk_scan_params sp; slk_table_view tb; slk_tax_view tx;
  const uint8_t* bases1; const uint64_t* off1; uint64_t shift1;
  const uint8_t* bases2; const uint64_t* off2; uint64_t shift2;
  const uint32_t* mask1; const uint32_t* len1; const uint32_t* mask2; const uint32_t* len2; bool packed;
  uint32_t n_reads; double confidence; int32_t min_hit_groups;
  int32_t* taxon_out; uint8_t* flags_out; slk_read_detail* detail_out;
  slk_hit* hits_base; const unsigned long long* hits_shift_ptr; uint64_t hits_cap; unsigned long long* hits_cursor;
  unsigned long long* counts; uint32_t* error_flag; unsigned long long* stats;
  bool hits; cudaStream_t stream;
};
struct slk_emit_args {
  slk_scan_params sp; const uint8_t* bases; const uint64_t* frag_off; uint64_t off_shift; const uint32_t* frag_dense;
  const uint64_t* item_prefix; uint32_t n_frag; uint64_t n_items; uint64_t* out; uint64_t cap; unsigned long long* cursor;
  cudaStream_t stream;
};
// split path, scan step: pass 1 counts the spans of every fragment, pass 2 (after an exclusive scan of the counts)
// writes the span words to span_off[r] ...
struct slk_spans_args {
  slk_scan_params sp;
  const uint8_t* bases1; const uint64_t* off1; const uint8_t* bases2; const uint64_t* off2;   // ASCII, device
  uint32_t n_reads;
  uint64_t* span_off;      // [n_reads + 1]: pass 1 writes counts to [0, n), pass 2 reads offsets
  uint64_t* spans;         // pass 2 output (nullptr = pass 1)
  cudaStream_t stream;
  // one-pass variant: counts to span_off AND span words to scratch[r * stride ...] (stride > 0); overflow = flag word
  uint64_t* scratch = nullptr; uint32_t stride = 0; uint32_t* overflow = nullptr;
};
// Bracken weights, scan step: pass 1 counts the hits of every genome fragment, pass 2 writes them at hit_off[f] ...
struct slk_bracken_scan_args {
  slk_scan_params sp;
  const uint8_t* bases; const uint64_t* frag_off; uint32_t n_frag;
  uint64_t* hit_off;       // [n_frag + 1]
  slk_bhit* hits;          // nullptr = pass 1
  cudaStream_t stream;
};
#define SLK_DECL_W(w) void slk_launch_classify_w##w(const slk_classify_args&); void slk_launch_emit_w##w(const slk_emit_args&); \
  void slk_launch_classify2_w##w(const slk_classify2_args&, cudaStream_t); \
  void slk_launch_spans_w##w(const slk_spans_args&); void slk_launch_bracken_scan_w##w(const slk_bracken_scan_args&);
SLK_DECL_W(1) SLK_DECL_W(2) SLK_DECL_W(3) SLK_DECL_W(4) SLK_DECL_W(5) SLK_DECL_W(6) SLK_DECL_W(7) SLK_DECL_W(8)

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t warp_agg_alloc(unsigned long long* cursor, uint32_t n) {
  // all 32 lanes must call; returns each lane's offset in a warp-wide contiguous allocation
  uint32_t lane = threadIdx.x & 31, incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += v;
  }
  uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 0 && total) base = atomicAdd(cursor, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + incl - n;
}

// ---------------------------------------------------------------------------------------------- build kernels
// K2 (build side) + K3a emit: one thread scans BUILD_WPT k-mer windows of one fragment and appends its cells.
template <int W>
__global__ void __launch_bounds__(128) emit_cells_kernel(const __grid_constant__ slk_scan_params sp, const uint8_t* __restrict__ bases,
                                                         const uint64_t* __restrict__ frag_off, uint64_t off_shift,
                                                         const uint32_t* __restrict__ frag_dense,
                                                         const uint64_t* __restrict__ item_prefix, uint32_t n_frag,
                                                         uint64_t n_items, uint64_t* __restrict__ out,
                                                         uint64_t cap, unsigned long long* cursor) {
  uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t local[BUILD_WPT];
  uint32_t nl = 0;
  if (it < n_items) {
    // fragment of this item: last f with item_prefix[f] <= it
    uint32_t lo = 0, hi = n_frag;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (item_prefix[mid] <= it) lo = mid; else hi = mid;
    }
    uint32_t f = lo;
    uint32_t dense = frag_dense[f];
    if (dense != 0) {
      uint64_t fs = frag_off[f] - off_shift, fe = frag_off[f + 1] - off_shift;
      uint64_t w0 = (it - item_prefix[f]) * BUILD_WPT;
      uint64_t nb = fe - fs - w0;
      uint64_t want = (uint64_t)BUILD_WPT + sp.k - 1;
      if (nb > want) nb = want;
      auto emit = [&](uint64_t cell) { local[nl++] = cell; };
      slk_emit_cells<W>(sp, bases + fs + w0, nb, dense, emit);
    }
  }
  uint64_t o = warp_agg_alloc(cursor, nl);
  for (uint32_t j = 0; j < nl; j++)
    if (o + j < cap) out[o + j] = local[j];
}

// ---------------------------------------------------------------------------------------------- split path: scan
template <int W, bool EMIT>
__global__ void __launch_bounds__(128) spans_kernel(const __grid_constant__ slk_scan_params sp, const uint8_t* __restrict__ bases1,
                                                    const uint64_t* __restrict__ off1, const uint8_t* __restrict__ bases2,
                                                    const uint64_t* __restrict__ off2, uint32_t n_reads,
                                                    uint64_t* __restrict__ span_off, uint64_t* __restrict__ spans) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t s1 = off1[r], e1 = off1[r + 1];
  const uint8_t* p2 = nullptr;
  uint32_t l2 = 0;
  if (bases2) { const uint64_t s2 = off2[r], e2 = off2[r + 1]; p2 = bases2 + s2; l2 = (uint32_t)(e2 - s2); }
  if (EMIT) {
    uint64_t* out = spans + span_off[r];
    slk_scan_fragment_spans<W>(sp, bases1 + s1, (uint32_t)(e1 - s1), p2, l2, bases2 != nullptr, [&](uint64_t w) { *out++ = w; });
  } else {
    uint64_t n = 0;
    slk_scan_fragment_spans<W>(sp, bases1 + s1, (uint32_t)(e1 - s1), p2, l2, bases2 != nullptr, [&](uint64_t) { n++; });
    span_off[r] = n;
  }
}

// One pass instead of two: fragment r writes its span words to its own row of `stride` slots (stride = the most spans
// a fragment of the batch's longest reads can have: one per k-mer window, plus the mate border) and its count to
// span_off[r]; after the prefix sum compact_spans_kernel moves the rows together, one warp per fragment.
template <int W>
__global__ void __launch_bounds__(128) spans_strided_kernel(const __grid_constant__ slk_scan_params sp, const uint8_t* __restrict__ bases1,
                                                            const uint64_t* __restrict__ off1, const uint8_t* __restrict__ bases2,
                                                            const uint64_t* __restrict__ off2, uint32_t n_reads, uint32_t stride,
                                                            uint64_t* __restrict__ span_off, uint64_t* __restrict__ scratch,
                                                            uint32_t* overflow) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_reads) return;
  const uint64_t s1 = off1[r], e1 = off1[r + 1];
  const uint8_t* p2 = nullptr;
  uint32_t l2 = 0;
  if (bases2) { const uint64_t s2 = off2[r], e2 = off2[r + 1]; p2 = bases2 + s2; l2 = (uint32_t)(e2 - s2); }
  uint64_t* out = scratch + (uint64_t)r * stride;
  uint32_t n = 0;
  slk_scan_fragment_spans<W>(sp, bases1 + s1, (uint32_t)(e1 - s1), p2, l2, bases2 != nullptr, [&](uint64_t w) {
    if (n < stride) out[n] = w;
    n++;
  });
  if (n > stride) atomicExch(overflow, 1u);
  span_off[r] = n;
}

// ---------------------------------------------------------------------------------------------- Bracken weights: scan
template <int W, bool EMIT>
__global__ void __launch_bounds__(32) bracken_scan_kernel(const __grid_constant__ slk_scan_params sp, const uint8_t* __restrict__ bases,
                                                          const uint64_t* __restrict__ frag_off, uint32_t n_frag,
                                                          uint64_t* __restrict__ hit_off, slk_bhit* __restrict__ hits) {
  const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_frag) return;
  const uint64_t s = frag_off[f], e = frag_off[f + 1];
  if (EMIT) {
    slk_bhit* out = hits + hit_off[f];
    slk_bracken_scan<W>(sp, bases + s, (uint32_t)(e - s), [&](const slk_bhit& h) { *out++ = h; });
  } else {
    uint64_t n = 0;
    slk_bracken_scan<W>(sp, bases + s, (uint32_t)(e - s), [&](const slk_bhit&) { n++; });
    hit_off[f] = n;
  }
}

// ---------------------------------------------------------------------------------------------- classify kernel
// Fast store of a warp in shared memory (see slk_store_local for the contract):
//   stage A [POOL][16 B] | stage B [POOL][16 B] | tile 0 | tile 1,  tile = keys [POOL][8 B] | metas [POOL][2 B] |
//   nexts [POOL][1 B] | pending [POOL][1 B]
// so that 32 lanes working on 32 consecutive slots touch consecutive words (no bank conflicts), and the two 16-byte
// halves of a bucket are the two cp.async (LDGSTS through L1: one sector request per bucket) destinations.
// It holds 32-bit shared-space addresses and uses ld.shared / st.shared explicitly: with generic pointers the
// compiler must assume that every store may alias the pointers themselves and reloads them around each access.
// commit/wait_group give every thread its own FIFO of outstanding copies.
#define SLK_CLS_THREADS 128
#define SLK_TILE_BYTES (12u * SLK_POOL)
#define SLK_WARP_BYTES (32u * SLK_POOL + 2u * SLK_TILE_BYTES)
#define SLK_SMEM_HITS ((SLK_CLS_THREADS / 32u) * SLK_WARP_BYTES)
#define SLK_SMEM_BYTES (SLK_SMEM_HITS + SLK_SHITS * 8u * SLK_CLS_THREADS)
static_assert(SLK_POOL % 32 == 0 && SLK_POOL >= 128 && SLK_POOL <= 256, "SLK_POOL: whole rounds of 32 slots, 8-bit links");
extern __shared__ __align__(16) uint8_t slk_smem[];
struct dev_store {
  uint32_t warp, hits, lane8;
  __device__ __forceinline__ static uint32_t sa(uint32_t off) { return (uint32_t)__cvta_generic_to_shared(slk_smem) + off; }
  __device__ __forceinline__ uint32_t tile(uint32_t t) const { return warp + 32u * SLK_POOL + t * SLK_TILE_BYTES; }
  __device__ __forceinline__ void put(uint32_t b, uint32_t s, uint64_t k, uint32_t m) const {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(sa(b + s * 8u)), "l"(k) : "memory");
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(sa(b + 8u * SLK_POOL + s * 2u)), "h"((uint16_t)m) : "memory");
  }
  __device__ __forceinline__ void set_key(uint32_t t, uint32_t s, uint64_t k) const {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(sa(t + s * 8u)), "l"(k) : "memory");
  }
  __device__ __forceinline__ uint64_t key(uint32_t t, uint32_t s) const {
    uint64_t k;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(k) : "r"(sa(t + s * 8u)) : "memory");
    return k;
  }
  __device__ __forceinline__ uint32_t meta(uint32_t t, uint32_t s) const {
    uint16_t m;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(m) : "r"(sa(t + 8u * SLK_POOL + s * 2u)) : "memory");
    return m;
  }
  __device__ __forceinline__ void set_next(uint32_t t, uint32_t s, uint32_t n) const {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa(t + 10u * SLK_POOL + s)), "r"(n) : "memory");
  }
  __device__ __forceinline__ uint32_t next(uint32_t t, uint32_t s) const {
    uint32_t n;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(n) : "r"(sa(t + 10u * SLK_POOL + s)) : "memory");
    return n;
  }
  __device__ __forceinline__ void set_pending(uint32_t t, uint32_t q, uint32_t s) const {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa(t + 11u * SLK_POOL + q)), "r"(s) : "memory");
  }
  __device__ __forceinline__ uint32_t pending(uint32_t t, uint32_t q) const {
    uint32_t s;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(s) : "r"(sa(t + 11u * SLK_POOL + q)) : "memory");
    return s;
  }
  __device__ __forceinline__ void fetch(uint32_t s, const uint64_t* src) const {
    const uint32_t d = sa(warp + s * 16u);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d + 16u * SLK_POOL), "l"(src + 2) : "memory");
  }
  __device__ __forceinline__ void bucket(uint32_t s, slk_bucket* o) const {
    const uint32_t d = sa(warp + s * 16u);
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(o->c0), "=l"(o->c1) : "r"(d) : "memory");
    asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(o->c2), "=l"(o->c3) : "r"(d + 16u * SLK_POOL) : "memory");
  }
  __device__ __forceinline__ void commit(uint32_t) const { asm volatile("cp.async.commit_group;" ::: "memory"); }
  __device__ __forceinline__ void wait_all(uint32_t) const { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

  __device__ __forceinline__ void set_hit(uint32_t i, int32_t label, int32_t count) const {
    asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(sa(hits + i * (8u * SLK_CLS_THREADS))), "r"(label), "r"(count) : "memory");
  }
  __device__ __forceinline__ void get_hit(uint32_t i, int32_t* label, int32_t* count) const {
    int32_t l, c;
    asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(l), "=r"(c) : "r"(sa(hits + i * (8u * SLK_CLS_THREADS))) : "memory");
    *label = l; *count = c;
  }
  __device__ __forceinline__ void hist_set(uint32_t i, uint32_t t, int32_t v) const {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa(warp + i * 256u + lane8)), "r"(t), "r"((uint32_t)v) : "memory");
  }
  __device__ __forceinline__ void hist_get(uint32_t i, uint32_t* t, int32_t* v) const {
    uint32_t a, b;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(sa(warp + i * 256u + lane8)) : "memory");
    *t = a; *v = (int32_t)b;
  }
};

// Spill path for a read with more merged hits than the per-thread buffers hold: it takes a worst-case block of the
// global hit buffer for itself (one atomic) and appends there. Ordinary reads never touch it; their hits are
// written from the buffers to a compact, warp-allocated block at the end of the kernel.
struct dev_hit_sink {
  uint32_t n;
  bool spilled, live;     // live: the lane holds a fragment of the batch (lanes past its end scan an empty one)
  uint64_t goff;
  uint64_t reserved;      // first slot of this lane's buffered hits in the compact, warp-allocated block
  slk_hit* gbase;         // scratch, indexed by (absolute index - gshift)
  uint64_t gshift, gcap;  // gcap: capacity of gbase in hits
  unsigned long long* cursor;
  __device__ __forceinline__ void put(uint64_t abs_idx, int32_t taxon, int32_t count) {
    uint64_t rel = abs_idx - gshift;
    slk_hit h;
    h.taxon = taxon; h.count = count;
    if (rel < gcap) gbase[rel] = h;
  }
  __device__ __forceinline__ void reserve(uint32_t n_hits);   // all 32 lanes call it
  __device__ __forceinline__ void push(int32_t taxon, int32_t count, uint32_t need) {
    if (!spilled) {
      goff = atomicAdd(cursor, (unsigned long long)need + 2ull);
      spilled = true;
    }
    put(goff + n, taxon, count);
    n++;
  }
};

__device__ __forceinline__ void dev_hit_sink::reserve(uint32_t n_hits) { reserved = warp_agg_alloc(cursor, live ? n_hits : 0u); }

#ifndef SLK_CLS_MINB
#define SLK_CLS_MINB 4   // 4 x 51 KB of tiles per SM; up to 128 registers per thread, so nothing spills
#endif
template <int W, bool HITS, bool PACKED>
__global__ void __launch_bounds__(SLK_CLS_THREADS, SLK_CLS_MINB) classify_kernel(const __grid_constant__ slk_scan_params sp,
                                                       const __grid_constant__ slk_table_view tb,
                                                       const __grid_constant__ slk_tax_view tx,
                                                       const uint8_t* __restrict__ bases1, const uint64_t* __restrict__ off1,
                                                       uint64_t shift1, const uint8_t* __restrict__ bases2,
                                                       const uint64_t* __restrict__ off2, uint64_t shift2,
                                                       const uint32_t* __restrict__ mask1, const uint32_t* __restrict__ len1,
                                                       const uint32_t* __restrict__ mask2, const uint32_t* __restrict__ len2,
                                                       uint32_t n_reads, double confidence, int32_t min_hit_groups,
                                                       int32_t* __restrict__ taxon_out, uint8_t* __restrict__ flags_out,
                                                       slk_read_detail* __restrict__ detail_out, slk_hit* hits_base,
                                                       const unsigned long long* hits_shift_ptr, uint64_t hits_cap,
                                                       unsigned long long* hits_cursor, unsigned long long* counts,
                                                       uint32_t* error_flag, unsigned long long* stats) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  bool active = r < n_reads;
  slk_frag_result res;
  res.taxon = 0; res.flags = 0; res.kmers1 = 0; res.kmers2 = 0; res.num_distinct = 0; res.n_hits = 0; res.n_probes = 0;
  typedef typename std::conditional<HITS, dev_hit_sink, slk_null_sink>::type sink_t;
  dev_store ent;
  ent.warp = (threadIdx.x >> 5) * SLK_WARP_BYTES;
  ent.hits = SLK_SMEM_HITS + threadIdx.x * 8u;
  ent.lane8 = (threadIdx.x & 31u) * 8u;
  sink_t sink;
  if constexpr (HITS) {
    sink.n = 0; sink.spilled = false; sink.live = active; sink.goff = 0; sink.reserved = 0; sink.gbase = hits_base;
    sink.gshift = hits_shift_ptr ? *hits_shift_ptr : 0ull;
    sink.gcap = hits_cap; sink.cursor = hits_cursor;
  }
  slk_frag_classifier<W, sink_t, dev_store> cl(tb, tx, sink, ent);
  {
    // every lane runs the classifier (lanes past the end of the batch get an empty read) so that the warp-wide
    // votes inside run() always see 32 participants
    slk_read_src r1, r2;
    r1.ascii = bases1; r1.codes = reinterpret_cast<const uint64_t*>(bases1); r1.mask = mask1; r1.len = 0;
    r2.ascii = bases2; r2.codes = reinterpret_cast<const uint64_t*>(bases2); r2.mask = mask2; r2.len = 0;
    if (active) {
      if (PACKED) {
        const uint64_t b1 = off1[r] - shift1;
        r1.codes += b1; r1.mask += b1; r1.len = len1[r];
        if (bases2) {
          const uint64_t b2 = off2[r] - shift2;
          r2.codes += b2; r2.mask += b2; r2.len = len2[r];
        }
      } else {
        const uint64_t s1 = off1[r], e1 = off1[r + 1];
        r1.ascii = bases1 + (s1 - shift1); r1.len = (uint32_t)(e1 - s1);
        if (bases2) {
          const uint64_t s2 = off2[r], e2 = off2[r + 1];
          r2.ascii = bases2 + (s2 - shift2); r2.len = (uint32_t)(e2 - s2);
        }
      }
    }
    if (sp.canonical) cl.template run<PACKED, true>(sp, r1, r2, bases2 != nullptr, confidence, min_hit_groups, res);
    else cl.template run<PACKED, false>(sp, r1, r2, bases2 != nullptr, confidence, min_hit_groups, res);
    if (active) {
      taxon_out[r] = res.taxon;
      flags_out[r] = (uint8_t)(res.flags & 3u);
      if (res.flags & SLK_F_OVERFLOW) atomicExch(error_flag, 1u);
    }
  }
  // The merged hits and their (warp-allocated, compact) place in the output. The allocation's atomic was issued inside
  // run(), before the resolve step; everything that does not need its result is done first.
  int32_t hl[SLK_SHITS], hc[SLK_SHITS];
  uint32_t n_buf = 0;
  if constexpr (HITS) {
    if (active && !sink.spilled) {
      n_buf = cl.nh;
#pragma unroll
      for (uint32_t i = 0; i < SLK_SHITS; i++)
        if (i < n_buf) ent.get_hit(i, &hl[i], &hc[i]);
#pragma unroll
      for (uint32_t i = 0; i < SLK_SHITS; i++)   // dense labels become raw taxon ids here: all lookups in flight together
        if (i < n_buf && hl[i] >= 0) hl[i] = __ldg(tx.raw + hl[i]);
    }
  }
  if (stats != nullptr) {  // probes and merged hits of this launch (the S and H of the roofline arithmetic)
    uint32_t np = __reduce_add_sync(0xffffffffu, active ? res.n_probes : 0u);
    uint32_t nh = __reduce_add_sync(0xffffffffu, active ? res.n_hits : 0u);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&stats[0], (unsigned long long)np);
      atomicAdd(&stats[1], (unsigned long long)nh);
    }
  }
  // K7: per-taxon report counters (groupBy(sampleId, taxon).count, slacken/Classifier.scala:214-217),
  // aggregated per warp with match_any so a hot taxon costs one atomic per warp
  if (counts != nullptr) {
    bool cnt = active && (res.flags & SLK_F_HAS_SPAN);
    uint32_t key = cnt ? (uint32_t)res.taxon : 0xFFFFFFFFu;
    uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (cnt && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1))
      atomicAdd(&counts[(uint32_t)res.taxon], (unsigned long long)__popc(peers));
  }
  if constexpr (HITS) {
    if (active) {
      const uint64_t o = sink.reserved;
#pragma unroll
      for (uint32_t i = 0; i < SLK_SHITS; i++)
        if (i < n_buf) sink.put(o + i, hl[i], hc[i]);
      for (uint32_t i = SLK_SHITS; i < n_buf; i++) {
        int32_t l, c;
        cl.buffered_hit(i, &l, &c);
        sink.put(o + i, l >= 0 ? tx.raw[l] : l, c);
      }
      slk_read_detail d;
      d.hit_off = sink.spilled ? sink.goff : o;
      d.hit_cnt = res.n_hits;
      d.len1 = res.kmers1 + (uint32_t)(sp.k - 1);
      d.len2 = bases2 ? res.kmers2 + (uint32_t)(sp.k - 1) : 0xFFFFFFFFu;
      d.num_distinct = res.num_distinct;
      detail_out[r] = d;
    }
  } else if (detail_out != nullptr && active) {
    slk_read_detail d;
    d.hit_off = 0; d.hit_cnt = 0;
    d.len1 = res.kmers1 + (uint32_t)(sp.k - 1);
    d.len2 = bases2 ? res.kmers2 + (uint32_t)(sp.k - 1) : 0xFFFFFFFFu;
    d.num_distinct = res.num_distinct;
    detail_out[r] = d;
  }
}

// ---------------------------------------------------------------------------------------------- classify kernel, 2nd generation
// One warp = 32 fragments (slk_group.h). Four warps per block, four blocks per SM: 16 x 13.6 KB of entry buffers
// (measured: 14 warps 712, 16 warps 741-750 M reads/s; more warps need fewer than 100 registers, and spilling costs more).
#ifndef SLK_G_THREADS
#define SLK_G_THREADS 128
#endif
#ifndef SLK_G_MINB
#define SLK_G_MINB 4
#endif
#define SLK_G_SMEM_BYTES ((SLK_G_THREADS / 32u) * SLK_G_WARP_BYTES)
template <int W, bool CANON>
__global__ void __launch_bounds__(SLK_G_THREADS, SLK_G_MINB) classify2_kernel(const __grid_constant__ slk_classify2_args a) {
  extern __shared__ __align__(16) uint8_t slk_smem2[];
  slk_classify2_thread<W, CANON>(a, slk_smem2);
}

#endif  // __CUDACC__
