// slk_group.h -- the fused classify kernel body, second generation: one WARP classifies 32 fragments (lane = fragment).
//
// What changed against the first-generation body (slk_frag_classifier in slk_core.h), and why (profiles/r01_*.md):
//  * SCAN. The first kernel rolled the m-mer one base per step and paid a ballot, a popcount and a conditional
//    shared-memory append for every base of every lane (77 warp instructions per base step). Here a lane takes a CHUNK of 16
//    k-mer windows of its read at a time and computes all of them position-parallel: every m-mer is cut straight out of
//    the 2-bit stream with funnel shifts whose amounts are compile-time constants (the reverse strand out of the
//    pair-reversed stream), the window minimum is a three-level tree shared between neighbouring windows, and there is
//    no loop-carried dependency except a two-instruction run counter: ~40 instructions per window, all independent.
//    Super-mer ends are collected in a bit mask; one warp prefix sum per chunk gives every lane the place of its entries
//    in the warp's entry buffer, and a per-(step, lane) table remembers where they are.
//  * LOOKUPS. Plain 32-byte loads (LDG.256) into registers, four in flight per lane, issued in flat order over the whole
//    buffer so that every lane is busy whatever its own read looks like. A stream of cp.async gathers that miss to DRAM
//    blocks the shared-memory path of the whole SM (profiles/r01_probe_microbench.md, section 6); plain loads do not, and
//    the warps that wait for them leave the issue slots to the warps that scan.
//  * MERGE / RESOLVE as before: every lane walks its own entries in span order (numDistinct, k-mer totals,
//    TaxonCounts.fromHits), resolveTree runs on a histogram in the (now idle) entry buffer.
//
// Reference semantics: see the header of slk_core.h; the statements below cite the same files.
// This file is compiled by nvcc for the device and by g++ for tests/host_emulation (with simt.h, which runs the 32 lanes
// of a warp as fibres); it uses CUDA's warp intrinsics directly.
#pragma once
#include "slk_core.h"

#ifndef SLK_G_CAP
#define SLK_G_CAP 1152        // entries in a warp's buffer (multiple of 32)
#endif
#define SLK_G_CHUNK 16        // k-mer windows per chunk
#ifndef SLK_G_STEPS
#define SLK_G_STEPS 16        // chunk steps between two closes of the buffer (size of the step table)
#endif
#define SLK_G_MAXNE 19        // entries one lane can produce in one step: border + cut run + 16 run ends + the run open at the mate's end
#ifndef SLK_G_DEPTH
#define SLK_G_DEPTH 4         // table lookups in flight per lane (measured: 2 -> 762, 3 -> 737, 4 -> 802, 6 -> 624 M reads/s)
#endif
#ifndef SLK_G_HIST
#define SLK_G_HIST 32         // (taxon, k-mers) pairs per lane that fit the idle entry buffer
#endif
#define SLK_G_CNT_MAX 63u     // k-mer windows per entry (6 bits); longer runs are split, which no output can see

// per-warp shared memory: keys u64[CAP] | table u16[STEPS][32] | metas u8[CAP] | hits (label, count)[SLK_SHITS][32] | pending u16[PEND]
#define SLK_G_OFF_TABLE (8u * SLK_G_CAP)
#define SLK_G_OFF_META (SLK_G_OFF_TABLE + 2u * 32u * SLK_G_STEPS)
#define SLK_G_OFF_HITS (SLK_G_OFF_META + SLK_G_CAP)
#define SLK_G_OFF_PEND (SLK_G_OFF_HITS + 8u * 32u * SLK_SHITS)
#define SLK_G_PEND 64u        // entries whose first bucket was full without a match, worked off 32 and more at a time
#define SLK_G_WARP_BYTES (SLK_G_OFF_PEND + 2u * SLK_G_PEND)
static_assert(SLK_G_CAP % 32 == 0 && SLK_G_CAP <= 2047 && SLK_G_CAP >= 32 * SLK_G_MAXNE, "SLK_G_CAP");
static_assert(SLK_G_HIST * 256 <= 8 * SLK_G_CAP, "the histogram must fit the key array");
static_assert(SLK_G_OFF_HITS % 8 == 0, "alignment of the hit slots");

// classes of a k-mer window beyond the 62-bit minimizer values
#define SLK_V_AMB (1ull << 62)    // every base of the window is ambiguous: part of an AMBIGUOUS span
#define SLK_V_NONE (2ull << 62)   // some but not all bases ambiguous (or no window at all): belongs to no span

#ifndef SLK_MAX_THRESHOLDS
#define SLK_MAX_THRESHOLDS 8
#endif
struct slk_group_thresholds {   // ClassifyParams.thresholds: row t of the outputs (t >= 1) starts at taxon_out + t * stride
  double confidence[SLK_MAX_THRESHOLDS];
  uint32_t n;
  uint64_t stride;
  int32_t* taxon_out; uint8_t* flags_out;
};
struct slk_group_in {   // one mate of the batch, packed form (include/slacken_gpu.h, "Packed input")
  const uint64_t* codes; const uint32_t* mask; const uint64_t* boff; const uint32_t* len; uint64_t shift;
};

#if defined(__CUDA_ARCH__)
#define SLK_G_LDG32(p) __ldg(p)
__device__ __forceinline__ void slk_g_load_bucket(const uint64_t* p, slk_bucket* o) {
#ifndef SLK_G_LD
#define SLK_G_LD "ld.global.nc.L1::no_allocate.v4.u64"
#endif
  asm volatile(SLK_G_LD " {%0,%1,%2,%3}, [%4];" : "=l"(o->c0), "=l"(o->c1), "=l"(o->c2), "=l"(o->c3) : "l"(p));
}
#else
#define SLK_G_LDG32(p) (*(p))
inline void slk_g_load_bucket(const uint64_t* p, slk_bucket* o) { o->c0 = p[0]; o->c1 = p[1]; o->c2 = p[2]; o->c3 = p[3]; }
#endif

__device__ __forceinline__ uint32_t slk_g_pairrev32(uint32_t x) {   // reverses the order of the 16 bit pairs of a word
  uint32_t y = __brev(x);
  return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}
// 64 bits of a little-endian word stream w[0..] starting at bit `bit` (compile-time constant after unrolling)
__device__ __forceinline__ uint64_t slk_g_extract(const uint32_t* w, uint32_t bit) {
  const uint32_t i = bit >> 5, s = bit & 31u;
  const uint32_t lo = s ? __funnelshift_r(w[i], w[i + 1], s) : w[i];
  const uint32_t hi = s ? __funnelshift_r(w[i + 1], w[i + 2], s) : w[i + 1];
  return ((uint64_t)hi << 32) | lo;
}

// The minimizers of the 16 k-mer windows of one chunk, position-parallel.
//   cw[0..3]: the 2-bit codes of bases 16c .. 16c+63 of the read (base i in bits 2i, 2i+1)
// V[j] = minimizer (right-aligned, < 2^62) of the window that starts at base 16c + j.
// kmers/minimizer/ShiftScanner.scala:90-159, MinimizerPriorities.scala:144-175,287-312, PosRankWindow.scala:47-74.
template <int W, bool CANON>
__device__ __forceinline__ void slk_g_chunk(const uint32_t* cw, int m, uint64_t mmask, uint64_t xor_mask, uint64_t sig_mask, uint64_t* V) {
  constexpr int NX = SLK_G_CHUNK + W - 1;   // m-mer positions the chunk's windows look at
  // forward strand: the m-mer as a number has its FIRST base on top, so it is cut out of the pair-reversed stream
  uint32_t rv[6];
  rv[0] = slk_g_pairrev32(cw[3]); rv[1] = slk_g_pairrev32(cw[2]); rv[2] = slk_g_pairrev32(cw[1]); rv[3] = slk_g_pairrev32(cw[0]);
  rv[4] = 0; rv[5] = 0;
  // rv holds base i at pair position 63 - i. Shifting it right by 2 * (31 - m) makes the cut position independent of m.
  uint32_t sh = 2u * (uint32_t)(31 - m);
  if (sh >= 32u) { rv[0] = rv[1]; rv[1] = rv[2]; rv[2] = rv[3]; rv[3] = 0; sh -= 32u; }
  if (sh) {
    rv[0] = __funnelshift_r(rv[0], rv[1], sh); rv[1] = __funnelshift_r(rv[1], rv[2], sh);
    rv[2] = __funnelshift_r(rv[2], rv[3], sh); rv[3] >>= sh;
  }
  uint32_t cx[6] = {cw[0], cw[1], cw[2], cw[3], 0u, 0u};
  uint64_t x[NX];
#pragma unroll
  for (int j = 0; j < NX; j++) {
    // reverse complement: base j sits lowest in the stream as it is, complemented (NTBitArray.scala:231-266)
    const uint64_t rc = ~slk_g_extract(cx, 2u * (uint32_t)j) & mmask;
    const uint64_t fw = slk_g_extract(rv, 2u * (uint32_t)(33 - j)) & mmask;
    const uint64_t cn = CANON ? slk_min62(fw, rc) : fw;
    x[j] = (cn ^ xor_mask) & sig_mask;
  }
  // minimum over the W m-mers of every window: a tree whose inner nodes neighbouring windows share
  constexpr int P = W >= 8 ? 8 : W >= 4 ? 4 : W >= 2 ? 2 : 1;   // largest power of two <= W
  uint64_t l1[NX], l2[NX], l3[NX];
#pragma unroll
  for (int j = 0; j < NX; j++) { l1[j] = x[j]; l2[j] = x[j]; l3[j] = x[j]; }
  if (P >= 2) {
#pragma unroll
    for (int j = 0; j + 1 < NX; j++) l1[j] = slk_min62(x[j], x[j + 1]);
  }
  if (P >= 4) {
#pragma unroll
    for (int j = 0; j + 3 < NX; j++) l2[j] = slk_min62(l1[j], l1[j + 2]);
  }
  if (P >= 8) {
#pragma unroll
    for (int j = 0; j + 7 < NX; j++) l3[j] = slk_min62(l2[j], l2[j + 4]);
  }
#pragma unroll
  for (int j = 0; j < SLK_G_CHUNK; j++) {
    const uint64_t a = P == 1 ? x[j] : P == 2 ? l1[j] : P == 4 ? l2[j] : l3[j];
    const uint64_t b = P == 1 ? x[j + W - P] : P == 2 ? l1[j + W - P] : P == 4 ? l2[j + W - P] : l3[j + W - P];
    V[j] = W == P ? a : slk_min62(a, b);
  }
}
// Supermers.splitByAmbiguity (slacken/Supermers.scala:150-189) on a chunk with ambiguous bases (am: bit i = base 16c+i is
// ambiguous): a window with an ambiguous base belongs to no sequence span; one that lies entirely inside an ambiguous
// stretch is one k-mer of an AMBIGUOUS span.
__device__ __forceinline__ void slk_g_chunk_ambiguity(uint64_t am, uint32_t k, uint64_t* V) {
  const uint64_t kmask = (1ull << k) - 1ull;   // k <= 38
#pragma unroll
  for (int j = 0; j < SLK_G_CHUNK; j++) {
    const uint64_t t = (am >> j) & kmask;
    V[j] = t == 0 ? V[j] : (t == kmask ? SLK_V_AMB : SLK_V_NONE);
  }
}

// slk_compress_fast for the same minimizer right-aligned by two bits (m = 31: fshift = 2)
__device__ __forceinline__ uint64_t slk_compress_fast_r2(uint64_t xr) {
  uint32_t y = (uint32_t)xr & 0x33333333u;
  y = (y | (y >> 2)) & 0x0f0f0f0fu;
  y = (y | (y >> 4)) & 0x00ff00ffu;
  y = (y | (y >> 8)) & 0x0000ffffu;
  return ((xr >> 30) << 16) | y;
}

// entry types in the meta byte (bits 6-7); bits 0-5 = k-mer windows
#define SLK_G_T_SEQ 0u
#define SLK_G_T_AMB 1u
#define SLK_G_T_BORDER 2u

// Merged hits of a lane's fragment beyond the SLK_SHITS kept in shared memory, and the slow-path histogram. These are
// indexed dynamically, so they live in local memory; only fragments with many hits / many taxa ever touch them.
struct slk_group_overflow {
  int32_t xh_label[SLK_XHITS], xh_count[SLK_XHITS];
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t nk;
  bool overflow;
  __device__ __forceinline__ uint32_t size() const { return nk; }
  __device__ __forceinline__ void at(uint32_t i, uint32_t* t, int32_t* v) const { *t = hk[i]; *v = hv[i]; }
  __device__ __forceinline__ int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) return hv[i];
    return 0;
  }
  __device__ __forceinline__ void add(uint32_t t, int32_t c) {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) { hv[i] += c; return; }
    if (nk == SLK_KMAX) { overflow = true; return; }
    hk[nk] = t; hv[nk] = c; nk++;
  }
};

// the per-warp shared memory, addressed through one base pointer that the compiler can see is shared memory
struct slk_group_smem {
  uint8_t* sm;
  uint32_t lane;
  __device__ __forceinline__ uint64_t& key(uint32_t s) const { return reinterpret_cast<uint64_t*>(sm)[s]; }
  __device__ __forceinline__ uint16_t& tab(uint32_t t) const { return reinterpret_cast<uint16_t*>(sm + SLK_G_OFF_TABLE)[t * 32u + lane]; }
  __device__ __forceinline__ uint8_t& meta(uint32_t s) const { return (sm + SLK_G_OFF_META)[s]; }
  __device__ __forceinline__ uint64_t& hit_slot(uint32_t i) const { return reinterpret_cast<uint64_t*>(sm + SLK_G_OFF_HITS)[i * 32u + lane]; }
  __device__ __forceinline__ uint64_t& hist_slot(uint32_t i) const { return reinterpret_cast<uint64_t*>(sm)[i * 32u + lane]; }
  __device__ __forceinline__ uint16_t& pend(uint32_t q) const { return reinterpret_cast<uint16_t*>(sm + SLK_G_OFF_PEND)[q]; }
};
// the first SLK_G_HIST pairs of a lane's (taxon -> k-mers) histogram live in the idle key array
struct slk_group_fast_hist {
  slk_group_smem S;
  uint32_t n;
  __device__ __forceinline__ uint32_t size() const { return n; }
  __device__ __forceinline__ void at(uint32_t i, uint32_t* t, int32_t* v) const { const uint64_t x = S.hist_slot(i); *t = (uint32_t)x; *v = (int32_t)(x >> 32); }
  __device__ __forceinline__ int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t x = S.hist_slot(i);
      if ((uint32_t)x == t) return (int32_t)(x >> 32);
    }
    return 0;
  }
  __device__ __forceinline__ bool add(uint32_t t, int32_t c) {
    for (uint32_t i = 0; i < n; i++) {
      const uint64_t x = S.hist_slot(i);
      if ((uint32_t)x == t) { S.hist_slot(i) = ((uint64_t)(uint32_t)((int32_t)(x >> 32) + c) << 32) | t; return true; }
    }
    if (n == SLK_G_HIST) return false;
    S.hist_slot(n) = ((uint64_t)(uint32_t)c << 32) | t; n++;
    return true;
  }
};

__device__ __forceinline__ void slk_group_buffered_hit(const slk_group_smem& S, const slk_group_overflow& ov, uint32_t i, int32_t* label,
                                                       int32_t* count) {
  if (i < SLK_SHITS) { const uint64_t h = S.hit_slot(i); *label = (int32_t)(uint32_t)h; *count = (int32_t)(uint32_t)(h >> 32); }
  else { *label = ov.xh_label[i - SLK_SHITS]; *count = ov.xh_count[i - SLK_SHITS]; }
}
// TaxonCounts.toMap (slacken/TaxonCounts.scala:70-81) into the slow-path histogram: skips AMBIGUOUS / MATE_PAIR_BORDER
static __device__ __noinline__ void slk_group_fold_hits(const slk_group_smem S, slk_group_overflow& ov, uint32_t n) {
  for (uint32_t i = 0; i < n; i++) {
    int32_t l, c;
    slk_group_buffered_hit(S, ov, i, &l, &c);
    if (l >= 0) ov.add((uint32_t)l, c);
  }
}
// The rare fragment whose merged hits outgrow the buffers: fold them into the slow-path histogram and move them to the sink.
template <class Sink>
static __device__ __noinline__ void slk_group_spill(const slk_group_smem S, slk_group_overflow& ov, Sink& sink, const int32_t* raw, uint32_t n,
                                             uint32_t need) {
  slk_group_fold_hits(S, ov, n);
  for (uint32_t i = 0; i < n; i++) {
    int32_t l, c;
    slk_group_buffered_hit(S, ov, i, &l, &c);
    sink.push(l >= 0 ? raw[l] : l, c, need + n);
  }
}
static __device__ __noinline__ uint32_t slk_group_resolve_slow(const slk_group_overflow& ov, const slk_tax_view& tx, double confidence, int32_t total) {
  return slk_resolve_tree(ov, tx, confidence, total);
}

// One fragment per lane, end to end. Every lane of the warp calls it (lanes without a fragment pass live = false).
// `sink` receives the hits of the rare fragment whose hit list outgrows the buffers (see dev_hit_sink), and reserves the
// place of everybody else's hits in the output.
template <int W, bool CANON, class Sink>
__device__ __forceinline__ void slk_group_classify(uint8_t* sm_warp, const slk_scan_params& sp, const slk_table_view& tb,
                                                   const slk_tax_view& tx, Sink& sink, slk_group_overflow& ov,
                                                   const slk_group_in& in1, const slk_group_in& in2, bool paired, bool live,
                                                   uint64_t r, const slk_group_thresholds& mt, int32_t min_hit_groups, slk_frag_result& res) {
  const uint32_t lane = threadIdx.x & 31u;
  const slk_group_smem S{sm_warp, lane};
  const uint32_t k = (uint32_t)sp.k, km1 = k - 1;
  const int m = sp.m;
  const slk_scan_masks smk(sp);
  const uint64_t mmask = smk.mmask, xor_mask = smk.xor_mask, sig_mask = smk.sig_mask;
  const int fshift = sp.fshift;
  const bool fast = sp.fast_compress != 0;
  const int32_t border_cnt = -(sp.k - 1);

  // ---- merge state of this lane's fragment (slacken/Classifier.scala:92-95, slacken/TaxonCounts.scala:31-48): registers
  uint64_t l_last = 0;
  bool l_have_last = false, l_have_cur = false, l_spilled = false, any = false;
  int32_t l_label = 0, l_count = 0;
  uint32_t l_mate = 0, l_k0 = 0, l_k1 = 0, l_nd = 0, l_np = 0, l_nh = 0, nh_spilled = 0;
  // the usual fragment hits ONE taxon (plus misses): resolveTree is then a comparison, no histogram needed
  uint32_t l_t1 = 0, l_c1 = 0;
  bool l_multi = false;
  ov.nk = 0; ov.overflow = false;

  // ---- this lane's fragment (no arrays indexed by the mate: they would live in local memory)
  uint32_t len0 = 0, len1 = 0;
  const uint32_t *cw0 = nullptr, *cw1 = nullptr, *mk0 = nullptr, *mk1 = nullptr;
  if (live) {
    const uint64_t b1 = in1.boff[r] - in1.shift;
    len0 = in1.len[r]; cw0 = reinterpret_cast<const uint32_t*>(in1.codes + b1); mk0 = in1.mask + b1;
    if (paired) {
      const uint64_t b2 = in2.boff[r] - in2.shift;
      len1 = in2.len[r]; cw1 = reinterpret_cast<const uint32_t*>(in2.codes + b2); mk1 = in2.mask + b2;
    }
  }
  // upper bound of the merged hits of the fragment (for the sink's spill allocation)
  const uint32_t max_hits = (len0 > km1 ? len0 - km1 : 0) + (paired ? (len1 > km1 ? len1 - km1 : 0) + 1 : 0) + 2;

  auto buffered_hit = [&](uint32_t i, int32_t* label, int32_t* count) { slk_group_buffered_hit(S, ov, i, label, count); };
  // only fragments with more than SLK_SHITS + SLK_XHITS merged hits come here: their hits go to a block of their own
  auto spill = [&](uint32_t need) {
    slk_group_spill(S, ov, sink, tx.raw, l_nh, need);
    nh_spilled += l_nh;
    l_nh = 0;
    l_spilled = true;
  };
  // The hit that is open (l_label, l_count) is always stored at index l_nh, so that closing it is only an increment:
  // TaxonCounts.fromHits merges adjacent hits with the same taxon (slacken/TaxonCounts.scala:31-48)
  auto store_open_hit = [&]() {
    if (l_nh < SLK_SHITS) S.hit_slot(l_nh) = ((uint64_t)(uint32_t)l_count << 32) | (uint32_t)l_label;
    else { ov.xh_label[l_nh - SLK_SHITS] = l_label; ov.xh_count[l_nh - SLK_SHITS] = l_count; }
  };

  // scan cursor: mate, chunk, and the run that is open (its minimizer / class, its k-mer count)
  uint32_t mate = 0, c = 0;
  uint32_t nwin_m = len0 >= k ? len0 - km1 : 0;                            // k-mer windows of the current mate
  uint32_t nch = (nwin_m + SLK_G_CHUNK - 1) / SLK_G_CHUNK;
  bool done = !live, pend_border = false;
  uint64_t run_val = SLK_V_NONE;
  uint32_t run_cnt = 0;
  uint32_t n_buf = 0, n_steps = 0;     // warp-uniform: entries in the buffer, steps since the last close
  uint32_t my_rows = 0, my_total = 0;  // this lane's rows of the step table (steps that brought it entries) and its entries
  bool flush = false, finished = false;
  // the chunk the lane scans next, requested one step ahead
  bool have = false;
  uint32_t cw[4] = {0, 0, 0, 0}, mw0 = 0, mw1 = 0, mw2 = 0;
  auto prepare = [&]() {
    have = false;
    if (!done) {
      if (c >= nch && mate == 0 && paired) {   // mate 1 without a single window: straight on to the border and mate 2
        pend_border = true; mate = 1; c = 0;
        nwin_m = len1 >= k ? len1 - km1 : 0;
        nch = (nwin_m + SLK_G_CHUNK - 1) / SLK_G_CHUNK;
      }
      have = c < nch;
      if (!have && !pend_border) done = true;
    }
    if (have) {
      const uint32_t L = mate ? len1 : len0, nmw = (L + 31u) >> 5, nw32 = nmw * 2u, b = c >> 1;   // mask words, 32-bit code words
      const uint32_t* p = mate ? cw1 : cw0;
      const uint32_t* q = mate ? mk1 : mk0;
#pragma unroll
      for (uint32_t i = 0; i < 4; i++) cw[i] = c + i < nw32 ? SLK_G_LDG32(p + c + i) : 0u;
      mw0 = b < nmw ? SLK_G_LDG32(q + b) : 0u; mw1 = b + 1 < nmw ? SLK_G_LDG32(q + b + 1) : 0u;
      mw2 = b + 2 < nmw ? SLK_G_LDG32(q + b + 2) : 0u;
    }
  };
  prepare();

  for (;;) {
    if (flush) {
      // ================= close the buffer: look every sequence entry up, then every lane merges its own entries
      __syncwarp();   // the entries other lanes wrote are visible
      // 1. lookups (spanToHit's join, slacken/KeyValueIndex.scala:176-185): flat over the buffer, a rolling pipeline of
      //    SLK_G_DEPTH buckets in flight per lane. The dense taxon goes to the top 16 bits of the entry's key slot. An
      //    entry whose bucket was full without a match goes to a short pending list; its chain goes on in the same
      //    128-byte line, and the list is worked off 32 entries at a time so that those loads overlap as well.
      {
        uint64_t ck[SLK_G_DEPTH];
        slk_bucket bk[SLK_G_DEPTH];
        uint32_t n_pend = 0;
        const uint32_t rounds = (n_buf + 31u) >> 5;
#pragma unroll
        for (int d = 0; d < SLK_G_DEPTH; d++) { ck[d] = ~0ull; bk[d].c0 = 0; bk[d].c1 = 0; bk[d].c2 = 0; bk[d].c3 = 0; }
        auto work_off = [&]() {
          __syncwarp();
          for (uint32_t q = lane; q < n_pend; q += 32u) {
            const uint32_t s = S.pend(q);
            const uint64_t kq = S.key(s) & 0xffffffffffffull;
            const uint32_t dense = slk_probe_rest(tb, slk_bucket_of(kq, tb), 1, kq);
            S.key(s) = kq | ((uint64_t)dense << 48);
          }
          __syncwarp();
          n_pend = 0;
        };
        for (uint32_t r0 = 0; r0 < rounds + SLK_G_DEPTH; r0 += SLK_G_DEPTH) {
#pragma unroll
          for (int d = 0; d < SLK_G_DEPTH; d++) {
            const uint32_t rr = r0 + (uint32_t)d;
            // the bucket requested SLK_G_DEPTH rounds ago has landed (slot d of the ring)
            bool pending = false;
            if (ck[d] != ~0ull) {
              const uint32_t s = (rr - SLK_G_DEPTH) * 32u + lane;
              uint32_t dense;
              pending = !slk_match_bucket(bk[d], ck[d], &dense);
              S.key(s) = ck[d] | ((uint64_t)dense << 48);
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, pending);
            if (bal) {
              if (pending) S.pend(n_pend + (uint32_t)__popc(bal & ((1u << lane) - 1u))) = (uint16_t)((rr - SLK_G_DEPTH) * 32u + lane);
              n_pend += (uint32_t)__popc(bal);
              if (n_pend > SLK_G_PEND - 32u) work_off();
            }
            // request the bucket of this round's entry
            const uint32_t s = rr * 32u + lane;
            ck[d] = ~0ull;
            if (s < n_buf && (S.meta(s) >> 6) == SLK_G_T_SEQ) {
              const uint64_t kr = S.key(s);   // right-aligned; the left-aligned priority is kr << fshift
              ck[d] = fast ? slk_compress_fast_r2(kr) : slk_compress_generic(sp, kr << fshift);
              slk_g_load_bucket(tb.cells + slk_bucket_of(ck[d], tb) * 4, &bk[d]);
            }
          }
        }
        if (n_pend) work_off();
      }
      __syncwarp();   // all labels are visible to the lanes that own the entries
      // 2. merge: every lane walks its own entries in span order, row by row of the step table: numDistinct
      //    (slacken/Classifier.scala:94), k-mer totals and TaxonCounts.fromHits (slacken/TaxonCounts.scala:31-48,83-87)
      {
        uint32_t t = 0, i = 0, base = 0, cnt = 0;
        for (uint32_t it = 0; it < my_total; it++) {
          if (i == cnt) { const uint32_t e = S.tab(t); t++; base = e >> 5; cnt = e & 31u; i = 0; }   // rows are never empty
          const uint64_t kl = S.key(base + i);
          const uint32_t mt = S.meta(base + i);
          i++;
          const uint32_t type = mt >> 6, n = mt & 63u;
          const bool is_seq = type == SLK_G_T_SEQ, is_border = type == SLK_G_T_BORDER;
          const uint64_t ck = kl & 0xffffffffffffull;
          const uint32_t dense = (uint32_t)(kl >> 48);
          l_nd += (is_seq && (!l_have_last || ck != l_last) && dense != 0) ? 1u : 0u;
          const bool nz = is_seq && dense != 0;
          l_t1 = (nz && l_t1 == 0) ? dense : l_t1;
          l_c1 += (nz && dense == l_t1) ? n : 0u;
          l_multi = l_multi || (nz && dense != l_t1);
          l_last = is_seq ? ck : l_last;
          l_have_last = l_have_last || is_seq;
          const int32_t label = is_seq ? (int32_t)dense : (is_border ? SLK_MATE_PAIR_BORDER : SLK_AMBIGUOUS_SPAN);
          const int32_t hcnt = is_border ? border_cnt : (int32_t)n;
          l_k0 += (!is_border && l_mate == 0) ? n : 0u;
          l_k1 += (!is_border && l_mate != 0) ? n : 0u;
          l_mate = is_border ? 1u : l_mate;
          const bool same = l_have_cur && label == l_label;
          if (l_have_cur && !same) {   // the open hit is complete (it is stored already): the next one opens
            l_nh++;
            if (l_nh == SLK_SHITS + SLK_XHITS) spill(max_hits - l_nh);
          }
          l_count = same ? l_count + hcnt : hcnt;
          l_label = label; l_have_cur = true;
          store_open_hit();
        }
      }
      __syncwarp();   // nobody writes the buffer before everybody has left it
      n_buf = 0; n_steps = 0; my_rows = 0; my_total = 0; flush = false;
      if (finished) break;
    }
    if (!__any_sync(0xffffffffu, !done)) { finished = true; flush = true; continue; }   // the last close, same code
    // ================= one step: every lane scans the chunk that prepare() requested
    uint32_t nwin = 0;
    bool fin = false;
    uint64_t am = 0;
    if (have) {
      const uint32_t s16 = (c & 1u) * 16u;
      am = ((uint64_t)(s16 ? __funnelshift_r(mw1, mw2, s16) : mw1) << 32) | (s16 ? __funnelshift_r(mw0, mw1, s16) : mw0);
      const uint32_t w0 = c * SLK_G_CHUNK;
      nwin = nwin_m - w0 < SLK_G_CHUNK ? nwin_m - w0 : SLK_G_CHUNK;
      fin = c + 1 == nch;
    }
    // ---- the chunk's minimizers, position-parallel (a warp without ambiguous bases takes the version without masks)
    uint64_t V[SLK_G_CHUNK];
    const bool dirty = __any_sync(0xffffffffu, am != 0 || run_val == SLK_V_AMB);   // (an open AMBIGUOUS run counts: its entry needs the type)
    slk_g_chunk<W, CANON>(cw, m, mmask, xor_mask, sig_mask, V);
    if (dirty) slk_g_chunk_ambiguity(am, k, V);
    // ---- run ends (MinSplitter.scala:180-216: consecutive windows with an equal minimizer VALUE are one super-mer).
    // A run that could outgrow the 6-bit count during this chunk is cut first (an entry of its own; invisible in every output).
    const bool split = have && run_cnt > SLK_G_CNT_MAX - SLK_G_CHUNK;
    const uint64_t split_val = run_val;
    const uint32_t split_cnt = run_cnt;
    if (split) run_cnt = 0;   // the run stays open (run_val), its count starts again
    // bit j of `ends`: the open run ends before window j; bit j of `emit`: ... and it is a span; bit j of `ambs`: an AMBIGUOUS one
    const uint32_t inm = (1u << nwin) - 1u;   // nwin <= 16
    uint32_t ends = V[0] != run_val ? 1u : 0u;
#pragma unroll
    for (int j = 1; j < SLK_G_CHUNK; j++) ends |= (V[j] != V[j - 1] ? 1u : 0u) << j;
    ends &= inm;
    uint32_t nones = run_cnt == 0 ? 1u : 0u, ambs = run_val == SLK_V_AMB ? 1u : 0u;   // class of the window BEFORE j
    if (dirty) {
#pragma unroll
      for (int j = 0; j + 1 < SLK_G_CHUNK; j++) {
        nones |= (V[j] >= SLK_V_NONE ? 1u : 0u) << (j + 1);
        ambs |= (V[j] == SLK_V_AMB ? 1u : 0u) << (j + 1);
      }
    }
    const uint32_t emit = ends & ~nones;
    // the run that is open after the chunk's last window
    uint64_t tail_val = run_val;
    uint32_t tail_cnt = run_cnt + nwin;
    if (nwin) {
      tail_val = V[SLK_G_CHUNK - 1];
      if (nwin < SLK_G_CHUNK) {
#pragma unroll
        for (int j = 0; j + 1 < SLK_G_CHUNK; j++) tail_val = (uint32_t)j + 1u == nwin ? V[j] : tail_val;
      }
      if (ends) tail_cnt = nwin - (31u - (uint32_t)__clz((int)ends));
    }
    // ---- entries of this step: [border] [cut run] [one per run end] [the run still open at the end of the mate]
    const bool fin_emit = fin && tail_val < SLK_V_NONE && tail_cnt != 0;
    const uint32_t ne = (pend_border ? 1u : 0u) + (split ? 1u : 0u) + (uint32_t)__popc(emit) + (fin_emit ? 1u : 0u);
    uint32_t incl = ne;   // place of this lane's entries: warp prefix sum
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (n_buf + total > SLK_G_CAP || n_steps == SLK_G_STEPS) {   // close first, then redo this step
      if (split) run_cnt = split_cnt;
      flush = true;
      continue;
    }
    uint32_t slot = n_buf + incl - ne;
    if (ne) { S.tab(my_rows) = (uint16_t)((slot << 5) | ne); my_rows++; my_total += ne; }
    any = any || ne != 0;
    if (pend_border) { S.key(slot) = 0; S.meta(slot) = (uint8_t)(SLK_G_T_BORDER << 6); slot++; pend_border = false; }
    if (split) {
      const bool amb = split_val == SLK_V_AMB;
      S.key(slot) = split_val; S.meta(slot) = (uint8_t)(split_cnt | ((amb ? SLK_G_T_AMB : SLK_G_T_SEQ) << 6)); slot++;
      l_np += amb ? 0u : 1u;
    }
    // run end j closes the run whose last window is j - 1: its minimizer is V[j - 1], its length j - (previous run end)
    if (ne) {
      uint64_t* kp = &S.key(slot);
      uint8_t* mp = &S.meta(slot);
      int32_t last = -(int32_t)run_cnt;
      if (!dirty) {   // no ambiguous base in the warp's chunks: every run end is emitted (but a mate's first), all are sequence spans
#pragma unroll
        for (int j = 0; j < SLK_G_CHUNK; j++) {
          if ((emit >> j) & 1u) {
            *kp++ = j == 0 ? run_val : V[j > 0 ? j - 1 : 0];
            *mp++ = (uint8_t)(j - last);
            last = j;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < SLK_G_CHUNK; j++) {
          if ((emit >> j) & 1u) {
            *kp++ = j == 0 ? run_val : V[j > 0 ? j - 1 : 0];
            *mp++ = (uint8_t)((uint32_t)(j - last) | (((ambs >> j) & 1u) << 6));   // SLK_G_T_AMB == 1
          }
          last = ((ends >> j) & 1u) ? j : last;
        }
      }
      slot += (uint32_t)__popc(emit);
      l_np += (uint32_t)__popc(emit & ~ambs);
    }
    if (fin_emit) {
      const bool amb = tail_val == SLK_V_AMB;
      S.key(slot) = tail_val;
      S.meta(slot) = (uint8_t)(tail_cnt | ((amb ? SLK_G_T_AMB : SLK_G_T_SEQ) << 6));
      l_np += amb ? 0u : 1u;
    }
    n_buf += total; n_steps++;
    // ---- advance the cursor and request the next chunk
    if (have) {
      run_val = tail_val; run_cnt = tail_val < SLK_V_NONE ? tail_cnt : 0u;   // windows outside any span are not counted
      c++;
      if (fin) {   // Supermers.splitFragment (slacken/Supermers.scala:49-97): mate 1, the border pseudo-span, mate 2
        run_val = SLK_V_NONE; run_cnt = 0;
        if (mate == 0 && paired) {
          pend_border = true; mate = 1; c = 0;
          nwin_m = len1 >= k ? len1 - km1 : 0;
          nch = (nwin_m + SLK_G_CHUNK - 1) / SLK_G_CHUNK;
        } else {
          done = true;
        }
      }
    }
    prepare();
  }
  if (l_have_cur) l_nh++;   // the last hit (stored already)
  // The number of buffered hits is final: their place in the output is reserved now (one atomic per warp, whose latency
  // the resolve step hides). All lanes call it.
  sink.reserve(l_spilled ? 0u : l_nh);
  // Classifier.classify (slacken/Classifier.scala:156-170) keeps the hits and classifies them once per confidence threshold:
  // here the histogram is built once and resolveTree runs once per threshold. Threshold 0 goes to `res`, the others
  // straight to their rows of the outputs.
  const int32_t total_kmers = (int32_t)(l_k0 + l_k1);
  int mode = 0;   // 0: one hit taxon at most; 1: histogram in the key array; 2: slow-path histogram
  slk_group_fast_hist fh{S, 0u};
  if (!l_multi) {
    if (l_spilled) spill(0);
  } else if (!l_spilled) {   // the histogram fits the (now idle) key array
    bool fits = true;
    for (uint32_t i = 0; i < l_nh && fits; i++) {
      int32_t l, cc;
      buffered_hit(i, &l, &cc);
      if (l >= 0) fits = fh.add((uint32_t)l, cc);
    }
    mode = fits ? 1 : 2;
    if (!fits) slk_group_fold_hits(S, ov, l_nh);
  } else {
    spill(0);   // a fragment that went to the sink keeps all its hits there (and folded them into the slow-path histogram)
    mode = 2;
  }
  for (uint32_t t = 0; t < mt.n; t++) {
    const double confidence = mt.confidence[t];
    uint32_t taxon;
    if (mode == 0) {
      // One hit taxon (or none): its root-path score is its own count and every clade on the way up holds exactly that
      // count, so LowestCommonAncestor.resolveTree (slacken/LowestCommonAncestor.scala:91-146) returns it when the count
      // reaches ceil(confidence * totalKmers) and NONE otherwise.
      const double required = ceil(confidence * (double)total_kmers);
      taxon = (l_t1 != 0 && (double)(int32_t)l_c1 >= required) ? l_t1 : 0u;
    } else if (mode == 1) {
      taxon = slk_resolve_tree(fh, tx, confidence, total_kmers);
    } else {
      taxon = slk_group_resolve_slow(ov, tx, confidence, total_kmers);
    }
    const bool classified = taxon != 0 && l_nd >= (uint32_t)min_hit_groups;   // slacken/Classifier.scala:446
    const int32_t raw = classified ? tx.raw[taxon] : 0;
    const uint32_t fl = (classified ? SLK_F_CLASSIFIED : 0u) | (any ? SLK_F_HAS_SPAN : 0u) | (ov.overflow ? SLK_F_OVERFLOW : 0u);
    if (t == 0) { res.taxon = raw; res.flags = fl; }
    else if (live) { mt.taxon_out[(uint64_t)t * mt.stride + r] = raw; mt.flags_out[(uint64_t)t * mt.stride + r] = (uint8_t)(fl & 3u); }
  }
  res.kmers1 = l_k0; res.kmers2 = l_k1;
  res.num_distinct = l_nd; res.n_hits = l_nh + nh_spilled; res.n_probes = l_np;
}

// ------------------------------------------------------------------------------------------------ kernel body
// all 32 lanes must call; returns each lane's offset in a warp-wide contiguous allocation
__device__ __forceinline__ uint64_t slk_g_warp_alloc(unsigned long long* cursor, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t incl = n;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= (uint32_t)d) incl += v;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (lane == 0 && total) base = atomicAdd(cursor, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 0);
  return base + incl - n;
}
// Where merged hits go. Ordinary fragments: a compact, warp-allocated block of the hit buffer, written at the end of the
// kernel from the per-lane buffers. A fragment with more merged hits than the buffers hold takes a worst-case block for
// itself (one atomic) and appends there.
struct slk_g_hit_sink {
  uint32_t n, alloc;      // hits pushed to the private block, and the block's size
  bool spilled, live, enabled;
  uint64_t goff, reserved;
  slk_hit* gbase;         // indexed by (absolute index - gshift)
  uint64_t gshift, gcap;
  unsigned long long* cursor;
  __device__ __forceinline__ void put(uint64_t abs_idx, int32_t taxon, int32_t count) {
    const uint64_t rel = abs_idx - gshift;
    slk_hit h;
    h.taxon = taxon; h.count = count;
    if (rel < gcap) gbase[rel] = h;
  }
  __device__ __forceinline__ void reserve(uint32_t n_hits) {   // all 32 lanes call it
    if (enabled) reserved = slk_g_warp_alloc(cursor, live ? n_hits : 0u);
  }
  __device__ __forceinline__ void push(int32_t taxon, int32_t count, uint32_t need) {
    if (!enabled) return;
    if (!spilled) { alloc = need + 2u; goff = atomicAdd(cursor, (unsigned long long)alloc); spilled = true; }
    put(goff + n, taxon, count);
    n++;
  }
};

struct slk_classify2_args {
  slk_scan_params sp; slk_table_view tb; slk_tax_view tx;
  slk_group_in in1, in2; uint32_t paired; uint32_t n_reads;
  slk_group_thresholds mt; int32_t min_hit_groups; uint32_t hits;
  int32_t* taxon_out; uint8_t* flags_out; slk_read_detail* detail_out;
  slk_hit* hits_base; const unsigned long long* hits_shift_ptr; uint64_t hits_cap; unsigned long long* hits_cursor;
  unsigned long long* hits_over;   // (optional) slots reserved beyond the hits written: cursor advance - *hits_over = merged hits
  unsigned long long* counts; uint32_t* error_flag; unsigned long long* stats;
};

// The work of one thread of classify2_kernel: fragment r = blockIdx.x * blockDim.x + threadIdx.x, its warp's slice of the
// block's shared memory at smem + (threadIdx.x / 32) * SLK_G_WARP_BYTES.
template <int W, bool CANON>
__device__ __forceinline__ void slk_classify2_thread(const slk_classify2_args& a, uint8_t* smem) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < a.n_reads, paired = a.paired != 0, want_hits = a.hits != 0;
  const uint32_t lane = threadIdx.x & 31u;
  uint8_t* sm_warp = smem + (threadIdx.x >> 5) * SLK_G_WARP_BYTES;
  slk_g_hit_sink sink;
  sink.n = 0; sink.alloc = 0; sink.spilled = false; sink.live = live; sink.enabled = want_hits; sink.goff = 0; sink.reserved = 0;
  sink.gbase = a.hits_base; sink.gshift = (want_hits && a.hits_shift_ptr) ? *a.hits_shift_ptr : 0ull;
  sink.gcap = a.hits_cap; sink.cursor = a.hits_cursor;
  slk_group_overflow ov;
  slk_frag_result res;
  res.taxon = 0; res.flags = 0; res.kmers1 = 0; res.kmers2 = 0; res.num_distinct = 0; res.n_hits = 0; res.n_probes = 0;
  slk_group_classify<W, CANON>(sm_warp, a.sp, a.tb, a.tx, sink, ov, a.in1, a.in2, paired, live, r, a.mt, a.min_hit_groups, res);
  if (live) {
    a.taxon_out[r] = res.taxon;
    a.flags_out[r] = (uint8_t)(res.flags & 3u);
    if (res.flags & SLK_F_OVERFLOW) atomicExch(a.error_flag, 1u);
  }
  if (a.stats != nullptr) {  // probes and merged hits of this launch (the S and H of the roofline arithmetic)
    const uint32_t np = __reduce_add_sync(0xffffffffu, live ? res.n_probes : 0u);
    const uint32_t nh = __reduce_add_sync(0xffffffffu, live ? res.n_hits : 0u);
    if (lane == 0) {
      atomicAdd(&a.stats[0], (unsigned long long)np);
      atomicAdd(&a.stats[1], (unsigned long long)nh);
    }
  }
  // K7: per-taxon report counters (groupBy(sampleId, taxon).count, slacken/Classifier.scala:214-217), aggregated per warp
  // with match_any so that a hot taxon costs one atomic per warp
  if (a.counts != nullptr) {
    const bool cnt = live && (res.flags & SLK_F_HAS_SPAN);
    const uint32_t key = cnt ? (uint32_t)res.taxon : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    if (cnt && lane == (uint32_t)(__ffs((int)peers) - 1)) atomicAdd(&a.counts[(uint32_t)res.taxon], (unsigned long long)__popc(peers));
  }
  if (a.detail_out != nullptr && live) {
    slk_read_detail d;
    d.hit_off = 0; d.hit_cnt = 0;
    if (want_hits) {
      const uint32_t n_buf = sink.spilled ? 0u : res.n_hits;
      const slk_group_smem S{sm_warp, lane};
      for (uint32_t i = 0; i < n_buf; i++) {
        int32_t l, c;
        if (i < SLK_SHITS) { const uint64_t h = S.hit_slot(i); l = (int32_t)(uint32_t)h; c = (int32_t)(uint32_t)(h >> 32); }
        else { l = ov.xh_label[i - SLK_SHITS]; c = ov.xh_count[i - SLK_SHITS]; }
        sink.put(sink.reserved + i, l >= 0 ? a.tx.raw[l] : l, c);   // dense labels become raw taxon ids here
      }
      d.hit_off = sink.spilled ? sink.goff : sink.reserved;
      d.hit_cnt = res.n_hits;
      if (sink.spilled && a.hits_over != nullptr) atomicAdd(a.hits_over, (unsigned long long)(sink.alloc - sink.n));
    }
    d.len1 = res.kmers1 + (uint32_t)(a.sp.k - 1);
    d.len2 = paired ? res.kmers2 + (uint32_t)(a.sp.k - 1) : 0xFFFFFFFFu;
    d.num_distinct = res.num_distinct;
    a.detail_out[r] = d;
  }
}
