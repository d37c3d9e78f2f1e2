// slk_core.h -- per-thread bodies of the sm_100a kernels of the Slacken classify/build hot path.
//
// Everything here is written as plain per-thread functions (no warp collectives), so that the very same
// source can also be compiled by g++ into tests/host_emulation (a TEST-ONLY harness that runs the kernel
// bodies thread by thread on the CPU; it is never loaded by the product path, which fails loudly without a GPU).
//
// Reference semantics reproduced (paths relative to /root/reference/src/main/scala/com/jnpersson/):
//   2-bit codes / validity ........ kmers/util/BitRepresentation.scala:35-39,127-143
//   m-mer priority ................ kmers/minimizer/MinimizerPriorities.scala:144-175,287-312
//                                   kmers/util/NTBitArray.scala:231-266,437-452
//   window minimum + super-mers ... kmers/minimizer/PosRankWindow.scala:47-74, MinSplitter.scala:180-216
//   ambiguity handling, spans ..... slacken/Supermers.scala:49-125,150-189
//   hit labels, numDistinct ....... slacken/KeyValueIndex.scala:176-185, slacken/Classifier.scala:92-95
//   merged hits, totals ........... slacken/TaxonCounts.scala:31-48,70-87,114-121
//   resolveTree, LCA .............. slacken/LowestCommonAncestor.scala:49-146
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SLK_HD __host__ __device__ __forceinline__
#define SLK_HD_NOINLINE __host__ __device__ __noinline__
#else
#define SLK_HD inline
#define SLK_HD_NOINLINE __attribute__((noinline))
#endif

#define SLK_MAX_W 8      // k - m + 1 supported by the kernels (35 - 31 + 1 = 5 for the Kraken 2 defaults)
#define SLK_ECAP 16      // span entries buffered per thread between scan and probe (shared memory on the device)
#define SLK_SHITS 8      // merged hits of a fragment kept in the fast store (shared memory on the device)
#define SLK_XHITS 56     // further merged hits kept in a per-thread overflow array before spilling to global memory
#define SLK_KMAX 128     // distinct taxa per fragment held in the per-thread histogram

// labels of merged hits (slacken/package.scala:28-29)
#define SLK_AMBIGUOUS_SPAN (-1)
#define SLK_MATE_PAIR_BORDER (-2)

// result flags
#define SLK_F_CLASSIFIED 1u
#define SLK_F_HAS_SPAN 2u
#define SLK_F_OVERFLOW 4u   // more than SLK_KMAX distinct taxa in one fragment: the call reports an error

struct slk_scan_params {
  int32_t k, m, w, canonical;
  int32_t fshift;      // 64 - 2m: position of the last base of a left-aligned m-mer
  int32_t key_bits;    // popcount(sig_mask) <= 48
  int32_t fast_compress;  // sig_mask == 0xffffffffcccccccc (m=31, s=7): 11-instruction compress
  int32_t pad_;
  uint64_t xor_mask;   // toggle mask aligned to the m-mer (RandomXOR.mask)
  uint64_t sig_mask;   // bits of a priority that can be non-zero: the space mask, or the m-mer fill mask
  uint64_t mmask;      // fill mask of an m-mer
  uint64_t cmv[6];     // parallel-suffix move masks that compress sig_mask's bits to the right
};

// SpacedSeed.spaceMask (kmers/minimizer/MinimizerPriorities.scala:287-301) for a single-word minimizer
SLK_HD uint64_t slk_space_mask(int m, int spaces) {
  uint64_t r = ~0ull;
  if (m % 32 != 0) r &= (~0ull) << (64 - (m % 32) * 2);
  uint64_t final_bits = 3ull << ((64 - (m % 32) * 2) & 63);
  for (int i = 0; i < spaces; i++) { r <<= 4; r |= final_bits; }
  return r;
}
// Derives the kernel-side parameters from an index's (k, m, minimizerSpaces, XORmask, canonical).
// Returns 0, or 1..4 for: m outside 1..31, k < m or k-m+1 > SLK_MAX_W, bad spaces, more than 48 significant key bits.
SLK_HD int slk_make_scan_params(int k, int m, int spaces, uint64_t toggle_mask, int canonical, slk_scan_params* sp) {
  if (m < 1 || m > 31) return 1;
  if (k < m || k - m + 1 > 8) return 2;
  if (spaces < 0 || spaces > m / 2) return 3;
  sp->k = k; sp->m = m; sp->w = k - m + 1; sp->canonical = canonical ? 1 : 0;
  sp->fshift = 64 - 2 * m;
  sp->mmask = (~0ull) << sp->fshift;
  sp->xor_mask = toggle_mask << sp->fshift;  // RandomXOR.mask (kmers/minimizer/MinimizerPriorities.scala:146-160)
  sp->sig_mask = spaces > 0 ? slk_space_mask(m, spaces) : sp->mmask;
  int bits = 0;
  for (uint64_t x = sp->sig_mask; x; x &= x - 1) bits++;
  sp->key_bits = bits;
  if (bits > 48) return 4;
  sp->fast_compress = sp->sig_mask == 0xffffffffccccccccull ? 1 : 0;
  sp->pad_ = 0;
  // parallel-suffix move masks (Hacker's Delight 7-4) for sig_mask
  uint64_t mm = sp->sig_mask, mk = ~mm << 1;
  for (int i = 0; i < 6; i++) {
    uint64_t mp = mk ^ (mk << 1);
    mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16; mp ^= mp << 32;
    uint64_t mv = mp & mm;
    sp->cmv[i] = mv;
    mm = (mm ^ mv) | (mv >> (1 << i));
    mk &= ~mp;
  }
  return 0;
}

struct slk_table_view {
  uint64_t* cells;     // n_buckets * 4 cells of (compressed key << 16 | dense taxon); 0 = empty
  uint64_t n_buckets;  // one bucket = one 32-byte sector
  uint32_t prefetch;   // issue an L2 prefetch for every bucket of a batch before probing it
  uint32_t pad_;
};

struct slk_tax_view {
  const uint16_t* parent;  // dense parent, dense 0 = NONE
  const uint8_t* depth;    // steps to NONE (depth[0] = 0, a root has depth 1)
  const int32_t* raw;      // dense -> raw taxon id (raw[0] = 0)
  uint32_t n;              // number of dense ids including 0
  uint32_t root;           // dense id of ROOT (raw 1)
};

#ifndef SLACKEN_GPU_H  // same layout as the public slk_hit of include/slacken_gpu.h
struct slk_hit {
  int32_t taxon;
  int32_t count;
};
#endif

// ------------------------------------------------------------------------------------------------ encode
// A=0 C=1 G=2 T=U=3, anything else (including whitespace: the boundary contract is whitespace-free input) = 4.
SLK_HD uint32_t slk_code(uint32_t c) {
  uint32_t u = (c | 0x20u) - 'a';   // a=0 c=2 g=6 t=19 u=20
  uint32_t code = (c >> 1) & 3u;
  code ^= code >> 1;
  bool ok = u < 32u && ((0x00180045u >> (u & 31u)) & 1u);
  return ok ? code : 4u;
}

// Calls f(byte) for every byte of [s, s+len) in order. On the device the bytes arrive through 16-byte aligned
// vector loads (a thread-per-read kernel would otherwise issue one LSU wavefront per base); the buffer must be
// readable up to the next 16-byte boundary, which every library-owned / cudaMalloc'ed buffer is.
template <class F>
SLK_HD void slk_for_each_byte(const uint8_t* s, uint64_t len, F&& f) {
#if defined(__CUDA_ARCH__)
  uintptr_t a0 = reinterpret_cast<uintptr_t>(s);
  uintptr_t a = a0 & ~(uintptr_t)15, aend = a0 + len;
  for (; a < aend; a += 16) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(a));
#pragma unroll 1
    for (int wi = 0; wi < 4; wi++) {
      uint32_t word = wi == 0 ? v.x : wi == 1 ? v.y : wi == 2 ? v.z : v.w;
      uintptr_t wa = a + 4 * wi;
      if (wa + 4 <= a0 || wa >= aend) continue;
#pragma unroll
      for (int bi = 0; bi < 4; bi++) {
        uintptr_t ba = wa + bi;
        if (ba >= a0 && ba < aend) f((word >> (8 * bi)) & 0xffu);
      }
    }
  }
#else
  for (uint64_t i = 0; i < len; i++) f((uint32_t)s[i]);
#endif
}

// ------------------------------------------------------------------------------------------------ key compression
// Hacker's-Delight style compress: gathers the bits of x selected by sig_mask at the low end, keeping their order.
SLK_HD uint64_t slk_compress(const slk_scan_params& sp, uint64_t x) {
  if (sp.fast_compress) {  // the Kraken 2 default mask: keep the high word, gather bit pairs 2-3 of every low nibble
    uint32_t y = ((uint32_t)x >> 2) & 0x33333333u;
    y = (y | (y >> 2)) & 0x0f0f0f0fu;
    y = (y | (y >> 4)) & 0x00ff00ffu;
    y = (y | (y >> 8)) & 0x0000ffffu;
    return ((x >> 32) << 16) | y;
  }
  x &= sp.sig_mask;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint64_t t = x & sp.cmv[i];
    x = (x ^ t) | (t >> (1 << i));
  }
  return x;
}
SLK_HD uint64_t slk_expand(const slk_scan_params& sp, uint64_t x) {
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    uint64_t mv = sp.cmv[i];
    uint64_t t = x << (1 << i);
    x = (x & ~mv) | (t & mv);
  }
  return x & sp.sig_mask;
}

// ------------------------------------------------------------------------------------------------ hash table
SLK_HD uint64_t slk_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
SLK_HD uint64_t slk_bucket_of(uint64_t ckey, uint64_t n_buckets) {
  uint64_t h = ckey * 0x9E3779B97F4A7C15ull;
  h ^= h >> 29;
  return slk_mulhi64(h * 0xD6E8FEB86659FD93ull, n_buckets);
}
SLK_HD void slk_prefetch_bucket(const slk_table_view& tb, uint64_t ckey) {
#if defined(__CUDA_ARCH__)
  const uint64_t* p = tb.cells + slk_bucket_of(ckey, tb.n_buckets) * 4;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)tb; (void)ckey;
#endif
}

// One bucket (= one 32-byte sector) of a probe: both halves are loaded before anything is compared.
// Returns true when the probe is decided: *dense = the key's taxon, or 0 if an empty cell proves its absence
// (cells of a bucket fill in order and are never deleted, so nothing can follow an empty cell).
SLK_HD bool slk_probe_bucket(const slk_table_view& tb, uint64_t b, uint64_t ckey, uint32_t* dense) {
  uint64_t c0, c1, c2, c3;
#if defined(__CUDA_ARCH__)
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(tb.cells + b * 4);
  const ulonglong2 v0 = __ldg(p), v1 = __ldg(p + 1);
  c0 = v0.x; c1 = v0.y; c2 = v1.x; c3 = v1.y;
#else
  c0 = tb.cells[b * 4]; c1 = tb.cells[b * 4 + 1]; c2 = tb.cells[b * 4 + 2]; c3 = tb.cells[b * 4 + 3];
#endif
  const bool m0 = c0 != 0 && (c0 >> 16) == ckey, m1 = c1 != 0 && (c1 >> 16) == ckey;
  const bool m2 = c2 != 0 && (c2 >> 16) == ckey, m3 = c3 != 0 && (c3 >> 16) == ckey;
  const uint64_t hit = m0 ? c0 : m1 ? c1 : m2 ? c2 : m3 ? c3 : 0ull;
  *dense = (uint32_t)(hit & 0xffffu);
  return hit != 0 || c3 == 0;   // c3 == 0 <=> the bucket has an empty cell
}

// Probe: returns the dense taxon of the key, 0 when absent (a left join miss -> Taxonomy.NONE).
SLK_HD uint32_t slk_probe(const slk_table_view& tb, uint64_t ckey) {
  uint64_t b = slk_bucket_of(ckey, tb.n_buckets);
  for (uint64_t tries = 0; tries < tb.n_buckets; tries++) {
    uint32_t dense;
    if (slk_probe_bucket(tb, b, ckey, &dense)) return dense;
    b = (b + 1 == tb.n_buckets) ? 0 : b + 1;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ taxonomy
// LowestCommonAncestor.apply (slacken/LowestCommonAncestor.scala:49-78) on the dense, ancestor-closed taxonomy:
// the first node of b's path that lies on a's path = the deepest common node; disjoint paths give ROOT.
SLK_HD uint32_t slk_lca(const slk_tax_view& tx, uint32_t a, uint32_t b) {
  if (a == 0 || b == 0) return b == 0 ? a : b;
  while (a != b) {
    if (tx.depth[a] >= tx.depth[b]) a = tx.parent[a];
    else b = tx.parent[b];
  }
  return a ? a : tx.root;
}
// Taxonomy.hasAncestor (slacken/Taxonomy.scala:236-244); anc != 0
SLK_HD bool slk_has_ancestor(const slk_tax_view& tx, uint32_t t, uint32_t anc) {
  uint32_t da = tx.depth[anc];
  while (t != 0 && tx.depth[t] > da) t = tx.parent[t];
  return t != 0 && t == anc;
}

// ------------------------------------------------------------------------------------------------ scanner
SLK_HD uint64_t slk_min64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// Rolling m-mer (forward and reverse complement, both left-aligned), its priority, and the minimum over the
// last W priorities = the minimizer of the k-mer window that ends at the base just pushed (the monotone deque of
// PosRankWindow.scala:47-74 collapses to this for a fixed, small W; ties do not matter because super-mers merge
// on equal VALUE, MinSplitter.scala:200-201). The window minimum is kept as suffix minima: sfx[j] = min of the
// last j+1 priorities, updated in place, so no ring buffer has to be shifted.
// Branch-free by design: an invalid character only resets `nvalid`; stale bits in fwd/rc/sfx are flushed by the
// k valid bases that must follow before the next window is reported.
template <int W>
struct slk_scanner {
  uint64_t fwd, rc;
  uint64_t sfx[W > 1 ? W - 1 : 1];
  uint32_t nvalid;  // consecutive valid bases seen

  SLK_HD void reset() {
    fwd = 0; rc = 0; nvalid = 0;
#pragma unroll
    for (int i = 0; i < (W > 1 ? W - 1 : 1); i++) sfx[i] = 0;
  }
  // c: 0..3 for a base, 4 for anything else. Returns true when a full k-mer window of valid bases ends here.
  SLK_HD bool push(uint32_t c, uint32_t k, int fshift, uint64_t mmask, uint64_t xor_mask, uint64_t sig_mask,
                   bool canonical, uint64_t* minv) {
    uint32_t b = c & 3u;
    fwd = (fwd << 2) | ((uint64_t)b << fshift);
    rc = ((rc >> 2) | ((uint64_t)(3u - b) << 62)) & mmask;
    nvalid = c < 4u ? nvalid + 1 : 0;
    uint64_t x = canonical ? slk_min64(fwd, rc) : fwd;
    x = (x ^ xor_mask) & sig_mask;
    uint64_t mn = x;
    if (W > 1) {
      mn = slk_min64(x, sfx[W > 1 ? W - 2 : 0]);
#pragma unroll
      for (int j = W - 2; j >= 1; j--) sfx[j] = slk_min64(x, sfx[j - 1]);
      sfx[0] = x;
    }
    *minv = mn;
    return nvalid >= k;
  }
};

// One mate of a fragment, in either of the two input forms of the C ABI:
//  * ASCII: `ascii[0..len)`, one byte per base;
//  * packed: 32-base blocks, block b = bases 32b..32b+31 of the read; base i of a block sits in bits [2i, 2i+1] of
//    codes[b] (A=0 C=1 G=2 T/U=3) and bit i of mask[b] is set when that character is ambiguous (its code bits are 0).
//    Every read starts a new block, so a batch is three flat arrays plus block offsets and lengths.
struct slk_read_src {
  const uint8_t* ascii;
  const uint64_t* codes;
  const uint32_t* mask;
  uint32_t len;
};

// Span entries: one per super-mer (SEQ), per run of >= k ambiguous bases (AMB) and per mate border.
#define SLK_E_SEQ 0u
#define SLK_E_AMB 1u
#define SLK_E_BORDER 2u
#define SLK_E_CNT_MAX 0x3fffu   // longer runs are split, which no output can see (equal labels merge again)

struct slk_frag_result {
  int32_t taxon;          // raw taxon reported (0 when unclassified)
  uint32_t flags;         // SLK_F_*
  uint32_t kmers1, kmers2;  // sum of span k-mer counts per mate (lengthString = kmers + k - 1)
  uint32_t num_distinct;
  uint32_t n_hits;        // merged hits
  uint32_t n_probes;      // table probes issued (= SEQ spans)
};

// A sink for merged hits that no longer fit the per-thread buffers (device: a worst-case block of the global hit
// buffer; emulation: a vector). `need` = upper bound of the hits this fragment can still produce.
struct slk_null_sink {
  SLK_HD void push(int32_t, int32_t, uint32_t) {}
};

// Per-thread fast store: span entries key[j]/meta[j] (count | type << 14) for j < SLK_ECAP and the first SLK_SHITS
// merged hits. On the device these are columns of shared-memory tiles; the emulation uses plain arrays.
struct slk_store_local {
  uint64_t key[SLK_ECAP];
  uint16_t meta[SLK_ECAP];
  int32_t hl[SLK_SHITS], hc[SLK_SHITS];
  SLK_HD void set(uint32_t j, uint64_t k, uint32_t m) { key[j] = k; meta[j] = (uint16_t)m; }
  SLK_HD uint64_t get_key(uint32_t j) const { return key[j]; }
  SLK_HD uint32_t get_meta(uint32_t j) const { return meta[j]; }
  SLK_HD void set_hit(uint32_t i, int32_t label, int32_t count) { hl[i] = label; hc[i] = count; }
  SLK_HD void get_hit(uint32_t i, int32_t* label, int32_t* count) const { *label = hl[i]; *count = hc[i]; }
};

template <int W, class Sink, class Entries>
struct slk_frag_classifier {
  const slk_table_view tb;   // by value: the hot loops copy what they need into registers anyway
  const slk_tax_view tx;
  Sink& sink;
  Entries ent;

  // per-fragment state of the drain side
  uint64_t last_seq_key;
  bool have_last_seq;
  int32_t cur_label, cur_count;
  bool have_cur;
  uint32_t mate, kmers[2], nd, nprobes;
  uint32_t windows_left;  // upper bound of merged hits still to come (for the sink's spill allocation)
  // merged hits (dense labels): the first SLK_SHITS in the fast store, then xh_*, then spilled through the sink
  uint32_t nh;            // buffered
  uint32_t nh_spilled;    // already pushed to the sink (0 for all but very long reads)
  int32_t xh_label[SLK_XHITS];
  int32_t xh_count[SLK_XHITS];
  // histogram: dense taxon -> k-mer count, insertion ordered (fastutil Int2IntArrayMap)
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t nk;
  bool overflow;

  SLK_HD slk_frag_classifier(const slk_table_view& tb_, const slk_tax_view& tx_, Sink& s, const Entries& e)
      : tb(tb_), tx(tx_), sink(s), ent(e) {}

  SLK_HD void hist_add(uint32_t t, int32_t c) {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) { hv[i] += c; return; }
    if (nk == SLK_KMAX) { overflow = true; return; }
    hk[nk] = t; hv[nk] = c; nk++;
  }
  SLK_HD int32_t hist_get(uint32_t t) const {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) return hv[i];
    return 0;
  }
  SLK_HD void buffered_hit(uint32_t i, int32_t* label, int32_t* count) const {
    if (i < SLK_SHITS) ent.get_hit(i, label, count);
    else { *label = xh_label[i - SLK_SHITS]; *count = xh_count[i - SLK_SHITS]; }
  }
  // TaxonCounts.toMap (slacken/TaxonCounts.scala:70-81) over the buffered merged hits, in order
  SLK_HD void fold_hits() {
    for (uint32_t i = 0; i < nh; i++) {
      int32_t l, c;
      buffered_hit(i, &l, &c);
      if (l >= 0) hist_add((uint32_t)l, c);  // skips AMBIGUOUS / MATE_PAIR_BORDER
    }
  }
  // only reads with more than SLK_SHITS + SLK_XHITS merged hits come here
  SLK_HD_NOINLINE void spill(uint32_t need) {
    fold_hits();
    for (uint32_t i = 0; i < nh; i++) {
      int32_t l, c;
      buffered_hit(i, &l, &c);
      sink.push(l >= 0 ? tx.raw[l] : l, c, need + nh);
    }
    nh_spilled += nh;
    nh = 0;
  }
  // TaxonCounts.fromHits (slacken/TaxonCounts.scala:31-48) has already merged adjacent equal taxa: one store
  SLK_HD void push_hit(int32_t label, int32_t count, uint32_t need) {
    if (nh == SLK_SHITS + SLK_XHITS) spill(need);
    if (nh < SLK_SHITS) ent.set_hit(nh, label, count);
    else { xh_label[nh - SLK_SHITS] = label; xh_count[nh - SLK_SHITS] = count; }
    nh++;
  }

  // spanToHit (slacken/KeyValueIndex.scala:176-185) + numDistinct (slacken/Classifier.scala:94) for the first
  // `ne` buffered entries. Out of line on purpose (the scan loop stays small); every lane of a warp calls it at the
  // same time. Its running state is copied into registers for the duration of the call, because stores through
  // the entry/sink pointers could otherwise alias the members and force a reload per entry.
  SLK_HD_NOINLINE void drain(const slk_scan_params& sp, uint32_t ne) {
    const slk_table_view tb = this->tb;   // registers, not members behind `this`
#if defined(__CUDA_ARCH__)
    const Entries ent = this->ent;
#else
    const Entries& ent = this->ent;
#endif
    // all buckets of this batch are requested from HBM first, so the probes below find them in L2
    for (uint32_t j = 0; j < ne; j++)
      if (tb.prefetch && (ent.get_meta(j) >> 14) == SLK_E_SEQ) slk_prefetch_bucket(tb, slk_compress(sp, ent.get_key(j)));
    uint64_t l_last = last_seq_key;
    bool l_have_last = have_last_seq, l_have_cur = have_cur;
    int32_t l_label = cur_label, l_count = cur_count;
    uint32_t l_mate = mate, l_k0 = kmers[0], l_k1 = kmers[1], l_nd = nd, l_np = nprobes, l_wl = windows_left;
    const int32_t border_cnt = -(sp.k - 1);
    // One loop for "next entry" and "next bucket of the current probe": every lane walks its own entries and its
    // own collision chains at its own pace, so a long chain in one lane does not hold back the other 31.
    uint32_t j = 0;
    bool probing = false;
    uint64_t key = 0, ckey = 0, bucket = 0;
    while (probing || j < ne) {
      const uint32_t meta = ent.get_meta(j), type = meta >> 14, cnt = meta & SLK_E_CNT_MAX;
      int32_t label, hcnt = (int32_t)cnt;
      if (type == SLK_E_SEQ) {
        if (!probing) {
          key = ent.get_key(j);
          ckey = slk_compress(sp, key);
          bucket = slk_bucket_of(ckey, tb.n_buckets);
          probing = true;
          l_np++;
        }
        uint32_t dense;
        if (!slk_probe_bucket(tb, bucket, ckey, &dense)) {
          bucket = (bucket + 1 == tb.n_buckets) ? 0 : bucket + 1;
          continue;
        }
        probing = false;
        l_nd += ((!l_have_last || key != l_last) && dense != 0) ? 1u : 0u;
        l_last = key; l_have_last = true;
        label = (int32_t)dense;
      } else if (type == SLK_E_AMB) {
        label = SLK_AMBIGUOUS_SPAN;
      } else {
        label = SLK_MATE_PAIR_BORDER; hcnt = border_cnt;
      }
      if (type != SLK_E_BORDER) { if (l_mate) l_k1 += cnt; else l_k0 += cnt; }
      // TaxonCounts.fromHits: adjacent hits with the same taxon merge (slacken/TaxonCounts.scala:31-48)
      if (l_have_cur && label == l_label) l_count += hcnt;
      else {
        if (l_have_cur) push_hit(l_label, l_count, l_wl + 2);
        l_label = label; l_count = hcnt; l_have_cur = true;
      }
      if (type == SLK_E_BORDER) l_mate = 1;
      l_wl = l_wl > cnt ? l_wl - cnt : 0;
      j++;
    }
    last_seq_key = l_last; have_last_seq = l_have_last; have_cur = l_have_cur; cur_label = l_label; cur_count = l_count;
    mate = l_mate; kmers[0] = l_k0; kmers[1] = l_k1; nd = l_nd; nprobes = l_np; windows_left = l_wl;
  }

  // LowestCommonAncestor.resolveTree (slacken/LowestCommonAncestor.scala:91-146)
  SLK_HD_NOINLINE uint32_t resolve(double confidence) {
    fold_hits();   // the buffered hits (a fragment that spilled has none left and was folded on the way)
    int32_t total = (int32_t)(kmers[0] + kmers[1]);  // totalKmers: ambiguous spans count, the border does not
    double required = ceil(confidence * (double)total);
    uint32_t max_taxon = 0;
    int32_t max_score = 0;
    // one hit taxon (plus, possibly, misses): its path score is its own count, nothing to walk
    uint32_t nz = 0, only = 0;
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] != 0) { nz++; only = hk[i]; }
    if (nz == 1) {
      max_taxon = only;
    } else if (nz > 1) {
      for (uint32_t i = 0; i < nk; i++) {
        uint32_t taxon = hk[i], node = taxon;
        int32_t score = 0;
        while (node != 0) { score += hist_get(node); node = tx.parent[node]; }
        if (score > max_score) { max_taxon = taxon; max_score = score; }
        else if (score == max_score) max_taxon = slk_lca(tx, max_taxon, taxon);
      }
    }
    max_score = hist_get(max_taxon);
    while (max_taxon != 0 && (double)max_score < required) {
      max_score = 0;
      for (uint32_t i = 0; i < nk; i++)
        if (slk_has_ancestor(tx, hk[i], max_taxon)) max_score += hv[i];
      if ((double)max_score >= required) return max_taxon;
      max_taxon = tx.parent[max_taxon];
    }
    return max_taxon;
  }

  // One fragment end to end (paired == false: r2 is ignored).
  // Scan: Supermers.splitByAmbiguity/splitFragment (slacken/Supermers.scala:113-189): valid runs >= k are cut into
  // super-mers (runs of k-mer windows with equal minimizer), runs of >= k ambiguous characters become one
  // AMBIGUOUS span of len-(k-1), anything shorter vanishes; the mates are separated by a MATE_PAIR_BORDER span.
  template <bool PACKED>
  SLK_HD void run(const slk_scan_params& sp, const slk_read_src& r1, const slk_read_src& r2, bool paired,
                  double confidence, int32_t min_hit_groups, slk_frag_result& r) {
    have_last_seq = false; last_seq_key = 0; have_cur = false; cur_label = 0; cur_count = 0; nh = 0; nh_spilled = 0;
    mate = 0; kmers[0] = 0; kmers[1] = 0; nd = 0; nk = 0; overflow = false; nprobes = 0;
    const uint32_t k = (uint32_t)sp.k, km1 = k - 1;
    const int fshift = sp.fshift;
    const uint64_t mmask = sp.mmask, xor_mask = sp.xor_mask, sig_mask = sp.sig_mask;
    const bool canonical = sp.canonical != 0;
    windows_left = (r1.len > km1 ? r1.len - km1 : 0) + (paired ? (r2.len > km1 ? r2.len - km1 : 0) + 1 : 0);
    uint32_t ne = 0;       // buffered entries (register)
    bool any = false;      // did the fragment yield any span at all
#if defined(__CUDA_ARCH__)
    const Entries ent = this->ent;   // the three shared-space addresses, in registers
#define SLK_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define SLK_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#else
    Entries& ent = this->ent;
#define SLK_WARP_ANY(p) (p)
#define SLK_WARP_MAX(x) (x)
#endif
#pragma unroll 1
    for (int mt = 0; mt < (paired ? 2 : 1); mt++) {  // one copy of the scan loop serves both mates
      if (mt) {
        if (SLK_WARP_ANY(ne > SLK_ECAP - 5)) { any = any || ne != 0; drain(sp, ne); ne = 0; }
        ent.set(ne, 0, SLK_E_BORDER << 14); ne++;
      }
      const slk_read_src& src = mt ? r2 : r1;
      const uint32_t len = src.len;
      slk_scanner<W> sc;
      sc.reset();
      uint64_t run_key = 0;
      uint32_t run_cnt = 0, ninv = 0, amb_cnt = 0;
      bool in_run = false;
      // One character (c = 0..3 for a base, 4 for anything else). Straight-line code: at most one entry is stored.
      auto step = [&](uint32_t c) {
        const bool valid = c < 4u;
        uint64_t mn;
        const bool window_ok = sc.push(c, k, fshift, mmask, xor_mask, sig_mask, canonical, &mn);
        ninv = valid ? 0u : ninv + 1u;
        const bool same = in_run && mn == run_key && run_cnt < SLK_E_CNT_MAX;
        const bool start_new = window_ok && !same;
        const bool emit_seq = in_run && (start_new || !valid);  // the open super-mer ends here
        const bool emit_amb = amb_cnt != 0 && (valid || amb_cnt == SLK_E_CNT_MAX);  // an ambiguous stretch ended
        if (emit_seq || emit_amb) {
          ent.set(ne, emit_seq ? run_key : 0ull, emit_seq ? run_cnt : (amb_cnt | (SLK_E_AMB << 14)));
          ne++;
        }
        amb_cnt = (valid || emit_amb) ? 0u : amb_cnt;
        amb_cnt += (!valid && ninv >= k) ? 1u : 0u;
        run_cnt = start_new ? 1u : run_cnt + ((window_ok && same) ? 1u : 0u);
        run_key = start_new ? mn : run_key;
        in_run = valid && (in_run || start_new);
      };
      // All lanes of a warp run the same number of iterations (the longest read of the warp decides) and drain
      // together as soon as one lane's entry tile is nearly full, so the warp never splits around drain().
      if (PACKED) {
        const uint32_t nblk = (len + 31u) >> 5;
        const uint32_t nblk_w = SLK_WARP_MAX(nblk);
        for (uint32_t b = 0; b < nblk_w; b++) {
          uint64_t cw = 0;
          uint32_t mw = 0, nb = 0;
          if (b < nblk) {
#if defined(__CUDA_ARCH__)
            cw = __ldg(src.codes + b); mw = __ldg(src.mask + b);
#else
            cw = src.codes[b]; mw = src.mask[b];
#endif
            nb = len - 32u * b; nb = nb > 32u ? 32u : nb;
          }
#pragma unroll 1
          for (uint32_t q = 0; q < 8; q++) {
            if (SLK_WARP_ANY(ne > SLK_ECAP - 5)) { any = any || ne != 0; drain(sp, ne); ne = 0; }  // 4 chars: <= 4 entries
            const uint32_t g = (uint32_t)(cw >> (8u * q)) & 0xffu, gm = (mw >> (4u * q)) & 0xfu;
            const uint32_t left = nb > 4u * q ? nb - 4u * q : 0u;
            if (left >= 4u) {
#pragma unroll
              for (uint32_t i = 0; i < 4; i++) step(((g >> (2u * i)) & 3u) | (((gm >> i) & 1u) << 2));
            } else {
#pragma unroll 1
              for (uint32_t i = 0; i < left; i++) step(((g >> (2u * i)) & 3u) | (((gm >> i) & 1u) << 2));
            }
          }
        }
      } else {
#if defined(__CUDA_ARCH__)
        // 16-byte aligned vector loads over [s, s+len). The buffer is readable up to the next 16-byte boundary
        // (library-owned and cudaMalloc'ed buffers are); bytes outside the read are skipped in the edge chunks.
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(src.ascii);
        const uintptr_t abase = a0 & ~(uintptr_t)15;
        const int32_t lo0 = (int32_t)(a0 - abase), total = lo0 + (int32_t)len;  // byte range [lo0, total) from abase
        const int32_t total_w = (int32_t)SLK_WARP_MAX((uint32_t)total);
        for (int32_t cb = 0; cb < total_w; cb += 16) {
          const bool have = cb < total;
          uint4 v = make_uint4(0u, 0u, 0u, 0u);
          if (have) v = __ldg(reinterpret_cast<const uint4*>(abase + cb));
          const bool interior = cb >= lo0 && cb + 16 <= total;
#pragma unroll 1
          for (int wi = 0; wi < 4; wi++) {
            if (SLK_WARP_ANY(ne > SLK_ECAP - 5)) { any = any || ne != 0; drain(sp, ne); ne = 0; }
            const uint32_t word = wi == 0 ? v.x : wi == 1 ? v.y : wi == 2 ? v.z : v.w;
            if (interior) {
#pragma unroll
              for (int bi = 0; bi < 4; bi++) step(slk_code((word >> (8 * bi)) & 0xffu));
            } else if (have) {
#pragma unroll 1
              for (int bi = 0; bi < 4; bi++) {
                const int32_t pos = cb + 4 * wi + bi;
                if (pos >= lo0 && pos < total) step(slk_code((word >> (8 * bi)) & 0xffu));
              }
            }
          }
        }
#else
        for (uint32_t i = 0; i < len; i++) {
          if ((i & 3) == 0 && ne > SLK_ECAP - 5) { any = true; drain(sp, ne); ne = 0; }
          step(slk_code(src.ascii[i]));
        }
#endif
      }
      // mate end: at most one pending entry (a run and an ambiguous stretch cannot both be open)
      if (in_run) { ent.set(ne, run_key, run_cnt); ne++; }
      if (amb_cnt) { ent.set(ne, 0, amb_cnt | (SLK_E_AMB << 14)); ne++; }
    }
    if (ne) any = true;
    drain(sp, ne);
    if (have_cur) { push_hit(cur_label, cur_count, 2); have_cur = false; }
    if (nh_spilled) spill(0);   // a fragment that went to the sink keeps all its hits there
    uint32_t taxon = resolve(confidence);
    bool classified = taxon != 0 && nd >= (uint32_t)min_hit_groups;  // slacken/Classifier.scala:446
    r.taxon = classified ? tx.raw[taxon] : 0;
    r.flags = (classified ? SLK_F_CLASSIFIED : 0u) | (any ? SLK_F_HAS_SPAN : 0u) | (overflow ? SLK_F_OVERFLOW : 0u);
    r.kmers1 = kmers[0]; r.kmers2 = kmers[1];
    r.num_distinct = nd; r.n_hits = nh + nh_spilled; r.n_probes = nprobes;
#undef SLK_WARP_ANY
#undef SLK_WARP_MAX
  }
};

// K1 as a stand-alone step: ASCII -> 2-bit codes + ambiguity mask in the packed block layout of slk_read_src.
// `emit(block index, codes, mask)` is called for every 32-base block of the read.
template <class Emit>
SLK_HD void slk_pack_read(const uint8_t* s, uint32_t len, Emit&& emit) {
  uint64_t cw = 0;
  uint32_t mw = 0, i = 0, b = 0;
  slk_for_each_byte(s, len, [&](uint32_t ch) {
    const uint32_t c = slk_code(ch);
    cw |= (uint64_t)(c & 3u) << (2u * i);
    mw |= (c >> 2) << i;
    if (++i == 32u) { emit(b, cw, mw); b++; cw = 0; mw = 0; i = 0; }
  });
  if (i) emit(b, cw, mw);
}

// ------------------------------------------------------------------------------------------------ build side
// SplitterMinimizers.find (slacken/Minimizers.scala:43-76): every super-mer of a genome fragment contributes
// (minimizer, taxon). One thread scans the k-mer windows [w0, w0+nw) of one fragment (bases s[w0 .. w0+nw+k-1))
// and emits one cell (compressed key << 16 | dense taxon) per run of equal window minima. Runs cut at thread
// borders only duplicate a cell, which the sort + LCA reduce absorbs. Invalid characters break the sequence
// exactly like InputReader.removeInvalid (kmers/input/InputReader.scala:60-72) does on the host.
template <int W, class Emit>
SLK_HD void slk_emit_cells(const slk_scan_params& sp, const uint8_t* s, uint64_t nbases, uint32_t dense_taxon,
                           Emit& out) {
  slk_scanner<W> sc;
  sc.reset();
  const uint32_t k = (uint32_t)sp.k;
  const int fshift = sp.fshift;
  const uint64_t mmask = sp.mmask, xor_mask = sp.xor_mask, sig_mask = sp.sig_mask;
  const bool canonical = sp.canonical != 0;
  uint64_t run_key = 0;
  bool in_run = false;
  slk_for_each_byte(s, nbases, [&](uint32_t ch) {
    const uint32_t c = slk_code(ch);
    uint64_t mn;
    const bool window_ok = sc.push(c, k, fshift, mmask, xor_mask, sig_mask, canonical, &mn);
    if (window_ok && (!in_run || mn != run_key)) {
      out((slk_compress(sp, mn) << 16) | dense_taxon);
      run_key = mn;
    }
    in_run = window_ok;
  });
}

// ------------------------------------------------------------------------------------------------ synthetic data
// Counter-based generators (not part of the reference): every byte is a pure function of (seed, index), so the
// CUDA generator can be compared byte for byte with the oracle's own independent copy (oracle/slk_oracle.c).
SLK_HD uint64_t slk_splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
SLK_HD uint64_t slk_rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
  return slk_splitmix64(slk_splitmix64(seed * 0x100000001b3ull + stream) ^ idx);
}
SLK_HD uint8_t slk_acgt(uint32_t c) { return c == 0 ? 'A' : c == 1 ? 'C' : c == 2 ? 'G' : 'T'; }
SLK_HD uint8_t slk_synth_genome_base(uint64_t seed, uint64_t g) {
  uint64_t r = slk_rnd(seed, 2, g >> 16);
  uint64_t st = r & 0xffff, ln = 1 + ((r >> 16) % 100), o = g & 0xffff;
  if (o >= st && o < st + ln) return 'N';
  uint64_t w = slk_rnd(seed, 1, g >> 5);
  return slk_acgt((uint32_t)((w >> (2 * (g & 31))) & 3));
}
SLK_HD uint8_t slk_comp_char(uint8_t c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c; }
SLK_HD uint8_t slk_synth_read_base(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                                   uint64_t r, uint32_t L, uint32_t j) {
  uint64_t nn = slk_rnd(rseed, 14, r);
  if ((nn % 200) == 0 && (nn >> 8) % L == j) return 'N';
  uint64_t h = slk_rnd(rseed, 10, r);
  if ((h % 10) < 8) {
    uint64_t g = (h >> 8) % n_genomes;
    uint64_t pos = slk_rnd(rseed, 11, r) % (genome_len - L + 1);
    uint64_t base = g * genome_len + pos;
    uint8_t c = ((h >> 40) & 1) ? slk_comp_char(slk_synth_genome_base(gseed, base + (L - 1 - j)))
                                : slk_synth_genome_base(gseed, base + j);
    uint64_t e = slk_rnd(rseed, 12, r * 1024 + j);
    if ((e % 100) == 0 && c != 'N') c = slk_acgt((uint32_t)((e >> 8) & 3));
    return c;
  }
  uint64_t w = slk_rnd(rseed, 13, r * 32 + (j >> 5));
  return slk_acgt((uint32_t)((w >> (2 * (j & 31))) & 3));
}
