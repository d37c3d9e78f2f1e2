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
#else
#define SLK_HD inline
#endif

#define SLK_MAX_W 8      // k - m + 1 supported by the kernels (35 - 31 + 1 = 5 for the Kraken 2 defaults)
#define SLK_ECAP 64      // span entries buffered per thread between scan and probe
#define SLK_HCAP 64      // merged hits buffered per thread before spilling to a worst-case global block
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
  uint32_t u = c | 0x20u;
  uint32_t code = (c >> 1) & 3u;
  code ^= code >> 1;
  bool ok = (u == 'a') | (u == 'c') | (u == 'g') | (u == 't') | (u == 'u');
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
  h ^= h >> 32;
  h *= 0xD6E8FEB86659FD93ull;
  h ^= h >> 32;
  return slk_mulhi64(h, n_buckets);
}

// Probe: returns the dense taxon of the key, 0 when absent (a left join miss -> Taxonomy.NONE).
SLK_HD uint32_t slk_probe(const slk_table_view& tb, uint64_t ckey) {
  uint64_t b = slk_bucket_of(ckey, tb.n_buckets);
  for (uint64_t tries = 0; tries < tb.n_buckets; tries++) {
    uint64_t c[4];
#if defined(__CUDA_ARCH__)
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(tb.cells + b * 4);
    ulonglong2 v0 = __ldg(p), v1 = __ldg(p + 1);
    c[0] = v0.x; c[1] = v0.y; c[2] = v1.x; c[3] = v1.y;
#else
    for (int i = 0; i < 4; i++) c[i] = tb.cells[b * 4 + i];
#endif
#pragma unroll
    for (int i = 0; i < 4; i++) {
      if (c[i] == 0) return 0;
      if ((c[i] >> 16) == ckey) return (uint32_t)(c[i] & 0xffffu);
    }
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
// Rolling m-mer (forward and reverse complement, both left-aligned), its priority, and the minimum over the
// last W priorities = the minimizer of the k-mer window that ends at the base just pushed.
template <int W>
struct slk_scanner {
  uint64_t fwd, rc;
  uint64_t ring[W];
  uint32_t nvalid;  // consecutive valid bases seen

  SLK_HD void reset() {
    fwd = 0; rc = 0; nvalid = 0;
#pragma unroll
    for (int i = 0; i < W; i++) ring[i] = 0;
  }
  // push one valid base (0..3); true when a full k-mer window ends here, *minv = its minimizer priority
  SLK_HD bool push(const slk_scan_params& sp, uint32_t c, uint64_t* minv) {
    fwd = (fwd << 2) | ((uint64_t)c << sp.fshift);
    rc = ((rc >> 2) | ((uint64_t)(3u - c) << 62)) & sp.mmask;
    nvalid++;
    uint64_t x = (sp.canonical && rc < fwd) ? rc : fwd;
    x = (x ^ sp.xor_mask) & sp.sig_mask;
#pragma unroll
    for (int i = 0; i + 1 < W; i++) ring[i] = ring[i + 1];
    ring[W - 1] = x;
    if (nvalid < (uint32_t)sp.k) return false;
    uint64_t mn = ring[0];
#pragma unroll
    for (int i = 1; i < W; i++) mn = ring[i] < mn ? ring[i] : mn;
    *minv = mn;
    return true;
  }
};

// Span entries: one per super-mer (SEQ), per run of >= k ambiguous bases (AMB) and per mate border.
#define SLK_E_SEQ 0u
#define SLK_E_AMB 1u
#define SLK_E_BORDER 2u

struct slk_frag_result {
  int32_t taxon;          // raw taxon reported (0 when unclassified)
  uint32_t flags;         // SLK_F_*
  uint32_t kmers1, kmers2;  // sum of span k-mer counts per mate (lengthString = kmers + k - 1)
  uint32_t num_distinct;
  uint32_t n_hits;        // merged hits
  uint32_t n_probes;      // table probes issued (= SEQ spans)
};

// A sink for merged hits. The device kernel's sink buffers SLK_HCAP hits per thread and spills to a global block.
struct slk_null_sink {
  SLK_HD void push(int32_t, int32_t, uint32_t) {}
};

template <int W, class Sink>
struct slk_frag_classifier {
  const slk_scan_params& sp;
  const slk_table_view& tb;
  const slk_tax_view& tx;
  Sink& sink;

  // span entries waiting for their probe
  uint64_t ekey[SLK_ECAP];
  uint32_t emeta[SLK_ECAP];  // count | type << 30
  uint32_t ne;
  // per-fragment state
  uint64_t last_seq_key;
  bool have_last_seq;
  int32_t cur_label, cur_count;
  bool have_cur;
  uint32_t mate, kmers[2], nd, nhits, total_entries, nprobes;
  uint32_t windows_left;  // upper bound of merged hits still to come (for the sink's spill allocation)
  // histogram: dense taxon -> k-mer count, insertion ordered (fastutil Int2IntArrayMap)
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t nk;
  bool overflow;

  SLK_HD slk_frag_classifier(const slk_scan_params& sp_, const slk_table_view& tb_, const slk_tax_view& tx_, Sink& s)
      : sp(sp_), tb(tb_), tx(tx_), sink(s) {}

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

  // TaxonCounts.fromHits: adjacent hits with the same taxon merge (slacken/TaxonCounts.scala:31-48)
  SLK_HD void flush_hit() {
    if (!have_cur) return;
    int32_t out_taxon = cur_label >= 0 ? tx.raw[cur_label] : cur_label;
    sink.push(out_taxon, cur_count, windows_left);
    nhits++;
    if (cur_label >= 0) hist_add((uint32_t)cur_label, cur_count);  // toMap skips AMBIGUOUS / MATE_PAIR_BORDER
    have_cur = false;
  }
  SLK_HD void add_hit(int32_t label, int32_t count) {
    if (have_cur && label == cur_label) { cur_count += count; return; }
    flush_hit();
    cur_label = label; cur_count = count; have_cur = true;
  }

  // spanToHit (slacken/KeyValueIndex.scala:176-185) + numDistinct (slacken/Classifier.scala:94)
  SLK_HD void drain() {
    for (uint32_t j = 0; j < ne; j++) {
      uint32_t type = emeta[j] >> 30, cnt = emeta[j] & 0x3fffffffu;
      if (type == SLK_E_SEQ) {
        uint64_t key = ekey[j];
        uint32_t dense = slk_probe(tb, slk_compress(sp, key));
        nprobes++;
        bool distinct = !have_last_seq || key != last_seq_key;
        if (distinct && dense != 0) nd++;
        last_seq_key = key; have_last_seq = true;
        kmers[mate] += cnt;
        add_hit((int32_t)dense, (int32_t)cnt);
      } else if (type == SLK_E_AMB) {
        kmers[mate] += cnt;
        add_hit(SLK_AMBIGUOUS_SPAN, (int32_t)cnt);
      } else {
        add_hit(SLK_MATE_PAIR_BORDER, -(sp.k - 1));
        mate = 1;
      }
      windows_left = windows_left > cnt ? windows_left - cnt : 0;
    }
    ne = 0;
  }
  SLK_HD void emit(uint32_t type, uint64_t key, uint32_t cnt) {
    if (ne == SLK_ECAP) drain();
    ekey[ne] = key; emeta[ne] = cnt | (type << 30); ne++;
    total_entries++;
  }

  struct mate_state {
    slk_scanner<W> sc;
    uint64_t run_key;
    uint32_t run_cnt, ninv, amb_cnt;
    bool in_run;
  };
  SLK_HD void mate_begin(mate_state& st) {
    st.sc.reset(); st.run_key = 0; st.run_cnt = 0; st.ninv = 0; st.amb_cnt = 0; st.in_run = false;
  }
  // one character of a read. Supermers.splitByAmbiguity/splitFragment: valid runs >= k are scanned for
  // super-mers, runs of >= k ambiguous characters become one AMBIGUOUS span of len-(k-1), anything shorter vanishes.
  SLK_HD void step(mate_state& st, uint32_t ch) {
    uint32_t c = slk_code(ch);
    if (c < 4u) {
      if (st.amb_cnt) { emit(SLK_E_AMB, 0, st.amb_cnt); st.amb_cnt = 0; }
      st.ninv = 0;
      uint64_t mn;
      if (st.sc.push(sp, c, &mn)) {
        if (st.in_run && mn == st.run_key && st.run_cnt < 0x3fffffffu) st.run_cnt++;
        else {
          if (st.in_run) emit(SLK_E_SEQ, st.run_key, st.run_cnt);
          st.run_key = mn; st.run_cnt = 1; st.in_run = true;
        }
      }
    } else {
      if (st.in_run) { emit(SLK_E_SEQ, st.run_key, st.run_cnt); st.in_run = false; }
      st.sc.nvalid = 0;
      st.ninv++;
      if (st.ninv >= (uint32_t)sp.k) st.amb_cnt++;
    }
  }
  SLK_HD void mate_end(mate_state& st) {
    if (st.in_run) { emit(SLK_E_SEQ, st.run_key, st.run_cnt); st.in_run = false; }
    if (st.amb_cnt) { emit(SLK_E_AMB, 0, st.amb_cnt); st.amb_cnt = 0; }
  }

  SLK_HD void scan_mate(const uint8_t* s, uint32_t len) {
    mate_state st;
    mate_begin(st);
    slk_for_each_byte(s, len, [&](uint32_t ch) { step(st, ch); });
    mate_end(st);
  }

  // LowestCommonAncestor.resolveTree (slacken/LowestCommonAncestor.scala:91-146)
  SLK_HD uint32_t resolve(double confidence) {
    int32_t total = (int32_t)(kmers[0] + kmers[1]);  // totalKmers: ambiguous spans count, the border does not
    double required = ceil(confidence * (double)total);
    uint32_t max_taxon = 0;
    int32_t max_score = 0;
    for (uint32_t i = 0; i < nk; i++) {
      uint32_t taxon = hk[i], node = taxon;
      int32_t score = 0;
      while (node != 0) { score += hist_get(node); node = tx.parent[node]; }
      if (score > max_score) { max_taxon = taxon; max_score = score; }
      else if (score == max_score) max_taxon = slk_lca(tx, max_taxon, taxon);
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

  // One fragment end to end. s2 == nullptr for single-end reads.
  SLK_HD void run(const uint8_t* s1, uint32_t len1, const uint8_t* s2, uint32_t len2, double confidence,
                  int32_t min_hit_groups, slk_frag_result& r) {
    ne = 0; have_last_seq = false; last_seq_key = 0; have_cur = false; cur_label = 0; cur_count = 0;
    mate = 0; kmers[0] = 0; kmers[1] = 0; nd = 0; nhits = 0; total_entries = 0; nk = 0; overflow = false; nprobes = 0;
    uint32_t km1 = (uint32_t)(sp.k - 1);
    windows_left = (len1 > km1 ? len1 - km1 : 0) + (s2 ? (len2 > km1 ? len2 - km1 : 0) + 1 : 0);
    scan_mate(s1, len1);
    if (s2) {
      emit(SLK_E_BORDER, 0, 0);
      scan_mate(s2, len2);
    }
    drain();
    flush_hit();
    uint32_t taxon = resolve(confidence);
    bool classified = taxon != 0 && nd >= (uint32_t)min_hit_groups;  // slacken/Classifier.scala:446
    r.taxon = classified ? tx.raw[taxon] : 0;
    r.flags = (classified ? SLK_F_CLASSIFIED : 0u) | (total_entries ? SLK_F_HAS_SPAN : 0u) | (overflow ? SLK_F_OVERFLOW : 0u);
    r.kmers1 = kmers[0]; r.kmers2 = kmers[1];
    r.num_distinct = nd; r.n_hits = nhits; r.n_probes = nprobes;
  }
};

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
  uint64_t run_key = 0;
  bool in_run = false;
  slk_for_each_byte(s, nbases, [&](uint32_t ch) {
    uint32_t c = slk_code(ch);
    if (c < 4u) {
      uint64_t mn;
      if (sc.push(sp, c, &mn)) {
        if (!in_run || mn != run_key) {
          out((slk_compress(sp, mn) << 16) | dense_taxon);
          run_key = mn; in_run = true;
        }
      }
    } else {
      sc.nvalid = 0; in_run = false;
    }
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
