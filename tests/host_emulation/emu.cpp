// emu.cpp -- TEST-ONLY host emulation of the kernel bodies in slacken_b200/csrc/slk_core.h.
// It runs the per-thread functions of the CUDA kernels one "thread" at a time on the CPU so that the device
// logic can be checked against the oracle on machines without a GPU (pytest -m "not gpu").
// It is NOT a fallback: nothing in slacken_b200/ loads it, and the product library refuses to run without CUDA.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

// a small tile, so that the tests close tiles (and move histograms out of the fast store) all the time
#define SLK_POOL 8
#include "../../slacken_b200/csrc/slk_core.h"

#define EMU_API extern "C" __attribute__((visibility("default")))

EMU_API int emu_scan_params(int k, int m, int spaces, uint64_t mask, int canonical, slk_scan_params* sp) {
  memset(sp, 0, sizeof(*sp));
  return slk_make_scan_params(k, m, spaces, mask, canonical, sp);
}
EMU_API int emu_sizeof_scan_params() { return (int)sizeof(slk_scan_params); }
EMU_API uint64_t emu_compress(const slk_scan_params* sp, uint64_t x) { return slk_compress(*sp, x); }
EMU_API uint64_t emu_expand(const slk_scan_params* sp, uint64_t x) { return slk_expand(*sp, x); }
EMU_API uint32_t emu_code(uint32_t c) { return slk_code(c); }
// number of words among words[0..n) whose slk_code4 differs from four slk_code calls
EMU_API uint64_t emu_code4_mismatches(const uint32_t* words, uint64_t n) {
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; i++) {
    uint32_t c8, i4, rc = 0, ri = 0;
    slk_code4(words[i], &c8, &i4);
    for (int j = 0; j < 4; j++) {
      const uint32_t c = slk_code((words[i] >> (8 * j)) & 0xffu);
      rc |= (c & 3u) << (2 * j);
      ri |= (c >> 2) << j;
    }
    bad += (c8 != rc) || (i4 != ri);
  }
  return bad;
}
EMU_API uint64_t emu_buckets_for(uint64_t n_keys) { return (((uint64_t)((double)n_keys / 0.70) + 64 + 15) / 16) * 4; }

// sequential twin of insert_cells_kernel (no atomics needed with one "thread")
// world > 1: the table is one shard of a library cut for `world` ranks (slk_table_view::mix_mul)
EMU_API void emu_insert_cells(uint64_t* cells, uint64_t n_buckets, const uint16_t* parent, const uint8_t* depth,
                              uint32_t root, const uint64_t* in, uint64_t n, uint32_t world) {
  slk_tax_view tx{parent, depth, nullptr, 0, root};
  for (uint64_t i = 0; i < n; i++) {
    uint64_t cell = in[i], ckey = cell >> 16;
    uint32_t taxon = (uint32_t)(cell & 0xffff);
    if (!taxon) continue;
    const slk_table_view tbv{cells, n_buckets, world ? world : 1u, 0};
    uint64_t b = slk_bucket_of(ckey, tbv);
    bool done = false;
    for (uint64_t tries = 1; !done; tries++) {
      for (int j = 0; j < 4 && !done; j++) {
        uint64_t& c = cells[b * 4 + j];
        if (c == 0) { c = cell; done = true; }
        else if ((c >> 16) == ckey) { c = (ckey << 16) | slk_lca(tx, (uint32_t)(c & 0xffff), taxon); done = true; }
      }
      b = slk_next_bucket(b, tries, n_buckets);
    }
  }
}

struct vec_sink {
  std::vector<slk_hit>* v;
  void push(int32_t taxon, int32_t count, uint32_t) { v->push_back(slk_hit{taxon, count}); }
  void reserve(uint32_t) {}
};

struct emu_result { int32_t taxon; uint32_t flags, kmers1, kmers2, num_distinct, n_hits; };

// one classify "thread" per read; hits are appended to hits_out (cap entries), hit_off[i] = first hit of read i
template <int W>
static int64_t classify_w(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent,
                          const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* b1,
                          const uint64_t* o1, const uint8_t* b2, const uint64_t* o2, uint32_t n, double confidence,
                          int min_hit_groups, emu_result* res, uint64_t* hit_off, slk_hit* hits_out, uint64_t cap, bool packed) {
  slk_table_view tb{cells, n_buckets, 1, 0};
  slk_tax_view tx{parent, depth, raw, n_dense, root};
  std::vector<slk_hit> hv;
  uint64_t used = 0;
  for (uint32_t r = 0; r < n; r++) {
    hv.clear();
    vec_sink sink{&hv};
    slk_store_local ent;
    slk_frag_classifier<W, vec_sink, slk_store_local> cl(tb, tx, sink, ent);
    slk_frag_result fr;
    slk_read_src r1{b1 + o1[r], nullptr, nullptr, (uint32_t)(o1[r + 1] - o1[r])};
    slk_read_src r2{b2 ? b2 + o2[r] : nullptr, nullptr, nullptr, b2 ? (uint32_t)(o2[r + 1] - o2[r]) : 0u};
    std::vector<uint64_t> c1, c2;
    std::vector<uint32_t> m1, m2;
    if (packed) {   // the same read through the packed input form: pack it with the K1 body first
      slk_pack_read(r1.ascii, r1.len, [&](uint32_t, uint64_t cw, uint32_t mw) { c1.push_back(cw); m1.push_back(mw); });
      r1.codes = c1.data(); r1.mask = m1.data();
      if (b2) {
        slk_pack_read(r2.ascii, r2.len, [&](uint32_t, uint64_t cw, uint32_t mw) { c2.push_back(cw); m2.push_back(mw); });
        r2.codes = c2.data(); r2.mask = m2.data();
      }
      if (sp->canonical) cl.template run<true, true>(*sp, r1, r2, b2 != nullptr, confidence, min_hit_groups, fr);
      else cl.template run<true, false>(*sp, r1, r2, b2 != nullptr, confidence, min_hit_groups, fr);
    } else {
      if (sp->canonical) cl.template run<false, true>(*sp, r1, r2, b2 != nullptr, confidence, min_hit_groups, fr);
      else cl.template run<false, false>(*sp, r1, r2, b2 != nullptr, confidence, min_hit_groups, fr);
    }
    res[r] = emu_result{fr.taxon, fr.flags, fr.kmers1, fr.kmers2, fr.num_distinct, fr.n_hits};
    if (cl.nh_spilled == 0)   // the kernel epilogue: buffered hits leave with raw taxon ids
      for (uint32_t i = 0; i < cl.nh; i++) {
        int32_t l, c;
        cl.buffered_hit(i, &l, &c);
        hv.push_back(slk_hit{l >= 0 ? raw[l] : l, c});
      }
    hit_off[r] = used;
    if (used + hv.size() > cap) return -1;
    std::copy(hv.begin(), hv.end(), hits_out + used);
    used += hv.size();
  }
  return (int64_t)used;
}

#define DISPATCH_W(w, CALL)                                                                                       \
  switch (w) {                                                                                                    \
    case 1: { constexpr int W_ = 1; CALL; break; } case 2: { constexpr int W_ = 2; CALL; break; }                 \
    case 3: { constexpr int W_ = 3; CALL; break; } case 4: { constexpr int W_ = 4; CALL; break; }                 \
    case 5: { constexpr int W_ = 5; CALL; break; } case 6: { constexpr int W_ = 6; CALL; break; }                 \
    case 7: { constexpr int W_ = 7; CALL; break; } default: { constexpr int W_ = 8; CALL; break; }                \
  }

EMU_API int64_t emu_classify(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent,
                             const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* b1,
                             const uint64_t* o1, const uint8_t* b2, const uint64_t* o2, uint32_t n, double confidence,
                             int min_hit_groups, emu_result* res, uint64_t* hit_off, slk_hit* hits_out, uint64_t cap,
                             int packed) {
  int64_t r = -2;
  DISPATCH_W(sp->w, r = classify_w<W_>(sp, cells, n_buckets, parent, depth, raw, n_dense, root, b1, o1, b2, o2, n,
                                       confidence, min_hit_groups, res, hit_off, hits_out, cap, packed != 0));
  return r;
}

// The split path (scan -> span words | probe | merge + resolve), one fragment at a time, with the same outputs as
// emu_classify: what slk_scan_spans_dev / slk_probe_spans_dev / slk_resolve_spans_dev do on the device.
template <int W>
static int64_t classify_split_w(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent,
                                const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* b1,
                                const uint64_t* o1, const uint8_t* b2, const uint64_t* o2, uint32_t n, double confidence,
                                int min_hit_groups, emu_result* res, uint64_t* hit_off, slk_hit* hits_out, uint64_t cap) {
  slk_table_view tb{cells, n_buckets, 1, 0};
  slk_tax_view tx{parent, depth, raw, n_dense, root};
  uint64_t used = 0;
  std::vector<uint64_t> spans;
  std::vector<uint16_t> dense;
  for (uint32_t r = 0; r < n; r++) {
    spans.clear();
    slk_scan_fragment_spans<W>(*sp, b1 + o1[r], (uint32_t)(o1[r + 1] - o1[r]), b2 ? b2 + o2[r] : nullptr,
                               b2 ? (uint32_t)(o2[r + 1] - o2[r]) : 0u, b2 != nullptr, [&](uint64_t w) { spans.push_back(w); });
    dense.assign(spans.size(), 0);
    for (size_t i = 0; i < spans.size(); i++)
      if (SLK_SPAN_TYPE(spans[i]) == SLK_E_SEQ) dense[i] = (uint16_t)slk_probe(tb, SLK_SPAN_KEY(spans[i]));
    slk_frag_result fr;
    hit_off[r] = used;
    bool full = false;
    slk_resolve_spans(tx, sp->k, spans.data(), dense.data(), (uint32_t)spans.size(), confidence, min_hit_groups,
                      [&](int32_t l, int32_t c) { if (used < cap) hits_out[used++] = slk_hit{l >= 0 ? raw[l] : l, c}; else full = true; }, fr);
    if (full) return -1;
    res[r] = emu_result{fr.taxon, fr.flags, fr.kmers1, fr.kmers2, fr.num_distinct, fr.n_hits};
  }
  return (int64_t)used;
}
EMU_API int64_t emu_classify_split(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent,
                                   const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* b1,
                                   const uint64_t* o1, const uint8_t* b2, const uint64_t* o2, uint32_t n, double confidence,
                                   int min_hit_groups, emu_result* res, uint64_t* hit_off, slk_hit* hits_out, uint64_t cap) {
  int64_t r = -2;
  DISPATCH_W(sp->w, r = classify_split_w<W_>(sp, cells, n_buckets, parent, depth, raw, n_dense, root, b1, o1, b2, o2, n,
                                             confidence, min_hit_groups, res, hit_off, hits_out, cap));
  return r;
}
EMU_API uint32_t emu_shard_of(uint64_t ckey, uint32_t world) { return slk_shard_of(ckey, world); }
EMU_API uint32_t emu_key_mix(uint64_t ckey) { return slk_key_mix(ckey); }
// home bucket of a key in one shard (n_buckets buckets) of a table cut for `world` GPUs
EMU_API uint64_t emu_bucket_of(uint64_t ckey, uint64_t n_buckets, uint32_t world) {
  const slk_table_view tb{nullptr, n_buckets, world, 0};
  return slk_bucket_of(ckey, tb);
}

// the three device steps of the split path one by one (what tests/test_dist_gloo.py plugs into ShardedClassifier)
template <int W>
static int64_t scan_spans_w(const slk_scan_params* sp, const uint8_t* b1, const uint64_t* o1, const uint8_t* b2,
                            const uint64_t* o2, uint32_t n, uint64_t* span_off, uint64_t* spans, uint64_t cap) {
  uint64_t used = 0;
  bool full = false;
  for (uint32_t r = 0; r < n; r++) {
    span_off[r] = used;
    slk_scan_fragment_spans<W>(*sp, b1 + o1[r], (uint32_t)(o1[r + 1] - o1[r]), b2 ? b2 + o2[r] : nullptr,
                               b2 ? (uint32_t)(o2[r + 1] - o2[r]) : 0u, b2 != nullptr,
                               [&](uint64_t w) { if (spans) { if (used < cap) spans[used] = w; else full = true; } used++; });
  }
  span_off[n] = used;
  return full ? -1 : (int64_t)used;
}
EMU_API int64_t emu_scan_spans(const slk_scan_params* sp, const uint8_t* b1, const uint64_t* o1, const uint8_t* b2,
                               const uint64_t* o2, uint32_t n, uint64_t* span_off, uint64_t* spans, uint64_t cap) {
  int64_t r = -2;
  DISPATCH_W(sp->w, r = scan_spans_w<W_>(sp, b1, o1, b2, o2, n, span_off, spans, cap));
  return r;
}
EMU_API void emu_probe_keys(uint64_t* cells, uint64_t n_buckets, const int32_t* raw, const uint64_t* keys, uint64_t n, int32_t* taxa,
                            uint32_t world) {
  slk_table_view tb{cells, n_buckets, world ? world : 1u, 0};
  for (uint64_t i = 0; i < n; i++) { uint32_t d = slk_probe(tb, keys[i]); taxa[i] = d ? raw[d] : 0; }
}
EMU_API void emu_resolve_spans(const uint16_t* parent, const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root,
                               int32_t k, const uint64_t* spans, const uint64_t* span_off, const uint16_t* dense, uint32_t n_reads,
                               double confidence, int min_hit_groups, emu_result* res, slk_hit* hits_out) {
  slk_tax_view tx{parent, depth, raw, n_dense, root};
  for (uint32_t r = 0; r < n_reads; r++) {
    slk_frag_result fr;
    slk_hit* ho = hits_out + span_off[r];
    slk_resolve_spans(tx, k, spans + span_off[r], dense + span_off[r], (uint32_t)(span_off[r + 1] - span_off[r]), confidence,
                      min_hit_groups, [&](int32_t l, int32_t c) { *ho++ = slk_hit{l >= 0 ? raw[l] : l, c}; }, fr);
    res[r] = emu_result{fr.taxon, fr.flags, fr.kmers1, fr.kmers2, fr.num_distinct, fr.n_hits};
  }
}

// Bracken weights of one genome fragment: scan -> hits, lookup, sliding window; dest_out[i] = raw destination taxon of the
// read that starts at i (what bracken_count/emit/probe/window kernels do on the device)
template <int W>
static int64_t bracken_w(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent, const uint8_t* depth,
                         const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* bases, uint32_t len, uint32_t read_len,
                         int32_t* dest_out) {
  slk_table_view tb{cells, n_buckets, 1, 0};
  slk_tax_view tx{parent, depth, raw, n_dense, root};
  std::vector<slk_bhit> hits;
  slk_bracken_scan<W>(*sp, bases, len, [&](const slk_bhit& h) { hits.push_back(h); });
  for (auto& h : hits)
    if (h.flags & SLK_BHIT_SEQ) h.taxon = slk_probe(tb, h.key);
  int64_t n = 0;
  bool ok = slk_bracken_window(tx, hits.data(), (uint32_t)hits.size(), len, read_len, (uint32_t)sp->k,
                               [&](uint32_t d) { dest_out[n++] = raw[d]; });
  return ok ? n : -1;
}
EMU_API int64_t emu_bracken(const slk_scan_params* sp, uint64_t* cells, uint64_t n_buckets, const uint16_t* parent,
                            const uint8_t* depth, const int32_t* raw, uint32_t n_dense, uint32_t root, const uint8_t* bases,
                            uint32_t len, uint32_t read_len, int32_t* dest_out) {
  int64_t r = -2;
  DISPATCH_W(sp->w, r = bracken_w<W_>(sp, cells, n_buckets, parent, depth, raw, n_dense, root, bases, len, read_len, dest_out));
  return r;
}

// the emit "threads" of one fragment, BUILD_WPT windows each, exactly like emit_cells_kernel partitions the work
EMU_API int64_t emu_emit_cells(const slk_scan_params* sp, const uint8_t* bases, uint64_t len, uint32_t dense_taxon,
                               uint32_t wpt, uint64_t* out, uint64_t cap) {
  uint64_t n = 0;
  bool overflow = false;
  auto emit = [&](uint64_t cell) { if (n < cap) out[n++] = cell; else overflow = true; };
  uint64_t nw = len >= (uint64_t)sp->k ? len - sp->k + 1 : 0;
  for (uint64_t w0 = 0; w0 < nw; w0 += wpt) {
    uint64_t nb = std::min<uint64_t>(len - w0, (uint64_t)wpt + sp->k - 1);
    DISPATCH_W(sp->w, (slk_emit_cells<W_>(*sp, bases + w0, nb, dense_taxon, emit)));
  }
  return overflow ? -1 : (int64_t)n;
}

EMU_API void emu_synth_genome(uint64_t seed, uint64_t start, uint64_t n, uint8_t* out) {
  for (uint64_t i = 0; i < n; i++) out[i] = slk_synth_genome_base(seed, start + i);
}
EMU_API void emu_synth_reads(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len, uint64_t first,
                             uint64_t n_reads, uint32_t L, uint8_t* out) {
  for (uint64_t r = 0; r < n_reads; r++)
    for (uint32_t j = 0; j < L; j++) out[r * L + j] = slk_synth_read_base(gseed, rseed, n_genomes, genome_len, first + r, L, j);
}
EMU_API void emu_synth_mates(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len, uint64_t first,
                             uint64_t n_reads, uint32_t L, uint32_t mate, uint8_t* out) {
  for (uint64_t r = 0; r < n_reads; r++)
    for (uint32_t j = 0; j < L; j++) out[r * L + j] = slk_synth_read_base(gseed, rseed, n_genomes, genome_len, first + r, L, j, mate);
}
