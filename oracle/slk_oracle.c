/*
 * slk_oracle.c -- CPU ORACLE for the Slacken Kraken-2-style build/classify hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT. Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it. The product path
 * (slacken_b200/, libslacken_gpu.so) never links, imports or calls anything in oracle/.
 *
 * It is a plain-C restatement of the reference's Scala algorithms (the reference itself
 * cannot be compiled or run here: no JVM/Scala/Spark in this image). Every function cites
 * the reference file:line it follows; paths are relative to
 * /root/reference/src/main/scala/com/jnpersson/.
 *
 * Pins (see tests/test_oracle_*.py): the MinSplitterTest known answer
 * (src/test/.../kmers/minimizer/MinSplitterTest.scala:25-32), the derived constants of
 * SURVEY.md section 8(a5), and the property suites of the reference's scalacheck tests
 * transcribed with hypothesis. Exact per-read lines / kreport text are NOT pinned by any
 * test the reference holds ("parity unpinned" for those; this restatement is the pin).
 *
 * Restrictions: minimizer width m <= 32 (one 64-bit word per minimizer, id1 only).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SLKO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * Parameters of the minimizer scheme (IndexParams / SplitterFormat: kmers/IndexParams.scala:63-91,
 * kmers/SplitterFormat.scala:55-77; defaults slacken/Slacken.scala:126-136).
 * ordering: 0 = RandomXOR (optionally canonical, optionally wrapped in SpacedSeed when spaces>0)
 *           1 = MinTable.ofLength(m) = plain lexicographic ordering of the forward m-mer
 *               (only used by the reference's MinSplitterTest known answer).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t k, m, spaces, canonical, ordering;
  uint64_t toggle_mask; /* XORmask property, e.g. DEFAULT_TOGGLE_MASK 0xe37e28c4271b5a2d */
} slko_params;

enum { TWOBIT_WHITESPACE = 4, TWOBIT_INVALID = 5 };
enum { SEQUENCE_FLAG = 1, AMBIGUOUS_FLAG = 2, MATE_PAIR_BORDER_FLAG = 3 }; /* slacken/package.scala:37-39 */
enum { TAXON_NONE = 0, TAXON_ROOT = 1, AMBIGUOUS_SPAN = -1, MATE_PAIR_BORDER = -2 }; /* slacken/package.scala:28-29, Taxonomy.scala:30-31 */

/* kmers/util/BitRepresentation.scala:127-165 (charToTwobitWithInvalid) */
static inline int char_to_twobit(unsigned char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'U': case 'u': return 3;
    case '\n': case '\r': return TWOBIT_WHITESPACE;
    default: return TWOBIT_INVALID;
  }
}
/* kmers/util/BitRepresentation.scala:140-143 (isValid) */
static inline int is_valid_char(unsigned char c) { return char_to_twobit(c) < 4; }

SLKO_API int slko_char_to_twobit(int c) { return char_to_twobit((unsigned char)c); }

/* kmers/util/BitRepresentation.scala:60-73 (swapNTSequence): reverse the order of 2-bit groups */
static inline uint64_t swap_nt_sequence(uint64_t kmer) {
  kmer = ((kmer & 0xCCCCCCCCCCCCCCCCull) >> 2) | ((kmer & 0x3333333333333333ull) << 2);
  kmer = ((kmer & 0xF0F0F0F0F0F0F0F0ull) >> 4) | ((kmer & 0x0F0F0F0F0F0F0F0Full) << 4);
  kmer = ((kmer & 0xFF00FF00FF00FF00ull) >> 8) | ((kmer & 0x00FF00FF00FF00FFull) << 8);
  kmer = ((kmer & 0xFFFF0000FFFF0000ull) >> 16) | ((kmer & 0x0000FFFF0000FFFFull) << 16);
  return (kmer >> 32) | (kmer << 32);
}

/* kmers/util/NTBitArray.scala:231-247 (writeReverseComplement) for a single left-aligned long:
 * l = 1, shiftAmt = (size%32)*2; data0 = swap(x) ^ -1; data0 <<= (64 - shiftAmt) (Java shifts mod 64) */
static inline uint64_t revcomp_left_aligned(uint64_t x, int size) {
  uint64_t r = swap_nt_sequence(x) ^ ~0ull;
  int shiftAmt = (size % 32) * 2;
  int sh = (64 - shiftAmt) & 63;
  return r << sh;
}
SLKO_API uint64_t slko_revcomp(uint64_t x, int size) { return revcomp_left_aligned(x, size); }

/* kmers/util/NTBitArray.scala:455-460 (apply) */
static inline int nt_at(uint64_t x, int pos) { return (int)((x >> (2 * (31 - pos))) & 3); }

/* kmers/util/NTBitArray.scala:437-452 (sliceIsForwardOrientation(0,size)) */
static inline int is_forward_orientation(uint64_t x, int size) {
  int st = 0, end = size - 1;
  while (st < end) {
    int a = nt_at(x, st);
    int b = (~nt_at(x, end)) & 3; /* complementOne, BitRepresentation.scala:47 */
    if (a < b) return 1;
    if (a > b) return 0;
    st++; end--;
  }
  return nt_at(x, st) < 2; /* apply(st) < G */
}

/* kmers/minimizer/MinimizerPriorities.scala:146-160 (RandomXOR.mask) for a single long */
static inline uint64_t xor_mask_for(const slko_params* p) {
  int w = p->m;
  if (w % 32 != 0) return p->toggle_mask << (64 - (w % 32) * 2);
  return p->toggle_mask;
}
SLKO_API uint64_t slko_xor_mask(const slko_params* p) { return xor_mask_for(p); }

/* kmers/minimizer/MinimizerPriorities.scala:287-301 (SpacedSeed.spaceMask) for a single long:
 * r = fill(-1,width) (NTBitArray.scala:104-111), then s times: r <<= 4; r |= finalBits */
static inline uint64_t space_mask_for(const slko_params* p) {
  int w = p->m;
  uint64_t r = ~0ull;
  if (w % 32 != 0) r &= (~0ull) << (64 - (w % 32) * 2);
  uint64_t finalBits = 3ull << ((64 - (w % 32) * 2) & 63);
  for (int i = 0; i < p->spaces; i++) { r <<= 4; r |= finalBits; }
  return r;
}
SLKO_API uint64_t slko_space_mask(const slko_params* p) { return space_mask_for(p); }

/* RandomXOR.writePriorityOf (MinimizerPriorities.scala:165-175) + NTBitArray.writeCanonical
 * (NTBitArray.scala:258-266) + SpacedSeed.writePriorityOf (MinimizerPriorities.scala:308-312).
 * ordering 1: MinTable.withAll(width): priorityLookup(motif) = motif, written left-aligned
 * (MinimizerPriorities.scala:246-255) = the forward m-mer itself. */
static inline uint64_t priority_of(const slko_params* p, uint64_t window) {
  if (p->ordering == 1) return window;
  uint64_t x = window;
  if (p->canonical && !is_forward_orientation(window, p->m)) x = revcomp_left_aligned(window, p->m);
  x ^= xor_mask_for(p);
  if (p->spaces > 0) x &= space_mask_for(p);
  return x;
}
SLKO_API uint64_t slko_priority(const slko_params* p, uint64_t window) { return priority_of(p, window); }

/* kmers/util/NTBitArray.scala:140-150 (shiftLongArrayKmerLeft), single long */
static inline uint64_t shift_add_bp(uint64_t w, int b, int m) {
  int kmod32 = m & 31;
  return (w << 2) | ((uint64_t)b << (((32 - kmod32) * 2) & 63));
}

/* Encode a string into a left-aligned window (test helper; NTBitArray.encode, NTBitArray.scala:80-98) */
SLKO_API uint64_t slko_encode_window(const char* s, int len) {
  uint64_t w = 0;
  for (int i = 0; i < len; i++) w = shift_add_bp(w, char_to_twobit((unsigned char)s[i]), len);
  return w;
}

/* ------------------------------------------------------------------------------------------
 * ShiftScanner.allMatches (kmers/minimizer/ShiftScanner.scala:90-159).
 * Returns the number of non-whitespace characters (validSize) or -1 if an invalid character is met
 * (the reference throws InvalidNucleotideException). pri[i], valid[i] for i < validSize.
 * ------------------------------------------------------------------------------------------ */
static int64_t scan_all_matches(const slko_params* p, const char* data, int64_t size,
                                uint64_t* pri, uint8_t* valid) {
  int width = p->m;
  int64_t validSize = 0, pos = 0;
  uint64_t window = 0;
  while (validSize < width - 1 && pos < size) {
    int x = char_to_twobit((unsigned char)data[pos]);
    if (x == TWOBIT_INVALID) return -1;
    if (x != TWOBIT_WHITESPACE) {
      pri[validSize] = 0; valid[validSize] = 0;
      window = shift_add_bp(window, x, width);
      validSize++;
    }
    pos++;
  }
  while (pos < size) {
    int x = char_to_twobit((unsigned char)data[pos]);
    if (x == TWOBIT_INVALID) return -1;
    if (x != TWOBIT_WHITESPACE) {
      window = shift_add_bp(window, x, width);
      pri[validSize] = priority_of(p, window); valid[validSize] = 1;
      validSize++;
    }
    pos++;
  }
  return validSize;
}

/* ------------------------------------------------------------------------------------------
 * PosRankWindow (kmers/minimizer/PosRankWindow.scala:33-97) over MinimizerPositions
 * (kmers/minimizer/MinimizerPositions.scala:55-77: unsigned 64-bit compare).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int m, k; int64_t n; const uint64_t* pri; uint8_t* valid; int64_t leftBound, rightBound;
} pos_rank_window;

static void prw_advance(pos_rank_window* w) {
  w->rightBound += 1;
  if (w->rightBound > w->n) return;
  int64_t inserted = w->rightBound - 1;
  if (w->valid[inserted]) {
    int64_t test = w->rightBound - 2;
    while (test >= w->leftBound + 1 && (!w->valid[test] || w->pri[test] > w->pri[inserted])) {
      w->valid[test] = 0;
      test--;
    }
    if (!w->valid[w->leftBound] || w->pri[inserted] < w->pri[w->leftBound]) w->leftBound += 1;
  }
  while (w->rightBound - w->leftBound > w->k - (w->m - 1) ||
         (w->leftBound < w->n && !w->valid[w->leftBound]))
    w->leftBound += 1;
}
static void prw_init(pos_rank_window* w, int m, int k, int64_t n, const uint64_t* pri, uint8_t* valid) {
  w->m = m; w->k = k; w->n = n; w->pri = pri; w->valid = valid; w->leftBound = 0; w->rightBound = 0;
  while (w->rightBound < k) prw_advance(w);
}
static inline int prw_has_next(const pos_rank_window* w) { return w->rightBound <= w->n; }

/* MinSplitter.superkmerPositions (kmers/minimizer/MinSplitter.scala:180-216).
 * Emits (location, rank, length) triples; returns their number, -1 on invalid char, -2 if cap too small,
 * -3 on "k-length window found with no minimizer" (cannot happen for the XOR orderings). */
static int64_t superkmer_positions(const slko_params* p, const char* data, int64_t size,
                                   int64_t* loc, uint64_t* rank, int32_t* length, int64_t cap,
                                   uint64_t* pri, uint8_t* valid) {
  int64_t n = scan_all_matches(p, data, size, pri, valid);
  if (n < 0) return -1;
  pos_rank_window w;
  prw_init(&w, p->m, p->k, n, pri, valid);
  int64_t regionStart = 0, out = 0;
  while (prw_has_next(&w)) {
    int64_t pos = w.leftBound;
    if (pos >= n || !valid[pos]) return -3;
    prw_advance(&w); /* window.next */
    uint64_t r = pri[pos];
    int64_t consumed = 1;
    while (prw_has_next(&w)) {
      if (w.leftBound >= n) return -3; /* the reference would throw (index out of bounds) */
      if (!(w.leftBound == pos || pri[w.leftBound] == r)) break;
      prw_advance(&w); /* window.next */
      consumed++;
    }
    int64_t thisStart = regionStart;
    regionStart += consumed;
    if (out >= cap) return -2;
    loc[out] = thisStart; rank[out] = r;
    length[out] = prw_has_next(&w) ? (int32_t)(consumed + (p->k - 1)) : (int32_t)(n - thisStart);
    out++;
  }
  return out;
}

SLKO_API int64_t slko_superkmers(const slko_params* p, const char* data, int64_t size,
                                 int64_t* loc, uint64_t* rank, int32_t* length, int64_t cap) {
  uint64_t* pri = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(size + 1));
  uint8_t* valid = (uint8_t*)malloc((size_t)(size + 1));
  int64_t r = superkmer_positions(p, data, size, loc, rank, length, cap, pri, valid);
  free(pri); free(valid);
  return r;
}

/* Raw scanner output for tests: pri[], valid[] per non-whitespace position. */
SLKO_API int64_t slko_all_matches(const slko_params* p, const char* data, int64_t size, uint64_t* pri, uint8_t* valid) {
  return scan_all_matches(p, data, size, pri, valid);
}

/* ------------------------------------------------------------------------------------------
 * Supermers.splitByAmbiguity (slacken/Supermers.scala:141,150-189): regex [actguACTGU\n\r]+.
 * Emits pieces (start, length, flag). Returns the number of pieces.
 * ------------------------------------------------------------------------------------------ */
static inline int is_nonambiguous_char(unsigned char c) { return char_to_twobit(c) != TWOBIT_INVALID; }

SLKO_API int64_t slko_split_by_ambiguity(const char* seq, int64_t len, int k,
                                         int64_t* start, int64_t* plen, int32_t* flag, int64_t cap) {
  int64_t at = 0, out = 0;
  while (at < len) {
    int64_t e = at;
    int f;
    if (is_nonambiguous_char((unsigned char)seq[at])) {
      int c = 0;
      while (e < len && is_nonambiguous_char((unsigned char)seq[e])) e++;
      /* enoughValidChars (Supermers.scala:180-189): count valid (non-newline) chars */
      f = AMBIGUOUS_FLAG;
      for (int64_t i = at; i < e; i++) {
        if (is_valid_char((unsigned char)seq[i])) c++;
        if (c == k) { f = SEQUENCE_FLAG; break; }
      }
    } else {
      while (e < len && !is_nonambiguous_char((unsigned char)seq[e])) e++;
      f = AMBIGUOUS_FLAG;
    }
    if (out >= cap) return -2;
    start[out] = at; plen[out] = e - at; flag[out] = f; out++;
    at = e;
  }
  return out;
}

/* A span, before the join: OrdinalSpan (slacken/package.scala:61-62). ordinal = index in the array. */
typedef struct { uint64_t minimizer; int32_t kmers; uint8_t flag; uint8_t distinct; } span_t;

typedef struct { span_t* v; int64_t n, cap; } span_vec;
static void sv_push(span_vec* s, span_t x) {
  if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->v = (span_t*)realloc(s->v, sizeof(span_t) * (size_t)s->cap); }
  s->v[s->n++] = x;
}

typedef struct { uint64_t* pri; uint8_t* valid; int64_t* loc; uint64_t* rank; int32_t* length; int64_t cap; } scratch_t;
static void scratch_reserve(scratch_t* s, int64_t n) {
  if (n <= s->cap) return;
  s->cap = n + 64;
  s->pri = (uint64_t*)realloc(s->pri, sizeof(uint64_t) * (size_t)s->cap);
  s->valid = (uint8_t*)realloc(s->valid, (size_t)s->cap);
  s->loc = (int64_t*)realloc(s->loc, sizeof(int64_t) * (size_t)s->cap);
  s->rank = (uint64_t*)realloc(s->rank, sizeof(uint64_t) * (size_t)s->cap);
  s->length = (int32_t*)realloc(s->length, sizeof(int32_t) * (size_t)s->cap);
}
static void scratch_free(scratch_t* s) { free(s->pri); free(s->valid); free(s->loc); free(s->rank); free(s->length); memset(s, 0, sizeof(*s)); }

/* Supermers.splitFragment(NTSeq) (slacken/Supermers.scala:113-125): pieces shorter than k vanish;
 * AMBIGUOUS piece -> one span of kmers = len-(k-1) (random minimizer, never looked at);
 * SEQUENCE piece -> splitter.splitEncode, span.kmers = supermer length - (k-1) (Supermers.scala:94). */
static int split_fragment_nt(const slko_params* p, const char* seq, int64_t len, span_vec* out, scratch_t* sc) {
  int64_t at = 0;
  int k = p->k;
  while (at < len) {
    int64_t e = at; int f;
    if (is_nonambiguous_char((unsigned char)seq[at])) {
      int c = 0;
      while (e < len && is_nonambiguous_char((unsigned char)seq[e])) e++;
      f = AMBIGUOUS_FLAG;
      for (int64_t i = at; i < e; i++) { if (is_valid_char((unsigned char)seq[i])) c++; if (c == k) { f = SEQUENCE_FLAG; break; } }
    } else {
      while (e < len && !is_nonambiguous_char((unsigned char)seq[e])) e++;
      f = AMBIGUOUS_FLAG;
    }
    int64_t plen = e - at;
    if (plen >= k) {
      if (f == AMBIGUOUS_FLAG) {
        span_t s = {0, (int32_t)(plen - (k - 1)), AMBIGUOUS_FLAG, 0};
        sv_push(out, s);
      } else {
        scratch_reserve(sc, plen + 1);
        int64_t ns = superkmer_positions(p, seq + at, plen, sc->loc, sc->rank, sc->length, sc->cap, sc->pri, sc->valid);
        if (ns < 0) return (int)ns;
        for (int64_t i = 0; i < ns; i++) {
          span_t s = {sc->rank[i], sc->length[i] - (k - 1), SEQUENCE_FLAG, 0};
          sv_push(out, s);
        }
      }
    }
    at = e;
  }
  return 0;
}

/* Supermers.splitFragment(InputFragment) + spans (slacken/Supermers.scala:49-97):
 * R1 spans, a MATE_PAIR_BORDER pseudo span (0 nucleotides -> kmers = -(k-1)), R2 spans;
 * distinct = SEQUENCE && (first || rank != lastMinimizer) with `first` cleared by ANY span and
 * lastMinimizer only updated by SEQUENCE spans. nt2 == NULL for single-end. */
static int fragment_spans(const slko_params* p, const char* nt1, int64_t len1, const char* nt2, int64_t len2,
                          span_vec* out, scratch_t* sc) {
  out->n = 0;
  int rc = split_fragment_nt(p, nt1, len1, out, sc);
  if (rc < 0) return rc;
  if (nt2) {
    span_t b = {0, -(p->k - 1), MATE_PAIR_BORDER_FLAG, 0};
    sv_push(out, b);
    rc = split_fragment_nt(p, nt2, len2, out, sc);
    if (rc < 0) return rc;
  }
  int first = 1, haveLast = 0; uint64_t last = 0;
  for (int64_t i = 0; i < out->n; i++) {
    span_t* s = &out->v[i];
    int isSeq = (s->flag != AMBIGUOUS_FLAG && s->flag != MATE_PAIR_BORDER_FLAG);
    /* lastMinimizer starts as an empty array, which never equals a rank (Supermers.scala:73,85) */
    s->distinct = (uint8_t)(isSeq && (first || !haveLast || s->minimizer != last));
    if (isSeq) { last = s->minimizer; haveLast = 1; }
    first = 0;
  }
  return 0;
}

SLKO_API int64_t slko_spans(const slko_params* p, const char* nt1, int64_t len1, const char* nt2, int64_t len2,
                            uint64_t* minimizer, uint8_t* distinct, int32_t* kmers, uint8_t* flag, int64_t cap) {
  span_vec sv = {0, 0, 0}; scratch_t sc; memset(&sc, 0, sizeof(sc));
  int rc = fragment_spans(p, nt1, len1, nt2, len2, &sv, &sc);
  int64_t n = rc < 0 ? rc : sv.n;
  if (rc >= 0) {
    if (n > cap) n = -2;
    else for (int64_t i = 0; i < n; i++) { minimizer[i] = sv.v[i].minimizer; distinct[i] = sv.v[i].distinct; kmers[i] = sv.v[i].kmers; flag[i] = sv.v[i].flag; }
  }
  free(sv.v); scratch_free(&sc);
  return n;
}

/* ------------------------------------------------------------------------------------------
 * LowestCommonAncestor.apply (slacken/LowestCommonAncestor.scala:49-78). parents[] indexed by raw taxid.
 * ------------------------------------------------------------------------------------------ */
#define PATH_MAX_LENGTH 256
static int32_t lca_apply(const int32_t* parents, int32_t tax1, int32_t tax2) {
  if (tax1 == TAXON_NONE || tax2 == TAXON_NONE) return tax2 == TAXON_NONE ? tax1 : tax2;
  int32_t path[PATH_MAX_LENGTH + 1];
  int32_t a = tax1; int i = 0;
  while (a != TAXON_NONE && i < PATH_MAX_LENGTH) { path[i++] = a; a = parents[a]; }
  path[i] = TAXON_NONE;
  int32_t b = tax2;
  while (b != TAXON_NONE) {
    for (i = 0; path[i] != TAXON_NONE; i++) if (path[i] == b) return b;
    b = parents[b];
  }
  return TAXON_ROOT;
}
SLKO_API int32_t slko_lca(const int32_t* parents, int32_t a, int32_t b) { return lca_apply(parents, a, b); }

/* Taxonomy.hasAncestor / stepsToAncestor (slacken/Taxonomy.scala:236-244): pathToRoot(tax) contains ancestor */
static int has_ancestor(const int32_t* parents, int32_t tax, int32_t ancestor) {
  int32_t t = tax;
  while (t != TAXON_NONE) { if (t == ancestor) return 1; t = parents[t]; }
  return 0;
}
SLKO_API int slko_has_ancestor(const int32_t* parents, int32_t tax, int32_t anc) { return has_ancestor(parents, tax, anc); }

/* ------------------------------------------------------------------------------------------
 * A hit, after the join: TaxonHit (slacken/KeyValueIndex.scala:436-441), and TaxonCounts
 * (slacken/TaxonCounts.scala). Insertion-ordered small map = fastutil Int2IntArrayMap.
 * ------------------------------------------------------------------------------------------ */
typedef struct { int32_t taxon; int32_t count; } slko_hit;

typedef struct { int32_t* key; int32_t* val; int n, cap; } int_map;
static int map_get(const int_map* m, int32_t k) { for (int i = 0; i < m->n; i++) if (m->key[i] == k) return m->val[i]; return 0; }
static void map_add(int_map* m, int32_t k, int32_t v) {
  for (int i = 0; i < m->n; i++) if (m->key[i] == k) { m->val[i] += v; return; }
  if (m->n == m->cap) { m->cap = m->cap ? 2 * m->cap : 32; m->key = (int32_t*)realloc(m->key, 4 * (size_t)m->cap); m->val = (int32_t*)realloc(m->val, 4 * (size_t)m->cap); }
  m->key[m->n] = k; m->val[m->n] = v; m->n++;
}

/* LowestCommonAncestor.resolveTree(Int2IntMap, requiredScore) (slacken/LowestCommonAncestor.scala:101-146) */
static int32_t resolve_tree_map(const int32_t* parents, const int_map* hc, double requiredScore) {
  int32_t maxTaxon = 0; int maxScore = 0;
  for (int i = 0; i < hc->n; i++) {
    int32_t taxon = hc->key[i], node = taxon; int score = 0;
    while (node != TAXON_NONE) { score += map_get(hc, node); node = parents[node]; }
    if (score > maxScore) { maxTaxon = taxon; maxScore = score; }
    else if (score == maxScore) maxTaxon = lca_apply(parents, maxTaxon, taxon);
  }
  maxScore = map_get(hc, maxTaxon);
  while (maxTaxon != TAXON_NONE && (double)maxScore < requiredScore) {
    maxScore = 0;
    for (int i = 0; i < hc->n; i++)
      if (has_ancestor(parents, hc->key[i], maxTaxon)) maxScore += hc->val[i];
    if ((double)maxScore >= requiredScore) return maxTaxon;
    maxTaxon = parents[maxTaxon];
  }
  return maxTaxon;
}

/* resolveTree(TaxonCounts, confidence) (LowestCommonAncestor.scala:91-96) with
 * TaxonCounts.toMap / totalKmers (slacken/TaxonCounts.scala:70-87) over MERGED hits. */
static int32_t resolve_tree_hits(const int32_t* parents, const slko_hit* merged, int n, double confidence, int_map* scratch) {
  scratch->n = 0;
  int total = 0;
  for (int i = 0; i < n; i++) {
    int32_t t = merged[i].taxon;
    if (t != AMBIGUOUS_SPAN && t != MATE_PAIR_BORDER) map_add(scratch, t, merged[i].count);
    if (t != MATE_PAIR_BORDER) total += merged[i].count;
  }
  double required = ceil(confidence * (double)total);
  return resolve_tree_map(parents, scratch, required);
}

/* Test entry: resolveTree over a raw hit list (TaxonCounts.fromHits merge first, TaxonCounts.scala:31-48). */
static int merge_hits(const slko_hit* hits, int n, slko_hit* out) {
  int m = 0;
  for (int i = 0; i < n; i++) {
    if (m > 0 && out[m - 1].taxon == hits[i].taxon) out[m - 1].count += hits[i].count;
    else out[m++] = hits[i];
  }
  return m;
}
SLKO_API int32_t slko_resolve_tree(const int32_t* parents, const int32_t* taxa, const int32_t* counts, int n, double confidence) {
  slko_hit* h = (slko_hit*)malloc(sizeof(slko_hit) * (size_t)(n + 1));
  slko_hit* mg = (slko_hit*)malloc(sizeof(slko_hit) * (size_t)(n + 1));
  for (int i = 0; i < n; i++) { h[i].taxon = taxa[i]; h[i].count = counts[i]; }
  int m = merge_hits(h, n, mg);
  int_map mp = {0, 0, 0, 0};
  int32_t r = resolve_tree_hits(parents, mg, m, confidence, &mp);
  free(mp.key); free(mp.val); free(h); free(mg);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Library records: the result of KeyValueIndex.makeRecords (slacken/KeyValueIndex.scala:85-93):
 * groupBy(id1).agg(TaxonLCA). Held as a CPU open-addressing table (the reference's own lookup is a
 * Spark sort-merge left join on id1, slacken/Classifier.scala:84; only its relational result matters).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  uint64_t* keys; int32_t* taxa; uint64_t nslots; uint64_t count;
  int update_only;   /* checking aid (see slko_lib_set_update_only): minimizers that are not in the table yet are skipped */
} slko_lib;
#define EMPTY_KEY (~0ull) /* a priority is never all ones for m<=31 (low pad bits are zero) */

static inline uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}

SLKO_API slko_lib* slko_lib_create(uint64_t expected_keys) {
  slko_lib* L = (slko_lib*)calloc(1, sizeof(slko_lib));
  uint64_t n = 16;
  while ((double)n * 0.7 < (double)expected_keys + 16.0) n <<= 1;
  L->nslots = n;
  L->keys = (uint64_t*)malloc(sizeof(uint64_t) * n);
  L->taxa = (int32_t*)calloc(n, sizeof(int32_t));
  if (!L->keys || !L->taxa) { free(L->keys); free(L->taxa); free(L); return NULL; }
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) L->keys[i] = EMPTY_KEY;
  return L;
}
SLKO_API void slko_lib_destroy(slko_lib* L) { if (L) { free(L->keys); free(L->taxa); free(L); } }
SLKO_API uint64_t slko_lib_size(const slko_lib* L) { return L->count; }

/* Thread-safe insert with LCA merge: TaxonLCA.reduce/merge (slacken/LowestCommonAncestor.scala:152-170);
 * zero = NONE and lca(NONE, x) = x, so a zero-initialised taxon slot is the aggregator's zero. */
static void lib_insert(slko_lib* L, const int32_t* parents, uint64_t key, int32_t taxon) {
  uint64_t mask = L->nslots - 1, i = mix64(key) & mask;
  for (;;) {
    uint64_t cur = __atomic_load_n(&L->keys[i], __ATOMIC_ACQUIRE);
    if (cur == EMPTY_KEY) {
      if (L->update_only) return;
      uint64_t exp = EMPTY_KEY;
      if (__atomic_compare_exchange_n(&L->keys[i], &exp, key, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) {
        __atomic_fetch_add(&L->count, 1, __ATOMIC_RELAXED);
        cur = key;
      } else cur = exp;
    }
    if (cur == key) {
      int32_t old = __atomic_load_n(&L->taxa[i], __ATOMIC_ACQUIRE);
      for (;;) {
        int32_t nw = lca_apply(parents, old, taxon);
        if (nw == old) return;
        if (__atomic_compare_exchange_n(&L->taxa[i], &old, nw, 0, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) return;
      }
    }
    i = (i + 1) & mask;
  }
}
static inline int lib_lookup(const slko_lib* L, uint64_t key, int32_t* taxon) {
  uint64_t mask = L->nslots - 1, i = mix64(key) & mask;
  for (;;) {
    uint64_t cur = L->keys[i];
    if (cur == key) { *taxon = L->taxa[i]; return 1; }
    if (cur == EMPTY_KEY) return 0;
    i = (i + 1) & mask;
  }
}
SLKO_API int slko_lib_lookup(const slko_lib* L, uint64_t key, int32_t* taxon) { return lib_lookup(L, key, taxon); }
/* Checking a sample of reads against a library too large for host memory (the 70 Gbp configuration): insert the minimizers
 * of the SAMPLE first, clear their taxa, switch the table to update-only and feed it every genome of the library: it then
 * holds exactly the library's records for the sample's minimizers (taxon 0 = the library does not have it). */
SLKO_API void slko_lib_clear_taxa(slko_lib* L) { memset(L->taxa, 0, sizeof(int32_t) * L->nslots); }
SLKO_API void slko_lib_set_update_only(slko_lib* L, int on) { L->update_only = on; }

/* Insert ready-made records (a library loaded from Parquet, KeyValueIndex.loadRecords :150-159). */
SLKO_API void slko_lib_add_records(slko_lib* L, const int32_t* parents, const uint64_t* id1, const int32_t* taxon, uint64_t n) {
  #pragma omp parallel for schedule(dynamic, 4096)
  for (int64_t i = 0; i < (int64_t)n; i++) lib_insert(L, parents, id1[i], taxon[i]);
}

/* SplitterMinimizers.find (slacken/Minimizers.scala:43-76) + makeRecords (KeyValueIndex.scala:85-93,118-120):
 * every super-mer of every (taxon, fragment) contributes (rank, taxon); fragments whose taxon is undefined
 * (Taxonomy.isDefined, slacken/Taxonomy.scala:175-176) are dropped. Fragments are "valid-only" sequences
 * (InputReader.removeInvalid, kmers/input/InputReader.scala:60-72) that may contain newlines.
 * Returns 0, or -1 if a fragment contains an invalid character. */
SLKO_API int slko_lib_add_fragments(slko_lib* L, const slko_params* p, const int32_t* parents, int32_t n_tax,
                                    const char* bases, const int64_t* off, const int32_t* frag_taxon, int64_t n_frag) {
  int err = 0;
  #pragma omp parallel
  {
    scratch_t sc; memset(&sc, 0, sizeof(sc));
    #pragma omp for schedule(dynamic, 1)
    for (int64_t f = 0; f < n_frag; f++) {
      int32_t t = frag_taxon[f];
      if (t < 0 || t >= n_tax) continue;
      if (!(parents[t] != TAXON_NONE || t == TAXON_ROOT)) continue;
      int64_t len = off[f + 1] - off[f];
      /* long fragments are cut into chunks overlapping by k-1 so scratch stays bounded; every k-mer window
       * is seen exactly once, and splitting a super-mer only duplicates a (rank, taxon) pair, which
       * groupBy+LCA absorbs. Whitespace-free chunks only, otherwise scan the fragment whole. */
      const char* s = bases + off[f];
      int has_ws = memchr(s, '\n', (size_t)len) != NULL || memchr(s, '\r', (size_t)len) != NULL;
      int64_t chunk = has_ws ? len : (1 << 20);
      for (int64_t st = 0; st < len; st += chunk) {
        int64_t e = st + chunk + (p->k - 1); if (e > len) e = len;
        int64_t l = e - st;
        scratch_reserve(&sc, l + 1);
        int64_t ns = superkmer_positions(p, s + st, l, sc.loc, sc.rank, sc.length, sc.cap, sc.pri, sc.valid);
        if (ns < 0) { err = -1; break; }
        for (int64_t i = 0; i < ns; i++) lib_insert(L, parents, sc.rank[i], t);
        if (e == len) break;
      }
    }
    scratch_free(&sc);
  }
  return err;
}

/* InputReader.removeInvalid (kmers/input/InputReader.scala:56-72) + slko_lib_add_fragments: regex
 * [ACTGUactgu][ACTGUactgu\n\r]* -- every maximal run that starts with a base and continues over bases and
 * newlines is a fragment of its own with the sequence's label. Used where the sequences still contain N. */
SLKO_API int slko_lib_add_sequences(slko_lib* L, const slko_params* p, const int32_t* parents, int32_t n_tax,
                                    const char* bases, const int64_t* off, const int32_t* seq_taxon, int64_t n_seq) {
  int64_t cap = 1024, n = 0;
  int64_t* poff = (int64_t*)malloc(sizeof(int64_t) * 2 * (size_t)cap);
  int32_t* ptax = (int32_t*)malloc(sizeof(int32_t) * (size_t)cap);
  for (int64_t s = 0; s < n_seq; s++) {
    int64_t i = off[s], e = off[s + 1];
    while (i < e) {
      if (!is_valid_char((unsigned char)bases[i])) { i++; continue; }
      int64_t j = i + 1;
      while (j < e && is_nonambiguous_char((unsigned char)bases[j])) j++;
      if (n == cap) { cap *= 2; poff = (int64_t*)realloc(poff, sizeof(int64_t) * 2 * (size_t)cap); ptax = (int32_t*)realloc(ptax, sizeof(int32_t) * (size_t)cap); }
      poff[2 * n] = i; poff[2 * n + 1] = j; ptax[n] = seq_taxon[s]; n++;
      i = j;
    }
  }
  int err = 0;
  #pragma omp parallel for schedule(dynamic, 1)
  for (int64_t f = 0; f < n; f++) {
    int64_t o2[2] = {0, poff[2 * f + 1] - poff[2 * f]};
    int rc = slko_lib_add_fragments(L, p, parents, n_tax, bases + poff[2 * f], o2, &ptax[f], 1);
    if (rc < 0) err = rc;
  }
  free(poff); free(ptax);
  return err;
}

/* Dump the records (unordered set of (id1, taxon) with unique id1). */
SLKO_API uint64_t slko_lib_records(const slko_lib* L, uint64_t* id1, int32_t* taxon, uint64_t cap) {
  uint64_t n = 0;
  for (uint64_t i = 0; i < L->nslots; i++) if (L->keys[i] != EMPTY_KEY) {
    if (n < cap) { id1[n] = L->keys[i]; taxon[n] = L->taxa[i]; }
    n++;
  }
  return n;
}

/* ------------------------------------------------------------------------------------------
 * Classify one fragment: KeyValueIndex.getSpans (:163-173) -> left join + spanToHit (:176-185) ->
 * spansToGroupedHits numDistinct (Classifier.scala:92-95) -> classifyHits sort by ordinal (:136) ->
 * Classifier.classify (:439-454) -> TaxonCounts.fromHits / lengthString (TaxonCounts.scala:31-48,114-121).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t taxon;        /* reported taxon (0 when unclassified) */
  uint8_t classified;
  uint8_t has_span;     /* 0 -> the read yields no output line at all (no span -> no group in the groupBy) */
  int32_t num_distinct;
  int32_t len1, len2;   /* lengthString parts; len2 = -1 for single-end */
  int32_t n_hits;       /* merged hits written */
} slko_result;

typedef struct { span_vec sv; scratch_t sc; int_map mp; slko_hit* hits; slko_hit* merged; int64_t hcap; } classify_ws;

static int classify_fragment(const slko_params* p, const int32_t* parents, const slko_lib* L,
                             const char* nt1, int64_t len1, const char* nt2, int64_t len2,
                             double confidence, int min_hit_groups,
                             classify_ws* ws, slko_result* res) {
  int rc = fragment_spans(p, nt1, len1, nt2, len2, &ws->sv, &ws->sc);
  if (rc < 0) return rc;
  int64_t n = ws->sv.n;
  memset(res, 0, sizeof(*res));
  res->len2 = -1;
  if (n == 0) return 0;            /* vanishes from all outputs (SURVEY 8a "vanishing reads") */
  res->has_span = 1;
  if (n > ws->hcap) { ws->hcap = n + 64; ws->hits = (slko_hit*)realloc(ws->hits, sizeof(slko_hit) * (size_t)ws->hcap); ws->merged = (slko_hit*)realloc(ws->merged, sizeof(slko_hit) * (size_t)ws->hcap); }
  int numDistinct = 0;
  for (int64_t i = 0; i < n; i++) {
    const span_t* s = &ws->sv.v[i];
    int32_t taxon;
    if (s->flag == AMBIGUOUS_FLAG) taxon = AMBIGUOUS_SPAN;
    else if (s->flag == MATE_PAIR_BORDER_FLAG) taxon = MATE_PAIR_BORDER;
    else { int32_t t; taxon = lib_lookup(L, s->minimizer, &t) ? t : TAXON_NONE; }
    ws->hits[i].taxon = taxon; ws->hits[i].count = s->kmers;
    if (s->distinct && taxon != TAXON_NONE) numDistinct++;
  }
  int m = merge_hits(ws->hits, (int)n, ws->merged);
  int32_t taxon = resolve_tree_hits(parents, ws->merged, m, confidence, &ws->mp);
  int classified = taxon != TAXON_NONE && numDistinct >= min_hit_groups;
  res->taxon = classified ? taxon : TAXON_NONE;
  res->classified = (uint8_t)classified;
  res->num_distinct = numDistinct;
  res->n_hits = m;
  /* lengthString (TaxonCounts.scala:114-121) */
  int border = -1;
  for (int i = 0; i < m; i++) if (ws->merged[i].taxon == MATE_PAIR_BORDER) { border = i; break; }
  int k = p->k;
  if (border < 0) { int s = 0; for (int i = 0; i < m; i++) s += ws->merged[i].count; res->len1 = s + (k - 1); }
  else {
    int s1 = 0, s2 = 0;
    for (int i = 0; i < border; i++) s1 += ws->merged[i].count;
    for (int i = border + 1; i < m; i++) s2 += ws->merged[i].count;
    res->len1 = s1 + (k - 1); res->len2 = s2 + (k - 1);
  }
  return 0;
}

/* Batch classify. bases2/off2 NULL for single-end. hit_off[i] gives each read's slot in hits_out
 * (caller sizes it by an upper bound); res[i].n_hits merged hits are written there.
 * Returns 0 or a negative error. Threads: OpenMP, `threads` <= 0 = all. */
SLKO_API int slko_classify_batch(const slko_params* p, const int32_t* parents, const slko_lib* L,
                                 const char* bases1, const int64_t* off1, const char* bases2, const int64_t* off2,
                                 int64_t n_reads, double confidence, int min_hit_groups,
                                 slko_result* res, const int64_t* hit_off, slko_hit* hits_out, int threads) {
  int err = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  #pragma omp parallel
  {
    classify_ws ws; memset(&ws, 0, sizeof(ws));
    #pragma omp for schedule(dynamic, 256)
    for (int64_t i = 0; i < n_reads; i++) {
      const char* n2 = bases2 ? bases2 + off2[i] : NULL;
      int64_t l2 = bases2 ? off2[i + 1] - off2[i] : 0;
      int rc = classify_fragment(p, parents, L, bases1 + off1[i], off1[i + 1] - off1[i], n2, l2,
                                 confidence, min_hit_groups, &ws, &res[i]);
      if (rc < 0) { err = rc; continue; }
      if (hits_out && hit_off) {
        int64_t cap = hit_off[i + 1] - hit_off[i];
        if (res[i].n_hits > cap) { err = -2; continue; }
        memcpy(hits_out + hit_off[i], ws.merged, sizeof(slko_hit) * (size_t)res[i].n_hits);
      }
    }
    free(ws.sv.v); scratch_free(&ws.sc); free(ws.mp.key); free(ws.mp.val); free(ws.hits); free(ws.merged);
  }
  return err;
}

SLKO_API int slko_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Synthetic workload generators (SURVEY.md section 8d shapes). Counter-based: every output byte is a pure
 * function of (seed, index), so the CUDA generator in the product library and this one can be compared
 * byte for byte. Not part of the reference.
 * ------------------------------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9e3779b97f4a7c15ull;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
static inline uint64_t rnd(uint64_t seed, uint64_t stream, uint64_t idx) {
  return splitmix64(splitmix64(seed * 0x100000001b3ull + stream) ^ idx);
}
static const char ACGT[4] = {'A', 'C', 'G', 'T'};

/* genome base at global position g (32 bases per random word). N-runs: positions are cut into blocks of
 * 65536; every block holds one N-run at a random offset with length 1..100 (~0.08% of positions). */
static inline char synth_genome_base(uint64_t seed, uint64_t g) {
  uint64_t blk = g >> 16, r = rnd(seed, 2, blk);
  {
    uint64_t st = r & 0xffff, ln = 1 + ((r >> 16) % 100);
    uint64_t o = g & 0xffff;
    if (o >= st && o < st + ln) return 'N';
  }
  uint64_t w = rnd(seed, 1, g >> 5);
  return ACGT[(w >> (2 * (g & 31))) & 3];
}
SLKO_API void slko_synth_genome(uint64_t seed, uint64_t start, uint64_t n, char* out) {
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n; i++) out[i] = synth_genome_base(seed, start + (uint64_t)i);
}
static inline char comp_char(char c) { switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; } }

/* read r of length L: 80% from the genome set (total G bases in n_genomes genomes of genome_len; half reverse
 * complemented; each base substituted with p = 1/100), 20% i.i.d. random; 1 read in 200 carries one N.
 * mate = 1 gives the read's mate of a read pair: same genome, 250 bases further along where the genome allows it,
 * the opposite strand, errors / N / random bases from streams of its own. */
static void synth_reads_mate(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                             uint64_t first_read, uint64_t n_reads, int L, int mate, char* out) {
  const uint64_t ms = mate ? 10u : 0u;
  #pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)n_reads; i++) {
    uint64_t r = first_read + (uint64_t)i;
    uint64_t h = rnd(rseed, 10, r);
    char* o = out + (uint64_t)i * (uint64_t)L;
    int from_genome = (h % 10) < 8;
    if (from_genome) {
      uint64_t g = (h >> 8) % n_genomes;
      uint64_t pos = rnd(rseed, 11, r) % (genome_len - (uint64_t)L + 1);
      if (mate && pos + 250 <= genome_len - (uint64_t)L) pos += 250;
      int rc = (int)((h >> 40) & 1) ^ (mate ? 1 : 0);
      uint64_t base = g * genome_len + pos;
      for (int j = 0; j < L; j++) {
        char c = rc ? comp_char(synth_genome_base(gseed, base + (uint64_t)(L - 1 - j))) : synth_genome_base(gseed, base + (uint64_t)j);
        uint64_t e = rnd(rseed, 12 + ms, r * 1024 + (uint64_t)j);
        if ((e % 100) == 0 && c != 'N') c = ACGT[((e >> 8) & 3)];
        o[j] = c;
      }
    } else {
      for (int j = 0; j < L; j++) {
        uint64_t w = rnd(rseed, 13 + ms, r * 32 + (uint64_t)(j >> 5));
        o[j] = ACGT[(w >> (2 * (j & 31))) & 3];
      }
    }
    uint64_t nn = rnd(rseed, 14 + ms, r);
    if ((nn % 200) == 0) o[(nn >> 8) % (uint64_t)L] = 'N';
  }
}
SLKO_API void slko_synth_reads(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                               uint64_t first_read, uint64_t n_reads, int L, char* out) {
  synth_reads_mate(gseed, rseed, n_genomes, genome_len, first_read, n_reads, L, 0, out);
}
SLKO_API void slko_synth_mates(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                               uint64_t first_read, uint64_t n_reads, int L, int mate, char* out) {
  synth_reads_mate(gseed, rseed, n_genomes, genome_len, first_read, n_reads, L, mate, out);
}
/* the OpenMP team size of every later call (torchrun exports OMP_NUM_THREADS=1, which the CPU arm of the bench must
 * not inherit) */
SLKO_API void slko_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}
