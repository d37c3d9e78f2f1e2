// slk_core.h -- per-thread bodies of the sm_100a kernels of the Slacken classify/build hot path.
//
// Everything here is written as per-thread functions whose few warp collectives (votes, ballots, warp barriers of
// the fused classifier) go through macros with a one-lane meaning, so that the very same source can also be compiled
// by g++ into tests/host_emulation (a TEST-ONLY harness that runs the kernel bodies one "lane" at a time on the CPU;
// it is never loaded by the product path, which fails loudly without a GPU).
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
//   Bracken weights ............... slacken/BrackenWeights.scala:46-284
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
#ifndef SLK_POOL
#define SLK_POOL 192     // span entries per tile; a WARP owns two tiles (filling / in flight) in shared memory
#endif
#define SLK_SHIST (SLK_POOL / 8)  // (taxon, k-mers) histogram pairs per lane that fit the idle staging area
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
  uint64_t* cells;     // n_buckets * 4 cells of (compressed key << 16 | dense taxon); 0 = empty; 128-byte aligned
  uint64_t n_buckets;  // one bucket = one 32-byte sector; always a multiple of 4 (four buckets = one 128-byte line)
  uint32_t mix_mul;    // 1; `world` for one shard of a table that is range-partitioned over `world` GPUs (slk_shard_of)
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

// slk_code for the four bytes of a word at once: *code8 = their 2-bit codes (byte 0 in bits 0-1, invalid characters as 0),
// *inv4 = one bit per byte that is none of A C G T U (either case). Byte-parallel arithmetic: an exact zero-byte test per
// accepted letter, and two multiplies that gather the per-byte fields (their partial products never overlap).
SLK_HD void slk_code4(uint32_t x, uint32_t* code8, uint32_t* inv4) {
  const uint32_t y = x & 0xDFDFDFDFu;   // upper case
  const uint32_t ta = y ^ 0x41414141u, tc = y ^ 0x43434343u, tg = y ^ 0x47474747u, tt = (y & 0xFEFEFEFEu) ^ 0x54545454u;
  // bit 7 of a byte of ((t & 0x7f) + 0x7f) | t is clear exactly when the byte of t is zero
  const uint32_t nz = (((ta & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | ta) & (((tc & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | tc) &
                      (((tg & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | tg) & (((tt & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | tt);
  const uint32_t invb = (nz >> 7) & 0x01010101u;
  uint32_t v = (x >> 1) & 0x03030303u;
  v ^= (v >> 1) & 0x01010101u;
  v &= ~(invb * 3u);
  *code8 = (v * 0x01041040u) >> 24;
  *inv4 = ((invb * 0x00204081u) >> 21) & 0xFu;
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
// the Kraken 2 default mask 0xffffffffcccccccc: keep the high word, gather bit pairs 2-3 of every low nibble
SLK_HD uint64_t slk_compress_fast(uint64_t x) {
  uint32_t y = ((uint32_t)x >> 2) & 0x33333333u;
  y = (y | (y >> 2)) & 0x0f0f0f0fu;
  y = (y | (y >> 4)) & 0x00ff00ffu;
  y = (y | (y >> 8)) & 0x0000ffffu;
  return ((x >> 32) << 16) | y;
}
SLK_HD uint64_t slk_compress_generic(const slk_scan_params& sp, uint64_t x) {
  x &= sp.sig_mask;
#pragma unroll
  for (int i = 0; i < 6; i++) {
    uint64_t t = x & sp.cmv[i];
    x = (x ^ t) | (t >> (1 << i));
  }
  return x;
}
SLK_HD uint64_t slk_compress(const slk_scan_params& sp, uint64_t x) {
  return sp.fast_compress ? slk_compress_fast(x) : slk_compress_generic(sp, x);
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
SLK_HD uint32_t slk_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
// Home bucket of a compressed key: a 128-byte line (n_buckets / 4 of them, fewer than 2^32 for any table that fits
// 180 GB) and one of its four 32-byte buckets. 32-bit arithmetic throughout (a 64-bit multiply is four instructions
// on the device): x, a murmur3-finalizer mix of all 48 key bits, picks the line by multiply-shift; y, a further
// mix that also separates keys with equal x, picks the bucket. The line is monotone in x: the build sorts its cells
// by x, so that equal keys meet AND the table is filled line after line instead of at random.
// x: a 32-bit mix of all 48 key bits (for a fixed high part a bijection of the low word)
SLK_HD uint32_t slk_key_mix(uint64_t ckey) {
  const uint32_t lo = (uint32_t)ckey, hi = (uint32_t)(ckey >> 32);
  uint32_t x = lo ^ (hi * 0x9E3779B1u);
  x *= 0x85EBCA6Bu; x ^= x >> 15;
  x *= 0xC2B2AE35u; x ^= x >> 13;
  return x;
}
// A table range-partitioned over `world` GPUs is one virtual table cut into `world` equal ranges of x: the 64-bit
// product x * world has the owner in its high word (slk_shard_of) and the position inside the owner's range in its low
// word, which takes the place of x when the owner picks the line (mix_mul = world; 1 for a whole table).
SLK_HD uint64_t slk_bucket_of(uint64_t ckey, const slk_table_view& tb) {
  const uint32_t x = slk_key_mix(ckey);
  const uint32_t y = (x ^ (uint32_t)(ckey >> 32)) * 0x27D4EB2Fu;
  return (uint64_t)slk_mulhi32(x * tb.mix_mul, (uint32_t)(tb.n_buckets >> 2)) * 4u + (y >> 30);
}
// Probe sequence. B200 answers a random 32-byte sector miss with the whole 128-byte line (measured: 127 B of DRAM
// traffic per random 32-byte gather, profiles/r01_probe_microbench.md) and is limited by the NUMBER of random
// line requests (~36 G/s), not by their bytes. So a collision chain first walks the other three buckets of its own
// line (cache hits), and only then moves on to the next line. `tries` = buckets already examined (>= 1).
SLK_HD uint64_t slk_next_bucket(uint64_t b, uint64_t tries, uint64_t n_buckets) {
  // bucket number t of a chain is (home line + t / 4, home sub-bucket XOR t % 4): the first step stays inside the
  // home bucket's 64-byte half line, the next two inside its 128-byte line. `b` is bucket number tries - 1.
  const uint64_t sub_home = (b ^ (tries - 1)) & 3ull, r = tries & 3ull;
  uint64_t nb = (b & ~3ull) | (sub_home ^ r);
  if (r == 0) { nb += 4; if (nb >= n_buckets) nb -= n_buckets; }
  return nb;
}
struct slk_bucket { uint64_t c0, c1, c2, c3; };
SLK_HD void slk_load_bucket(const slk_table_view& tb, uint64_t b, slk_bucket* o) {
#if defined(__CUDA_ARCH__)
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(tb.cells + b * 4);
  const ulonglong2 v0 = __ldg(p), v1 = __ldg(p + 1);
  o->c0 = v0.x; o->c1 = v0.y; o->c2 = v1.x; o->c3 = v1.y;
#else
  o->c0 = tb.cells[b * 4]; o->c1 = tb.cells[b * 4 + 1]; o->c2 = tb.cells[b * 4 + 2]; o->c3 = tb.cells[b * 4 + 3];
#endif
}
// A stored cell is (ckey << 16 | taxon) with taxon != 0, an empty cell is 0. Cell c holds the key iff its high word
// equals the key's and the low words differ in the taxon bits only; an empty cell that "matches" key 0 contributes
// taxon 0. At most one cell of a table holds a given key, so the taxa of the matching cells can simply be OR-ed.
SLK_HD bool slk_match_bucket(const slk_bucket& k, uint64_t ckey, uint32_t* dense) {
  const uint32_t khi = (uint32_t)(ckey >> 16), klo = (uint32_t)ckey << 16;
  const uint32_t l0 = (uint32_t)k.c0, l1 = (uint32_t)k.c1, l2 = (uint32_t)k.c2, l3 = (uint32_t)k.c3;
  const bool m0 = (uint32_t)(k.c0 >> 32) == khi && (l0 ^ klo) < 0x10000u;
  const bool m1 = (uint32_t)(k.c1 >> 32) == khi && (l1 ^ klo) < 0x10000u;
  const bool m2 = (uint32_t)(k.c2 >> 32) == khi && (l2 ^ klo) < 0x10000u;
  const bool m3 = (uint32_t)(k.c3 >> 32) == khi && (l3 ^ klo) < 0x10000u;
  const uint32_t d = ((m0 ? l0 : 0u) | (m1 ? l1 : 0u) | (m2 ? l2 : 0u) | (m3 ? l3 : 0u)) & 0xffffu;
  *dense = d;
  return d != 0 || k.c3 == 0;   // c3 == 0 <=> the bucket has an empty cell: cells fill in order, nothing is deleted
}
SLK_HD bool slk_probe_bucket(const slk_table_view& tb, uint64_t b, uint64_t ckey, uint32_t* dense) {
  slk_bucket k;
  slk_load_bucket(tb, b, &k);
  return slk_match_bucket(k, ckey, dense);
}
// Continues a probe whose first `tries` buckets (the last of them `b`) were full without a match.
static SLK_HD_NOINLINE uint32_t slk_probe_rest(const slk_table_view& tb, uint64_t b, uint64_t tries, uint64_t ckey) {
  for (; tries < tb.n_buckets; tries++) {
    b = slk_next_bucket(b, tries, tb.n_buckets);
    uint32_t dense;
    if (slk_probe_bucket(tb, b, ckey, &dense)) return dense;
  }
  return 0;
}
// Probe: returns the dense taxon of the key, 0 when absent (a left join miss -> Taxonomy.NONE).
SLK_HD uint32_t slk_probe(const slk_table_view& tb, uint64_t ckey) {
  const uint64_t b = slk_bucket_of(ckey, tb);
  uint32_t dense;
  if (slk_probe_bucket(tb, b, ckey, &dense)) return dense;
  return slk_probe_rest(tb, b, 1, ckey);
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
// Minimum of two values below 2^62. The scan is bound by the SM's integer pipe (ncu: ALU pipe 96 % busy, FMA pipe 11 %,
// profiles/r01_scan_pipes.md), and a 64-bit integer minimum is two compares and two selects there. Bit patterns below
// 2^62 are non-negative finite doubles whose order is the order of the integers, so ONE compare on the otherwise idle
// FP64 pipe replaces the two integer compares; the selects stay.
SLK_HD uint64_t slk_min62(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)a) < __longlong_as_double((long long)b) ? a : b;
#else
  return a < b ? a : b;
#endif
}

// Rolling m-mer (forward and reverse complement), its priority, and the minimum over the last W priorities = the
// minimizer of the k-mer window that ends at the base just pushed (the monotone deque of PosRankWindow.scala:47-74
// collapses to this for a fixed, small W; ties do not matter because super-mers merge on equal VALUE,
// MinSplitter.scala:200-201). The window minimum is kept as suffix minima: sfx[j] = min of the last j+1 priorities,
// updated in place, so no ring buffer has to be shifted.
// Everything here is RIGHT-aligned (the m-mer in bits 2m-1..0, at most 62 bits): the reference's left-aligned priority is
// this value << (64 - 2m), an order-preserving shift, applied where a minimizer leaves the scan (slk_scan_masks::fshift).
// Branch-free by design: an invalid character only resets `nvalid`; stale bits in fwd/rc/sfx are flushed by the
// k valid bases that must follow before the next window is reported.
struct slk_scan_masks {
  int fshift;        // left-aligned = right-aligned << fshift
  int rshift;        // 2m - 2: position of the first base of a right-aligned m-mer
  uint64_t mmask, xor_mask, sig_mask;   // the masks of slk_scan_params, right-aligned
  SLK_HD explicit slk_scan_masks(const slk_scan_params& sp)
      : fshift(sp.fshift), rshift(62 - sp.fshift), mmask(sp.mmask >> sp.fshift), xor_mask(sp.xor_mask >> sp.fshift),
        sig_mask(sp.sig_mask >> sp.fshift) {}
};
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
  // c: 0..3 for a base, 4 for anything else. Returns true when a full k-mer window of valid bases ends here;
  // *minv = the window's minimizer, right-aligned.
  SLK_HD bool push(uint32_t c, uint32_t k, int rshift, uint64_t mmask, uint64_t xor_mask, uint64_t sig_mask,
                   bool canonical, uint64_t* minv) {
    uint32_t b = c & 3u;
    fwd = ((fwd << 2) | (uint64_t)b) & mmask;
    rc = (rc >> 2) | ((uint64_t)(b ^ 3u) << rshift);
    nvalid = c < 4u ? nvalid + 1 : 0;
    uint64_t x = canonical ? slk_min62(fwd, rc) : fwd;
    x = (x ^ xor_mask) & sig_mask;
    uint64_t mn = x;
    if (W > 1) {
      mn = slk_min62(x, sfx[W > 1 ? W - 2 : 0]);
#pragma unroll
      for (int j = W - 2; j >= 1; j--) sfx[j] = slk_min62(x, sfx[j - 1]);
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
#define SLK_LABEL_PENDING 0xffffu  // a SEQ entry whose first bucket was full without a match (dense taxa are <= 65534)

struct slk_frag_result {
  int32_t taxon;          // raw taxon reported (0 when unclassified)
  uint32_t flags;         // SLK_F_*
  uint32_t kmers1, kmers2;  // sum of span k-mer counts per mate (lengthString = kmers + k - 1)
  uint32_t num_distinct;
  uint32_t n_hits;        // merged hits
  uint32_t n_probes;      // table lookups issued (= SEQ spans)
};

// A sink for merged hits that no longer fit the per-thread buffers (device: a worst-case block of the global hit
// buffer; emulation: a vector). `need` = upper bound of the hits this fragment can still produce.
struct slk_null_sink {
  SLK_HD void push(int32_t, int32_t, uint32_t) {}
  SLK_HD void reserve(uint32_t) {}
};

// LowestCommonAncestor.resolveTree (slacken/LowestCommonAncestor.scala:91-146) over a histogram H of
// (dense taxon -> k-mers) pairs in insertion order (fastutil Int2IntArrayMap): size(), at(i, &taxon, &count), count(taxon).
template <class H>
SLK_HD uint32_t slk_resolve_tree(const H& h, const slk_tax_view& tx, double confidence, int32_t total) {
  const double required = ceil(confidence * (double)total);
  const uint32_t n = h.size();
  uint32_t max_taxon = 0;
  int32_t max_score = 0;
  // one hit taxon (plus, possibly, misses): its path score is its own count, nothing to walk
  uint32_t nz = 0, only = 0;
  for (uint32_t i = 0; i < n; i++) {
    uint32_t t; int32_t v;
    h.at(i, &t, &v);
    if (t != 0) { nz++; only = t; }
  }
  if (nz == 1) {
    max_taxon = only;
  } else if (nz > 1) {
    for (uint32_t i = 0; i < n; i++) {
      uint32_t taxon; int32_t v;
      h.at(i, &taxon, &v);
      uint32_t node = taxon;
      int32_t score = 0;
      while (node != 0) { score += h.count(node); node = tx.parent[node]; }
      if (score > max_score) { max_taxon = taxon; max_score = score; }
      else if (score == max_score) max_taxon = slk_lca(tx, max_taxon, taxon);
    }
  }
  max_score = h.count(max_taxon);
  while (max_taxon != 0 && (double)max_score < required) {
    max_score = 0;
    for (uint32_t i = 0; i < n; i++) {
      uint32_t t; int32_t v;
      h.at(i, &t, &v);
      if (slk_has_ancestor(tx, t, max_taxon)) max_score += v;
    }
    if ((double)max_score >= required) return max_taxon;
    max_taxon = tx.parent[max_taxon];
  }
  return max_taxon;
}
// the first SLK_SHIST pairs of a lane's histogram, in the idle bucket staging area of the fast store
template <class Entries>
struct slk_fast_hist {
  Entries& ent;
  uint32_t n;
  SLK_HD uint32_t size() const { return n; }
  SLK_HD void at(uint32_t i, uint32_t* t, int32_t* v) const { ent.hist_get(i, t, v); }
  SLK_HD int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < n; i++) {
      uint32_t ti; int32_t vi;
      ent.hist_get(i, &ti, &vi);
      if (ti == t) return vi;
    }
    return 0;
  }
  SLK_HD bool add(uint32_t t, int32_t c) {   // false: no room (the caller falls back to the per-thread arrays)
    for (uint32_t i = 0; i < n; i++) {
      uint32_t ti; int32_t vi;
      ent.hist_get(i, &ti, &vi);
      if (ti == t) { ent.hist_set(i, t, vi + c); return true; }
    }
    if (n == SLK_SHIST) return false;
    ent.hist_set(n, t, c); n++;
    return true;
  }
};

// Fast store of a WARP (shared memory on the device, plain arrays in the one-lane emulation):
//  * two tiles (one being filled, one whose lookups are in flight) of SLK_POOL span entries each. The lanes of a warp
//    append to the tile in step order; an entry is key (64 bit), meta (count | type << 14) and the slot of the same
//    lane's next entry;
//  * ONE staging area of SLK_POOL x 32 bytes: slot s receives the table bucket of entry s of the tile in flight through
//    an ASYNCHRONOUS copy (the tile being filled needs none yet, so both tiles can be 1.5x larger for the same memory);
//  * per lane, the first SLK_SHITS merged hits of its fragment;
//  * once the scan is over, the staging area is reused as each lane's first SLK_SHIST (taxon, k-mers) histogram pairs.
struct slk_store_local {
  uint64_t keys[2][SLK_POOL];
  uint16_t metas[2][SLK_POOL];
  uint8_t nexts[2][SLK_POOL], pend[2][SLK_POOL];
  uint64_t stage[SLK_POOL][4];
  int32_t hl[SLK_SHITS], hc[SLK_SHITS];
  SLK_HD uint32_t tile(uint32_t t) const { return t; }   // tile handle (the device's is a shared-space address)
  SLK_HD void put(uint32_t t, uint32_t s, uint64_t k, uint32_t m) { keys[t][s] = k; metas[t][s] = (uint16_t)m; }
  SLK_HD uint64_t key(uint32_t t, uint32_t s) const { return keys[t][s]; }
  SLK_HD void set_key(uint32_t t, uint32_t s, uint64_t k) { keys[t][s] = k; }
  SLK_HD uint32_t meta(uint32_t t, uint32_t s) const { return metas[t][s]; }
  SLK_HD void set_next(uint32_t t, uint32_t s, uint32_t n) { nexts[t][s] = (uint8_t)n; }
  SLK_HD uint32_t next(uint32_t t, uint32_t s) const { return nexts[t][s]; }
  SLK_HD void set_pending(uint32_t t, uint32_t q, uint32_t s) { pend[t][q] = (uint8_t)s; }
  SLK_HD uint32_t pending(uint32_t t, uint32_t q) const { return pend[t][q]; }
  SLK_HD void fetch(uint32_t s, const uint64_t* src) { for (int i = 0; i < 4; i++) stage[s][i] = src[i]; }
  SLK_HD void bucket(uint32_t s, slk_bucket* o) const {
    o->c0 = stage[s][0]; o->c1 = stage[s][1]; o->c2 = stage[s][2]; o->c3 = stage[s][3];
  }
  SLK_HD void commit(uint32_t) const {}      // n: buckets this lane requested since the last commit
  SLK_HD void wait_all(uint32_t) const {}    // parity: number of commits so far, mod 2
  SLK_HD void set_hit(uint32_t i, int32_t label, int32_t count) { hl[i] = label; hc[i] = count; }
  SLK_HD void get_hit(uint32_t i, int32_t* label, int32_t* count) const { *label = hl[i]; *count = hc[i]; }
  SLK_HD void hist_set(uint32_t i, uint32_t t, int32_t v) { (&stage[0][0])[i] = ((uint64_t)(uint32_t)v << 32) | t; }
  SLK_HD void hist_get(uint32_t i, uint32_t* t, int32_t* v) const { uint64_t x = (&stage[0][0])[i]; *t = (uint32_t)x; *v = (int32_t)(x >> 32); }
};

// One fragment (a read or a read pair) per thread, end to end: scan -> span entries -> table lookups -> merged
// hits -> resolveTree, with the lookups decoupled from the thread that needs them:
//  * the lanes of a warp append their span entries to a tile of the warp. When the tile is full the warp closes it:
//    the buckets of all its SEQ entries are requested with asynchronous global->shared copies (32 entries per
//    round, every lane busy whatever its own fragment looks like) and are only examined when the NEXT tile closes,
//    after the warp has scanned on. A warp keeps up to SLK_POOL random line requests in flight without holding a
//    register for them and without waiting, and the B200's random-request ceiling (~36 G lines/s,
//    profiles/r01_probe_microbench.md) is reached with 12-16 warps per SM;
//  * matching a landed bucket against its key is again done 32 entries per round; only the merge of a lane's own
//    hits (adjacent equal taxa, numDistinct, k-mer totals) walks that lane's entries in order.
template <int W, class Sink, class Entries>
struct slk_frag_classifier {
  const slk_table_view tb;
  const slk_tax_view tx;
  Sink& sink;
  Entries ent;

  // totals of the fragment (written once, at the end of the scan; resolve() and the kernel epilogue read them)
  uint32_t kmers[2], nd, nprobes;
  // merged hits (dense labels): the first SLK_SHITS in the fast store, then xh_*, then spilled through the sink
  uint32_t nh;            // buffered
  uint32_t nh_spilled;    // already pushed to the sink (0 for all but very long reads)
  int32_t xh_label[SLK_XHITS];
  int32_t xh_count[SLK_XHITS];
  // histogram of the slow path (a fragment that spilled hits while its tiles were busy, or hit > SLK_SHIST taxa)
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t nk;
  bool overflow;

  SLK_HD slk_frag_classifier(const slk_table_view& tb_, const slk_tax_view& tx_, Sink& s, const Entries& e)
      : tb(tb_), tx(tx_), sink(s), ent(e) {}

  SLK_HD uint32_t size() const { return nk; }
  SLK_HD void at(uint32_t i, uint32_t* t, int32_t* v) const { *t = hk[i]; *v = hv[i]; }
  SLK_HD int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) return hv[i];
    return 0;
  }
  SLK_HD void hist_add(uint32_t t, int32_t c) {
    for (uint32_t i = 0; i < nk; i++)
      if (hk[i] == t) { hv[i] += c; return; }
    if (nk == SLK_KMAX) { overflow = true; return; }
    hk[nk] = t; hv[nk] = c; nk++;
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
  // the slow path of the final step: histogram in the per-thread arrays
  SLK_HD_NOINLINE uint32_t resolve_slow(double confidence) {
    fold_hits();   // the buffered hits (a fragment that spilled has none left and was folded on the way)
    return slk_resolve_tree(*this, tx, confidence, (int32_t)(kmers[0] + kmers[1]));
  }

  // One fragment end to end (paired == false: r2 is ignored). Every lane of a warp must call it (lanes without a
  // fragment pass empty reads): tiles are filled, closed and matched by the warp as a whole.
  // Scan: Supermers.splitByAmbiguity/splitFragment (slacken/Supermers.scala:113-189): valid runs >= k are cut into
  // super-mers (runs of k-mer windows with equal minimizer), runs of >= k ambiguous characters become one
  // AMBIGUOUS span of len-(k-1), anything shorter vanishes; the mates are separated by a MATE_PAIR_BORDER span.
  template <bool PACKED, bool CANON>
  SLK_HD void run(const slk_scan_params& sp, const slk_read_src& r1, const slk_read_src& r2, bool paired,
                  double confidence, int32_t min_hit_groups, slk_frag_result& r) {
    // The running state of the merge side lives in registers for the whole scan; the members (which sit in local
    // memory because spill() and resolve() are out of line) are only written at the end and around a spill.
    uint64_t l_last = 0;
    bool l_have_last = false, l_have_cur = false;
    int32_t l_label = 0, l_count = 0;
    uint32_t l_mate = 0, l_k0 = 0, l_k1 = 0, l_nd = 0, l_np = 0, l_nh = 0;
    nh = 0; nh_spilled = 0; nk = 0; overflow = false;
    bool l_spilled = false;
    const uint32_t k = (uint32_t)sp.k, km1 = k - 1;
    const slk_scan_masks sm(sp);   // right-aligned masks: run keys are right-aligned until the issue pass compresses them
    const int fshift = sm.fshift, rshift = sm.rshift;
    const uint64_t mmask = sm.mmask, xor_mask = sm.xor_mask, sig_mask = sm.sig_mask;
    const bool fast = sp.fast_compress != 0;
    const int32_t border_cnt = -(sp.k - 1);
    // upper bound of the merged hits of the fragment (for the sink's spill allocation)
    const uint32_t max_hits = (r1.len > km1 ? r1.len - km1 : 0) + (paired ? (r2.len > km1 ? r2.len - km1 : 0) + 1 : 0) + 2;
    uint32_t n_cur = 0, n_prev = 0;        // entries of the warp in the tile being filled / in flight (warp-uniform)
    uint32_t head_cur = 0, tail_cur = 0, cnt_cur = 0;   // this lane's own entries in the tile being filled
    uint32_t head_prev = 0, cnt_prev = 0;               // ... and in the tile in flight
    bool any = false;                      // did the fragment yield any span at all
    uint32_t n_commits = 0;                // closes so far (the parity of the store's completion barrier)
#if defined(__CUDA_ARCH__)
    const Entries ent = this->ent;   // shared-space addresses, in registers
    const slk_table_view tb = this->tb;
    const uint32_t lane = threadIdx.x & 31u, lanes_below = (1u << lane) - 1u;
#define SLK_LANES 32u
#define SLK_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define SLK_WARP_ALL(p) __all_sync(0xffffffffu, (p))
#define SLK_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#define SLK_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define SLK_POPC(x) ((uint32_t)__popc(x))
#define SLK_SYNCWARP() __syncwarp()
#else
    Entries& ent = this->ent;
    const slk_table_view& tb = this->tb;
    const uint32_t lane = 0, lanes_below = 0;
#define SLK_LANES 1u
#define SLK_WARP_ANY(p) (p)
#define SLK_WARP_ALL(p) (p)
#define SLK_WARP_MAX(x) (x)
#define SLK_BALLOT(p) ((p) ? 1u : 0u)
#define SLK_POPC(x) ((uint32_t)__builtin_popcount(x))
#define SLK_SYNCWARP() do {} while (0)
#endif
    const uint32_t pool_limit = SLK_POOL - SLK_LANES;   // a tile with more entries cannot take another step
    const uint32_t ent_tile0 = ent.tile(0), ent_tile1 = ent.tile(1);
    uint32_t cur = ent_tile0;              // handle of the tile being filled

    // TaxonCounts.fromHits (slacken/TaxonCounts.scala:31-48) has already merged adjacent equal taxa: one store
    auto push_hit = [&](int32_t label, int32_t count) {
      if (l_nh == SLK_SHITS + SLK_XHITS) { nh = l_nh; spill(max_hits - l_nh); l_nh = 0; l_spilled = true; }   // spill allocates need + nh + 2
      if (l_nh < SLK_SHITS) ent.set_hit(l_nh, label, count);
      else { xh_label[l_nh - SLK_SHITS] = label; xh_count[l_nh - SLK_SHITS] = count; }
      l_nh++;
    };
    // Closes the tile being filled. The whole warp comes here together.
    //  1. the buckets of the PREVIOUS tile, requested one tile ago, have landed in the staging area;
    //  2. match pass over the previous tile, 32 entries per round: spanToHit's join (slacken/KeyValueIndex.scala:
    //     176-185). The dense taxon (0 = miss -> Taxonomy.NONE) goes to the top 16 bits of the entry's key slot.
    //     The few entries whose bucket was full without a match are collected in the tile's pending list;
    //  3. issue pass over the tile just filled, 32 entries per round: compress the key, request its bucket (the
    //     staging area is free again: step 2 has read all of it);
    //  4. pending pass, 32 entries per round: the chains go on with ordinary loads; their next bucket is in the same
    //     128-byte line, which the first request already brought into the caches;
    //  5. merge pass: every lane walks ITS entries of the previous tile in span order: numDistinct
    //     (slacken/Classifier.scala:94), k-mer totals, TaxonCounts.fromHits.
    // A lane only ever waits for copies it issued itself: slot s is requested and matched by lane s % 32.
    auto close = [&]() {
      any = any || cnt_cur != 0;
      const uint32_t prev = cur ^ ent_tile0 ^ ent_tile1;
      SLK_SYNCWARP();   // the entries other lanes appended are visible
      uint32_t n_pend = 0;
      ent.wait_all(n_commits & 1u);
#pragma unroll 2
      for (uint32_t s0 = 0; s0 < n_prev; s0 += SLK_LANES) {
        const uint32_t s = s0 + lane;
        bool pend = false;
        if (s < n_prev && (ent.meta(prev, s) >> 14) == SLK_E_SEQ) {
          const uint64_t ck = ent.key(prev, s);
          slk_bucket bk;
          ent.bucket(s, &bk);
          uint32_t dense;
          pend = !slk_match_bucket(bk, ck, &dense);
          if (!pend) ent.set_key(prev, s, ck | ((uint64_t)dense << 48));
        }
        const uint32_t bal = SLK_BALLOT(pend);
        if (pend) ent.set_pending(prev, n_pend + SLK_POPC(bal & lanes_below), s);
        n_pend += SLK_POPC(bal);
      }
      SLK_SYNCWARP();   // the pending list is complete, and nobody reads the staging area any more
      uint32_t n_fetched = 0;
#pragma unroll 2
      for (uint32_t s = lane; s < n_cur; s += SLK_LANES) {
        if ((ent.meta(cur, s) >> 14) == SLK_E_SEQ) {
          const uint64_t key = ent.key(cur, s) << fshift;
          const uint64_t ck = fast ? slk_compress_fast(key) : slk_compress_generic(sp, key);
          ent.set_key(cur, s, ck);
          ent.fetch(s, tb.cells + slk_bucket_of(ck, tb) * 4);
          n_fetched++;
        }
      }
      ent.commit(n_fetched);
      n_commits++;
      for (uint32_t q = lane; q < n_pend; q += SLK_LANES) {
        const uint32_t s = ent.pending(prev, q);
        const uint64_t ck = ent.key(prev, s);
        const uint32_t dense = slk_probe_rest(tb, slk_bucket_of(ck, tb), 1, ck);
        ent.set_key(prev, s, ck | ((uint64_t)dense << 48));
      }
      SLK_SYNCWARP();   // all labels of the previous tile are visible to the lanes that own the entries
      uint32_t s = head_prev;
      for (uint32_t q = 0; q < cnt_prev; q++) {
        const uint32_t meta = ent.meta(prev, s), type = meta >> 14, cnt = meta & SLK_E_CNT_MAX;
        int32_t label, hcnt = (int32_t)cnt;
        if (type == SLK_E_SEQ) {
          const uint64_t kl = ent.key(prev, s), ck = kl & 0xffffffffffffull;
          const uint32_t dense = (uint32_t)(kl >> 48);
          l_np++;
          l_nd += ((!l_have_last || ck != l_last) && dense != 0) ? 1u : 0u;
          l_last = ck; l_have_last = true;
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
          if (l_have_cur) push_hit(l_label, l_count);
          l_label = label; l_count = hcnt; l_have_cur = true;
        }
        if (type == SLK_E_BORDER) l_mate = 1;
        s = ent.next(prev, s);
      }
      SLK_SYNCWARP();   // nobody appends to the previous tile before everybody has left it
      n_prev = n_cur; head_prev = head_cur; cnt_prev = cnt_cur;
      n_cur = 0; cnt_cur = 0; cur = prev;
    };
    // Appends one entry per lane with `emit` set, in lane order. All lanes call it; the tile has room for all of
    // them (n_cur <= pool_limit). The caller closes the tile when the return value says it is full.
    auto append = [&](bool emit, uint64_t key, uint32_t meta) -> bool {
      const uint32_t bal = SLK_BALLOT(emit);
      if (emit) {
        const uint32_t s = n_cur + SLK_POPC(bal & lanes_below);
        ent.put(cur, s, key, meta);
        ent.set_next(cur, cnt_cur ? tail_cur : s, s);   // the first entry links to itself; its link is set by the second
        head_cur = cnt_cur ? head_cur : s;
        tail_cur = s; cnt_cur++;
      }
      n_cur += SLK_POPC(bal);
      return n_cur > pool_limit;
    };

#pragma unroll 1
    for (int mt = 0; mt < (paired ? 2 : 1); mt++) {  // one copy of the scan loop serves both mates
      if (mt && append(true, 0, SLK_E_BORDER << 14)) close();
      const slk_read_src& src = mt ? r2 : r1;
      const uint32_t len = src.len;
      slk_scanner<W> sc;
      sc.reset();
      uint64_t run_key = 0;
      uint32_t run_cnt = 0, ninv = 0, amb_cnt = 0;
      bool in_run = false;
      // One character (c = 0..3 for a base, 4 for anything else) of the lanes with `act` set; at most one entry
      // each. Returns true when the tile is full.
      auto step = [&](uint32_t c, bool act) -> bool {
        bool valid = true, window_ok = false;
        uint64_t mn = 0;
        if (act) {
          valid = c < 4u;
          window_ok = sc.push(c, k, rshift, mmask, xor_mask, sig_mask, CANON, &mn);
          ninv = valid ? 0u : ninv + 1u;
        }
        const bool same = in_run && mn == run_key && run_cnt < SLK_E_CNT_MAX;
        const bool start_new = window_ok && !same;
        const bool emit_seq = in_run && (start_new || !valid);  // the open super-mer ends here
        const bool emit_amb = act && amb_cnt != 0 && (valid || amb_cnt == SLK_E_CNT_MAX);  // an ambiguous stretch ended
        const uint32_t amb_meta = amb_cnt | (SLK_E_AMB << 14);
        if (act) {
          amb_cnt = (valid || emit_amb) ? 0u : amb_cnt;
          amb_cnt += (!valid && ninv >= k) ? 1u : 0u;
        }
        const uint64_t ek = emit_seq ? run_key : 0ull;
        const uint32_t em = emit_seq ? run_cnt : amb_meta;
        run_cnt = start_new ? 1u : run_cnt + ((window_ok && same) ? 1u : 0u);
        run_key = start_new ? mn : run_key;
        in_run = valid && (in_run || start_new);
        return append(emit_seq || emit_amb, ek, em);
      };
      // The same for a base (b = 0..3) of a block in which no lane of the warp has an ambiguous character or an
      // open ambiguous stretch: no validity bookkeeping at all. GUARD: some lanes may be outside their read.
      auto fstep = [&](uint32_t b, bool act) -> bool {
        bool window_ok = false;
        uint64_t mn = 0;
        if (act) window_ok = sc.push(b, k, rshift, mmask, xor_mask, sig_mask, CANON, &mn);
        const bool same = in_run && mn == run_key && run_cnt < SLK_E_CNT_MAX;
        const bool start_new = window_ok && !same;
        const bool emit = in_run && start_new;
        const uint64_t ek = run_key;
        const uint32_t em = run_cnt;
        run_cnt = start_new ? 1u : run_cnt + (window_ok ? 1u : 0u);
        run_key = start_new ? mn : run_key;
        in_run = in_run || start_new;
        return append(emit, ek, em);
      };
      // Bases [i0, i1) of a block of 2-bit codes `cw` (base i in bits 2i, 2i+1) with ambiguity bits `mw`. All lanes
      // of a warp run the same number of iterations. The scan loops leave when the tile is full, so that there is
      // one copy of close() here, and come back to where they were.
      auto scan_block = [&](uint64_t cw, uint32_t mw, uint32_t i0, uint32_t i1) {
        const bool dirty = SLK_WARP_ANY(mw != 0u || ninv != 0u || amb_cnt != 0u);
        const uint32_t i1w = SLK_WARP_MAX(i1);
        const bool whole = !dirty && SLK_WARP_ALL(i0 == 0u && i1 == i1w);   // every lane has all i1w bases
        uint32_t i = 0;
        while (i < i1w) {
          bool full = false;
          if (whole) {
            while (i + 4u <= i1w && !full) {   // four bases per round, no per-lane guards
              const uint32_t g = (uint32_t)(cw >> (2u * i)) & 0xffu;
              if (fstep(g & 3u, true)) { i += 1u; full = true; continue; }
              if (fstep((g >> 2) & 3u, true)) { i += 2u; full = true; continue; }
              if (fstep((g >> 4) & 3u, true)) { i += 3u; full = true; continue; }
              full = fstep(g >> 6, true);
              i += 4u;
            }
            while (i < i1w && !full) { full = fstep((uint32_t)(cw >> (2u * i)) & 3u, true); i++; }
          } else if (!dirty) {
            while (i < i1w && !full) {
              full = fstep((uint32_t)(cw >> (2u * i)) & 3u, i >= i0 && i < i1); i++;
            }
          } else {
            while (i < i1w && !full) {
              full = step(((uint32_t)(cw >> (2u * i)) & 3u) | (((mw >> i) & 1u) << 2), i >= i0 && i < i1);
              i++;
            }
          }
          if (full) close();
        }
      };
      if (PACKED) {
        const uint32_t nblk = (len + 31u) >> 5;
        const uint32_t nblk_w = SLK_WARP_MAX(nblk);
        uint64_t cw_next = 0;
        uint32_t mw_next = 0;
#if defined(__CUDA_ARCH__)
#define SLK_LOAD_BLOCK(b) do { cw_next = __ldg(src.codes + (b)); mw_next = __ldg(src.mask + (b)); } while (0)
#else
#define SLK_LOAD_BLOCK(b) do { cw_next = src.codes[b]; mw_next = src.mask[b]; } while (0)
#endif
        if (nblk) SLK_LOAD_BLOCK(0);
        for (uint32_t b = 0; b < nblk_w; b++) {   // block b+1 is requested before block b is scanned
          const uint64_t cw = cw_next;
          const uint32_t mw = mw_next;
          uint32_t nb = 0;
          if (b < nblk) { nb = len - 32u * b; nb = nb > 32u ? 32u : nb; }
          cw_next = 0; mw_next = 0;
          if (b + 1 < nblk) SLK_LOAD_BLOCK(b + 1);
          scan_block(cw, mw, 0u, nb);
        }
#undef SLK_LOAD_BLOCK
      } else {
        // 16-byte chunks over [s, s+len), aligned for the device's vector loads. The buffer is readable up to the
        // next 16-byte boundary (library-owned and cudaMalloc'ed buffers are); bytes outside the read are skipped.
        const uintptr_t a0 = reinterpret_cast<uintptr_t>(src.ascii);
        const uintptr_t abase = a0 & ~(uintptr_t)15;
        const int32_t lo0 = (int32_t)(a0 - abase), total = lo0 + (int32_t)len;  // byte range [lo0, total) from abase
        const int32_t total_w = (int32_t)SLK_WARP_MAX((uint32_t)(len != 0 ? total : 0));
        for (int32_t cb = 0; cb < total_w; cb += 16) {
          uint64_t cw = 0;
          uint32_t mw = 0, i0 = 0, i1 = 0;
          if (cb < total && len != 0) {
            i0 = lo0 > cb ? (uint32_t)(lo0 - cb) : 0u;
            i1 = total - cb > 16 ? 16u : (uint32_t)(total - cb);
#if defined(__CUDA_ARCH__)
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(abase + cb));
            const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (uint32_t i = 0; i < 16; i++) {
              const uint32_t c = slk_code((wd[i >> 2] >> (8u * (i & 3u))) & 0xffu);
              cw |= (uint64_t)(c & 3u) << (2u * i);
              mw |= (c >> 2) << i;
            }
#else
            for (uint32_t i = i0; i < i1; i++) {
              const uint32_t c = slk_code(reinterpret_cast<const uint8_t*>(abase)[cb + i]);
              cw |= (uint64_t)(c & 3u) << (2u * i);
              mw |= (c >> 2) << i;
            }
#endif
          }
          scan_block(cw, mw, i0, i1);
        }
      }
      // mate end: at most one pending entry (a run and an ambiguous stretch cannot both be open)
      if (append(in_run || amb_cnt != 0, in_run ? run_key : 0ull, in_run ? run_cnt : (amb_cnt | (SLK_E_AMB << 14)))) close();
    }
#pragma unroll 1
    for (int e = 0; e < 2; e++) close();   // the first requests the last tile and merges the one before it
    if (l_have_cur) push_hit(l_label, l_count);
    // The number of buffered hits is final: the sink may reserve their place in the output now (on the device one
    // atomic per warp, whose latency the resolve step below hides). All lanes call it.
    sink.reserve(l_spilled ? 0u : l_nh);
    nh = l_nh; kmers[0] = l_k0; kmers[1] = l_k1; nd = l_nd; nprobes = l_np;
    // totalKmers: ambiguous spans count, the border does not (slacken/TaxonCounts.scala:83-87)
    uint32_t taxon;
    bool fast_done = false;
    if (!l_spilled) {   // nearly every fragment: the histogram fits the (now idle) staging area of the fast store
      Entries hstore = ent;   // device: four shared-space addresses; emulation: see below
      slk_fast_hist<Entries> fh{hstore, 0u};
      bool fits = true;
      for (uint32_t i = 0; i < l_nh && fits; i++) {
        int32_t l, c;
        buffered_hit(i, &l, &c);
        if (l >= 0) fits = fh.add((uint32_t)l, c);
      }
      if (fits) { taxon = slk_resolve_tree(fh, tx, confidence, (int32_t)(l_k0 + l_k1)); fast_done = true; }
    } else {
      spill(0);   // a fragment that went to the sink keeps all its hits there
    }
    if (!fast_done) taxon = resolve_slow(confidence);
    bool classified = taxon != 0 && nd >= (uint32_t)min_hit_groups;  // slacken/Classifier.scala:446
    r.taxon = classified ? tx.raw[taxon] : 0;
    r.flags = (classified ? SLK_F_CLASSIFIED : 0u) | (any ? SLK_F_HAS_SPAN : 0u) | (overflow ? SLK_F_OVERFLOW : 0u);
    r.kmers1 = kmers[0]; r.kmers2 = kmers[1];
    r.num_distinct = nd; r.n_hits = nh + nh_spilled; r.n_probes = nprobes;
#undef SLK_LANES
#undef SLK_WARP_ANY
#undef SLK_WARP_ALL
#undef SLK_WARP_MAX
#undef SLK_BALLOT
#undef SLK_POPC
#undef SLK_SYNCWARP
  }
};

// ------------------------------------------------------------------------------------------------ split path
// The same classification as slk_frag_classifier, cut at the table lookup, for libraries that do not fit one GPU:
// the query side scans its reads into span words, the span keys travel to the GPU that owns their hash range and
// come back as taxa, and the query side merges and resolves (the join of
// slacken/Classifier.scala:84 done by an all-to-all instead of a shuffle).
// A span word is (compressed key << 16 | type << 14 | k-mer count); AMBIGUOUS and BORDER spans carry key 0.
#define SLK_SPAN_TYPE(w) ((uint32_t)((w) >> 14) & 3u)
#define SLK_SPAN_CNT(w) ((uint32_t)(w) & SLK_E_CNT_MAX)
#define SLK_SPAN_KEY(w) ((uint64_t)(w) >> 16)

// Scan of one mate: calls emit(span word) for every span, in order (the per-character logic of
// slk_frag_classifier::run, without tiles).
template <int W, class Emit>
SLK_HD void slk_scan_spans(const slk_scan_params& sp, const uint8_t* s, uint32_t len, Emit&& emit) {
  const uint32_t k = (uint32_t)sp.k;
  const slk_scan_masks sm(sp);
  const int fshift = sm.fshift, rshift = sm.rshift;
  const uint64_t mmask = sm.mmask, xor_mask = sm.xor_mask, sig_mask = sm.sig_mask;
  const bool canonical = sp.canonical != 0;
  slk_scanner<W> sc;
  sc.reset();
  uint64_t run_key = 0;
  uint32_t run_cnt = 0, ninv = 0, amb_cnt = 0;
  bool in_run = false;
  slk_for_each_byte(s, len, [&](uint32_t ch) {
    const uint32_t c = slk_code(ch);
    const bool valid = c < 4u;
    uint64_t mn;
    const bool window_ok = sc.push(c, k, rshift, mmask, xor_mask, sig_mask, canonical, &mn);
    ninv = valid ? 0u : ninv + 1u;
    const bool same = in_run && mn == run_key && run_cnt < SLK_E_CNT_MAX;
    const bool start_new = window_ok && !same;
    if (in_run && (start_new || !valid)) emit((slk_compress(sp, run_key << fshift) << 16) | (SLK_E_SEQ << 14) | run_cnt);
    if (amb_cnt != 0 && (valid || amb_cnt == SLK_E_CNT_MAX)) { emit((uint64_t)((SLK_E_AMB << 14) | amb_cnt)); amb_cnt = 0; }
    amb_cnt = valid ? 0u : amb_cnt;
    amb_cnt += (!valid && ninv >= k) ? 1u : 0u;
    run_cnt = start_new ? 1u : run_cnt + ((window_ok && same) ? 1u : 0u);
    run_key = start_new ? mn : run_key;
    in_run = valid && (in_run || start_new);
  });
  if (in_run) emit((slk_compress(sp, run_key << fshift) << 16) | (SLK_E_SEQ << 14) | run_cnt);
  if (amb_cnt) emit((uint64_t)((SLK_E_AMB << 14) | amb_cnt));
}
// ... of a fragment: mate 1, the MATE_PAIR_BORDER pseudo-span, mate 2 (slacken/Supermers.scala:49-97)
template <int W, class Emit>
SLK_HD void slk_scan_fragment_spans(const slk_scan_params& sp, const uint8_t* s1, uint32_t len1, const uint8_t* s2,
                                    uint32_t len2, bool paired, Emit&& emit) {
  slk_scan_spans<W>(sp, s1, len1, emit);
  if (paired) {
    emit((uint64_t)(SLK_E_BORDER << 14));
    slk_scan_spans<W>(sp, s2, len2, emit);
  }
}

// histogram of the split path's resolve step: per-thread arrays
struct slk_array_hist {
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t n;
  bool overflow;
  SLK_HD uint32_t size() const { return n; }
  SLK_HD void at(uint32_t i, uint32_t* t, int32_t* v) const { *t = hk[i]; *v = hv[i]; }
  SLK_HD int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < n; i++)
      if (hk[i] == t) return hv[i];
    return 0;
  }
  SLK_HD void add(uint32_t t, int32_t c) {
    for (uint32_t i = 0; i < n; i++)
      if (hk[i] == t) { hv[i] += c; return; }
    if (n == SLK_KMAX) { overflow = true; return; }
    hk[n] = t; hv[n] = c; n++;
  }
};

// Merge + resolve of one fragment from its span words and the dense taxon of every span (0 = miss; ignored for
// AMBIGUOUS / BORDER spans): the merge pass and the final step of slk_frag_classifier::run. hit(label, count) is
// called for every merged hit in order, label = dense taxon or SLK_AMBIGUOUS_SPAN / SLK_MATE_PAIR_BORDER.
template <class Hit>
SLK_HD void slk_resolve_spans(const slk_tax_view& tx, int32_t k, const uint64_t* spans, const uint16_t* dense, uint32_t n_spans,
                              double confidence, int32_t min_hit_groups, Hit&& hit, slk_frag_result& r) {
  slk_array_hist h;
  h.n = 0; h.overflow = false;
  uint64_t last = 0;
  bool have_last = false, have_cur = false;
  int32_t cur_label = 0, cur_count = 0;
  uint32_t mate = 0, k0 = 0, k1 = 0, nd = 0, np = 0, nh = 0;
  auto flush = [&]() {
    hit(cur_label, cur_count);
    nh++;
    if (cur_label >= 0) h.add((uint32_t)cur_label, cur_count);   // TaxonCounts.toMap skips AMBIGUOUS / BORDER
  };
  for (uint32_t i = 0; i < n_spans; i++) {
    const uint64_t w = spans[i];
    const uint32_t type = SLK_SPAN_TYPE(w), cnt = SLK_SPAN_CNT(w);
    int32_t label, hcnt = (int32_t)cnt;
    if (type == SLK_E_SEQ) {
      const uint64_t ck = SLK_SPAN_KEY(w);
      const uint32_t d = dense[i];
      np++;
      nd += ((!have_last || ck != last) && d != 0) ? 1u : 0u;
      last = ck; have_last = true;
      label = (int32_t)d;
    } else if (type == SLK_E_AMB) {
      label = SLK_AMBIGUOUS_SPAN;
    } else {
      label = SLK_MATE_PAIR_BORDER; hcnt = -(k - 1);
    }
    if (type != SLK_E_BORDER) { if (mate) k1 += cnt; else k0 += cnt; }
    if (have_cur && label == cur_label) cur_count += hcnt;
    else {
      if (have_cur) flush();
      cur_label = label; cur_count = hcnt; have_cur = true;
    }
    if (type == SLK_E_BORDER) mate = 1;
  }
  if (have_cur) flush();
  const uint32_t taxon = slk_resolve_tree(h, tx, confidence, (int32_t)(k0 + k1));
  const bool classified = taxon != 0 && nd >= (uint32_t)min_hit_groups;
  r.taxon = classified ? tx.raw[taxon] : 0;
  r.flags = (classified ? SLK_F_CLASSIFIED : 0u) | (n_spans != 0 ? SLK_F_HAS_SPAN : 0u) | (h.overflow ? SLK_F_OVERFLOW : 0u);
  r.kmers1 = k0; r.kmers2 = k1; r.num_distinct = nd; r.n_hits = nh; r.n_probes = np;
}

// Owner of a compressed key among `world` hash-range shards: the range of the table-line mix x the key falls into.
// The build's cells, ordered by x, are thereby already grouped by owner, and the owner inserts what it receives front
// to back (slk_bucket_of with mix_mul = world keeps every shard's table uniformly loaded).
SLK_HD uint32_t slk_shard_of(uint64_t ckey, uint32_t world) {
  return slk_mulhi32(slk_key_mix(ckey), world);
}

// ------------------------------------------------------------------------------------------------ Bracken weights
// slacken/BrackenWeights.scala: every read of length readLen of every genome is "classified" against the library with
// a window that slides over the genome's taxon hits.
struct slk_bhit {      // TaxonHit (slacken/package.scala) of a genome fragment
  uint64_t key;        // compressed minimizer of a sequence super-mer (0 for the NONE quasi-hits)
  uint32_t ordinal;    // k-mer start position in the fragment (for the filler after a sequence segment: the reference's
                       // seq.length - (k-1), which lacks the segment's own position)
  uint32_t count;      // k-mer positions covered
  uint32_t flags;      // bit 0: distinct, bit 1: sequence super-mer (has a key to look up)
  uint32_t taxon;      // dense taxon, filled by the lookup step (0 = NONE)
};
#define SLK_BHIT_DISTINCT 1u
#define SLK_BHIT_SEQ 2u

// TaxonFragment.taxonHits (slacken/BrackenWeights.scala:199-236) without the taxa: calls emit(hit) for every hit of the
// fragment in order. Pieces as Supermers.splitByAmbiguity cuts them (slacken/Supermers.scala:150-177): a valid run with at
// least k characters gives its super-mers plus one NONE filler of k-1 positions, a shorter valid run and every gap give
// one NONE hit of their length.
template <int W, class Emit>
SLK_HD void slk_bracken_scan(const slk_scan_params& sp, const uint8_t* s, uint32_t len, Emit&& emit) {
  const uint32_t k = (uint32_t)sp.k;
  const slk_scan_masks sm(sp);
  const int fshift = sm.fshift, rshift = sm.rshift;
  const uint64_t mmask = sm.mmask, xor_mask = sm.xor_mask, sig_mask = sm.sig_mask;
  const bool canonical = sp.canonical != 0;
  slk_scanner<W> sc;
  sc.reset();
  uint64_t run_key = 0, last_key = 0;
  uint32_t run_cnt = 0, run_ord = 0, i = 0, piece_start = 0;
  bool in_run = false, first = true, piece_valid = false, have_piece = false;
  auto none_hit = [&](uint32_t ordinal, uint32_t count) {
    slk_bhit h; h.key = 0; h.ordinal = ordinal; h.count = count; h.flags = 0; h.taxon = 0;
    emit(h);
  };
  auto seq_hit = [&]() {
    slk_bhit h;
    h.key = slk_compress(sp, run_key << fshift); h.ordinal = run_ord; h.count = run_cnt; h.taxon = 0;
    h.flags = SLK_BHIT_SEQ | ((first || run_key != last_key) ? SLK_BHIT_DISTINCT : 0u);
    first = false; last_key = run_key;
    emit(h);
  };
  auto close_piece = [&](uint32_t end) {   // the piece [piece_start, end) is over
    const uint32_t plen = end - piece_start;
    if (piece_valid && plen >= k) {
      if (in_run) seq_hit();
      none_hit(plen - (k - 1), k - 1);
    } else {
      none_hit(piece_start, plen);
    }
    in_run = false;
  };
  slk_for_each_byte(s, len, [&](uint32_t ch) {
    const uint32_t c = slk_code(ch);
    const bool valid = c < 4u;
    if (!have_piece) { have_piece = true; piece_valid = valid; piece_start = i; }
    else if (valid != piece_valid) { close_piece(i); piece_valid = valid; piece_start = i; }
    uint64_t mn;
    const bool window_ok = sc.push(c, k, rshift, mmask, xor_mask, sig_mask, canonical, &mn);
    if (window_ok) {
      if (!(in_run && mn == run_key)) {
        if (in_run) seq_hit();
        run_key = mn; run_cnt = 1; run_ord = i - (k - 1); in_run = true;
      } else {
        run_cnt++;
      }
    }
    i++;
  });
  if (have_piece) close_piece(len);
}

// a small (taxon -> k-mers) map with removal: FragmentWindow.countSummary
struct slk_window_hist {
  uint32_t hk[SLK_KMAX];
  int32_t hv[SLK_KMAX];
  uint32_t n;
  bool overflow;
  SLK_HD uint32_t size() const { return n; }
  SLK_HD void at(uint32_t i, uint32_t* t, int32_t* v) const { *t = hk[i]; *v = hv[i]; }
  SLK_HD int32_t count(uint32_t t) const {
    for (uint32_t i = 0; i < n; i++)
      if (hk[i] == t) return hv[i];
    return 0;
  }
  SLK_HD void inc(uint32_t t) {
    for (uint32_t i = 0; i < n; i++)
      if (hk[i] == t) { hv[i]++; return; }
    if (n == SLK_KMAX) { overflow = true; return; }
    hk[n] = t; hv[n] = 1; n++;
  }
  SLK_HD void dec(uint32_t t) {   // put(updated) if updated > 0, else remove(key); an absent key stays absent
    for (uint32_t i = 0; i < n; i++)
      if (hk[i] == t) {
        if (--hv[i] <= 0) { n--; hk[i] = hk[n]; hv[i] = hv[n]; }
        return;
      }
  }
};

// FragmentWindow + TaxonFragment.readClassifications (slacken/BrackenWeights.scala:46-137,251-284): slides a window of
// read_len - (k-1) k-mer positions over the hits and calls dest(dense taxon) for every read start 0 .. len - read_len.
template <class Dest>
SLK_HD bool slk_bracken_window(const slk_tax_view& tx, const slk_bhit* hits, uint32_t n_hits, uint32_t len, uint32_t read_len,
                               uint32_t k, Dest&& dest) {
  if (len < read_len || n_hits == 0) return true;
  const uint32_t n_reads = len - read_len + 1;
  uint32_t w_start = 0, w_end = read_len - (k - 1);
  uint32_t head = 0, next = 0;
  while (next < n_hits && hits[next].ordinal < w_end) next++;   // hits.span(inWindow)
  slk_window_hist h;
  h.n = 0; h.overflow = false;
  uint32_t groups = 0;
  for (uint32_t i = head; i < next; i++) {
    const slk_bhit& x = hits[i];
    if ((x.flags & SLK_BHIT_DISTINCT) && x.taxon != 0) groups++;
    // k-mer positions of the hit that lie in [w_start, w_end)
    const uint64_t lo = x.ordinal > w_start ? x.ordinal : w_start;
    const uint64_t hi = (uint64_t)x.ordinal + x.count < w_end ? (uint64_t)x.ordinal + x.count : w_end;
    for (uint64_t q = lo; q < hi; q++) h.inc(x.taxon);
  }
  uint32_t last_in = next - 1;
  for (uint32_t start = 0; start < n_reads; start++) {
    if (start > 0) {   // advance()
      const slk_bhit rm = hits[head];
      h.dec(rm.taxon);
      w_start++; w_end++;
      if ((uint64_t)hits[head].ordinal + (hits[head].count - 1) < w_start) {
        head++;
        if ((rm.flags & SLK_BHIT_DISTINCT) && rm.taxon != 0) groups--;
      }
      if ((uint64_t)hits[last_in].ordinal + hits[last_in].count < w_end && next < n_hits) {
        last_in = next;
        if ((hits[next].flags & SLK_BHIT_DISTINCT) && hits[next].taxon != 0) groups++;
        next++;
      }
      h.inc(hits[last_in].taxon);
    }
    // TaxonFragment.classify: confidence 0, minHitGroups 2
    int32_t total = 0;
    dest(groups >= 2 ? slk_resolve_tree(h, tx, 0.0, total) : 0u);
  }
  return !h.overflow;
}

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
  const slk_scan_masks sm(sp);
  const int fshift = sm.fshift, rshift = sm.rshift;
  const uint64_t mmask = sm.mmask, xor_mask = sm.xor_mask, sig_mask = sm.sig_mask;
  const bool canonical = sp.canonical != 0;
  uint64_t run_key = 0;
  bool in_run = false;
  slk_for_each_byte(s, nbases, [&](uint32_t ch) {
    const uint32_t c = slk_code(ch);
    uint64_t mn;
    const bool window_ok = sc.push(c, k, rshift, mmask, xor_mask, sig_mask, canonical, &mn);
    if (window_ok && (!in_run || mn != run_key)) {
      out((slk_compress(sp, mn << fshift) << 16) | dense_taxon);
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
// mate = 0: the read itself; mate = 1: its mate of a read pair -- the same genome (or the same "random" decision), 250 bases
// further along where the genome allows it, the opposite strand, and independent errors.
SLK_HD uint8_t slk_synth_read_base(uint64_t gseed, uint64_t rseed, uint64_t n_genomes, uint64_t genome_len,
                                   uint64_t r, uint32_t L, uint32_t j, uint32_t mate = 0) {
  const uint64_t ms = mate ? 10u : 0u;   // mate 2 draws its N, its errors and its random bases from streams of its own
  uint64_t nn = slk_rnd(rseed, 14 + ms, r);
  if ((nn % 200) == 0 && (nn >> 8) % L == j) return 'N';
  uint64_t h = slk_rnd(rseed, 10, r);
  if ((h % 10) < 8) {
    uint64_t g = (h >> 8) % n_genomes;
    uint64_t pos = slk_rnd(rseed, 11, r) % (genome_len - L + 1);
    if (mate && pos + 250 <= genome_len - L) pos += 250;
    uint64_t base = g * genome_len + pos;
    uint8_t c = ((((h >> 40) & 1) != 0) != (mate != 0)) ? slk_comp_char(slk_synth_genome_base(gseed, base + (L - 1 - j)))
                                                        : slk_synth_genome_base(gseed, base + j);
    uint64_t e = slk_rnd(rseed, 12 + ms, r * 1024 + j);
    if ((e % 100) == 0 && c != 'N') c = slk_acgt((uint32_t)((e >> 8) & 3));
    return c;
  }
  uint64_t w = slk_rnd(rseed, 13 + ms, r * 32 + (j >> 5));
  return slk_acgt((uint32_t)((w >> (2 * (j & 31))) & 3));
}
