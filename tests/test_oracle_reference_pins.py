"""Pins the CPU oracle against what the reference's own tests hold for this path:
 - the MinSplitterTest known answer (src/test/scala/com/jnpersson/kmers/minimizer/MinSplitterTest.scala:25-32),
 - the constants derived in SURVEY.md section 8(a5),
 - the scalacheck property suites, transcribed with hypothesis:
     BitRepresentationProps / NTBitArrayProps, ShiftScannerProps, MinSplitterProps, SupermersProps,
     LowestCommonAncestorProps (resolveTree vs the reference's own slow model), TaxonomyProps, KrakenReportProps.
A second, pure-Python restatement of the priority function and the window minimum cross-checks the C code."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests.util import make_taxonomy, revcomp

MASK64 = (1 << 64) - 1
dna = st.text(alphabet="ACGT", min_size=0, max_size=300)
dna_mixed = st.text(alphabet="ACGTacgtUu", min_size=1, max_size=200)


# ------------------------------------------------------------------ pure-python second restatement
def py_encode(s: str) -> int:
    code = {"A": 0, "C": 1, "G": 2, "T": 3, "U": 3}
    x = 0
    for c in s.upper():
        x = (x << 2) | code[c]
    return (x << (64 - 2 * len(s))) & MASK64


def py_priority(mmer: str, spaces: int, mask: int, canonical: bool) -> int:
    m = len(mmer)
    s = mmer.upper().replace("U", "T")
    rc = revcomp(s.encode()).decode()
    canon = min(s, rc) if canonical else s            # lexicographic = numeric on the 2-bit codes A<C<G<T
    x = py_encode(canon) ^ ((mask << (64 - 2 * m)) & MASK64)
    sm = ((1 << (2 * m)) - 1) << (64 - 2 * m)
    for i in range(spaces):                           # positions 2,4,.. counted from the right end are zeroed
        sm &= ~(3 << (64 - 2 * m + 2 * (2 * i + 1)))
    return x & sm & MASK64 if spaces else x & MASK64


def py_superkmers(seq: str, k: int, m: int, spaces: int, mask: int, canonical: bool):
    """group consecutive k-mer windows by equal minimum m-mer priority"""
    if len(seq) < k:
        return []
    pri = [py_priority(seq[i:i + m], spaces, mask, canonical) for i in range(len(seq) - m + 1)]
    mins = [min(pri[q:q + k - m + 1]) for q in range(len(seq) - k + 1)]
    out, start = [], 0
    for q in range(1, len(mins) + 1):
        if q == len(mins) or mins[q] != mins[start]:
            out.append((start, mins[start], q - start + k - 1))
            start = q
    return out


# ------------------------------------------------------------------ known answers
def test_minsplitter_known_answer():
    p = oracle.params(k=5, m=2, spaces=0, canonical=False, ordering=1)   # MinTable.ofLength(2): lexicographic
    seq = "AATTTACTTTAGTTAC"
    got = [seq[l:l + n] for l, _, n in oracle.superkmers(p, seq)]
    assert got == ["AATTT", "ATTTA", "TTTACTTT", "CTTTA", "TTTAGTTA", "GTTAC"]


def test_derived_constants():
    p = oracle.params()
    assert oracle.lib().slko_xor_mask(p) == 0x8DF8A3109C6D68B4     # DEFAULT_TOGGLE_MASK << 2
    assert oracle.lib().slko_space_mask(p) == 0xFFFFFFFFCCCCCCCC   # s = 7, m = 31
    assert oracle.DEFAULT_TOGGLE_MASK == (-2054159557099562451) & MASK64   # XORmask in <idx>.properties
    # SpacedSeed doc example (MinimizerPriorities.scala:277-280): TTCTGTGGG, s = 3 -> TTCAGAGAG
    q = oracle.params(k=9, m=9, spaces=3, canonical=False, toggle_mask=0)
    assert oracle.priority(q, oracle.encode_window("TTCTGTGGG")) == oracle.encode_window("TTCAGAGAG")


# ------------------------------------------------------------------ codec (BitRepresentationProps, NTBitArrayProps)
@given(st.text(alphabet="ACGT", min_size=1, max_size=32))
def test_revcomp_matches_string_revcomp_and_is_involution(s):
    x = oracle.encode_window(s)
    assert x == py_encode(s)
    r = oracle.lib().slko_revcomp(x, len(s))
    assert r == py_encode(revcomp(s.encode()).decode())
    assert oracle.lib().slko_revcomp(r, len(s)) == x


@given(dna_mixed)
def test_char_codes_case_insensitive(s):
    for c in s:
        assert oracle.lib().slko_char_to_twobit(ord(c)) == "ACGT".index(c.upper().replace("U", "T"))
    for c in "NRYKM-*.0 ":
        assert oracle.lib().slko_char_to_twobit(ord(c)) == 5
    assert oracle.lib().slko_char_to_twobit(ord("\n")) == 4


@given(st.text(alphabet="ACGT", min_size=31, max_size=31))
def test_canonical_priority_is_rc_invariant_and_minimal(s):
    p = oracle.params()
    a = oracle.priority(p, oracle.encode_window(s))
    b = oracle.priority(p, oracle.encode_window(revcomp(s.encode()).decode()))
    assert a == b == py_priority(s, 7, oracle.DEFAULT_TOGGLE_MASK, True)
    assert a & ~0xFFFFFFFFCCCCCCCC == 0   # 48 significant bits


# ------------------------------------------------------------------ scanner (ShiftScannerProps)
@settings(max_examples=60, deadline=None)
@given(st.text(alphabet="ACGTacgt", min_size=0, max_size=200), st.integers(1, 31), st.booleans())
def test_scanner_finds_every_mmer_at_its_position(s, m, canonical):
    p = oracle.params(k=m + 4, m=m, spaces=0, canonical=canonical)
    pri, valid = oracle.all_matches(p, s)
    assert len(pri) == len(s)
    for i in range(len(s)):
        if i < m - 1:
            assert not valid[i]
        else:
            assert valid[i]
            assert int(pri[i]) == py_priority(s[i - m + 1:i + 1], 0, oracle.DEFAULT_TOGGLE_MASK, canonical)


def test_scanner_skips_newlines():
    p = oracle.params()
    rng = np.random.default_rng(3)
    s = "".join(rng.choice(list("ACGT"), 400))
    broken = "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\r\n"
    assert oracle.superkmers(p, s) == oracle.superkmers(p, broken)


# ------------------------------------------------------------------ splitter (MinSplitterProps)
splitter_params = st.tuples(st.integers(1, 31), st.integers(0, 12), st.integers(0, 7), st.booleans()).map(
    lambda t: (t[0] + t[1], t[0], min(t[2], t[0] // 2), t[3]))


@settings(max_examples=80, deadline=None)
@given(dna, splitter_params)
def test_superkmers_match_window_model_and_tile_the_read(s, kmsc):
    k, m, spaces, canonical = kmsc
    p = oracle.params(k=k, m=m, spaces=spaces, canonical=canonical)
    got = oracle.superkmers(p, s)
    assert got == py_superkmers(s, k, m, spaces, oracle.DEFAULT_TOGGLE_MASK, canonical)
    if len(s) < k:
        assert got == []
        return
    # tiling with k-1 overlap, adjacent minimizers differ
    assert got[0][0] == 0 and got[-1][0] + got[-1][2] == len(s)
    for (l0, r0, n0), (l1, r1, n1) in zip(got, got[1:]):
        assert l1 == l0 + n0 - (k - 1)
        assert r0 != r1


@settings(max_examples=40, deadline=None)
@given(st.text(alphabet="ACGT", min_size=35, max_size=250))
def test_superkmer_ranks_reverse_under_reverse_complement(s):
    p = oracle.params()
    fwd = [(r, n) for _, r, n in oracle.superkmers(p, s)]
    rc = [(r, n) for _, r, n in oracle.superkmers(p, revcomp(s.encode()).decode())]
    assert fwd == rc[::-1]


# ------------------------------------------------------------------ ambiguity (SupermersProps)
@settings(max_examples=80, deadline=None)
@given(st.text(alphabet="ACGTN", min_size=0, max_size=200), st.integers(3, 40))
def test_split_by_ambiguity(s, k):
    pieces = oracle.split_by_ambiguity(s, k)
    assert "".join(t for t, _, _ in pieces) == s
    pos = 0
    for text, flag, at in pieces:
        assert at == pos
        pos += len(text)
        if flag == oracle.SEQUENCE_FLAG:
            assert "N" not in text and len(text) >= k
        else:
            assert "N" in text or len(text) < k


def test_spans_quirks():
    p = oracle.params()
    rng = np.random.default_rng(9)
    a = "".join(rng.choice(list("ACGT"), 80))
    # short pieces vanish, >= k ambiguous run gives one span of len-(k-1)
    sp = oracle.spans(p, a[:20] + "N" * 50 + a + "N" * 10 + a[:30])
    flags = [f for _, _, _, f in sp]
    assert flags.count(oracle.AMBIGUOUS_FLAG) == 1
    amb = [x for x in sp if x[3] == oracle.AMBIGUOUS_FLAG][0]
    assert amb[2] == 50 - 34 and sp[0][3] == oracle.AMBIGUOUS_FLAG
    assert sum(c for _, _, c, f in sp if f == oracle.SEQUENCE_FLAG) == 80 - 34
    # a read shorter than k yields nothing; a pair always has the border span with kmers = -(k-1)
    assert oracle.spans(p, a[:34]) == []
    sp = oracle.spans(p, "ACG", "TTT")
    assert [(c, f) for _, _, c, f in sp] == [(-34, oracle.MATE_PAIR_BORDER_FLAG)]
    # `distinct` survives the mate border: the first span of mate 2 repeats mate 1's last minimizer
    sp = oracle.spans(p, a[:40], a[:40])
    seq = [x for x in sp if x[3] == oracle.SEQUENCE_FLAG]
    n1 = len(oracle.spans(p, a[:40]))
    assert seq[0][1] is True and all(d for _, d, _, _ in seq[:n1])
    m1_last, m2_first = seq[n1 - 1], seq[n1]
    assert m2_first[1] == (m2_first[0] != m1_last[0])


# ------------------------------------------------------------------ taxonomy / LCA / resolveTree
def path_to_root(parents, t):
    out = []
    while t != 0:
        out.append(t)
        t = int(parents[t])
    return out


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 10_000))
def test_lca_properties(seed):
    parents, _, _ = make_taxonomy(100, seed)
    rng = np.random.default_rng(seed)
    taxa = [t for t in range(1, len(parents)) if parents[t] != 0 or t == 1]
    for _ in range(30):
        a, b = (int(x) for x in rng.choice(taxa, 2))
        l = oracle.lca(parents, a, b)
        pa, pb = path_to_root(parents, a), path_to_root(parents, b)
        common = [x for x in pa if x in pb]
        assert l == common[0] == oracle.lca(parents, b, a)
        assert oracle.lca(parents, a, 0) == a and oracle.lca(parents, 0, b) == b
        assert oracle.has_ancestor(parents, a, l) and oracle.has_ancestor(parents, b, l)
        assert oracle.has_ancestor(parents, a, a) and not oracle.has_ancestor(parents, 0, a)


def correct_classification(parents, hits, threshold):
    """LowestCommonAncestorProps.correctClassification (the reference's own easy-to-trust model)."""
    def has_anc(t, a):
        return a in path_to_root(parents, t)
    total = sum(c for _, c in hits)
    distinct = []
    for t, _ in hits:
        if t != 0 and t not in distinct:
            distinct.append(t)

    def frac_above(t):
        return sum(c for h, c in hits if has_anc(t, h)) / total if hits else 0

    def frac_below(t):
        return sum(c for h, c in hits if has_anc(h, t)) / total if hits else 0
    best = sorted(((frac_above(t), t) for t in distinct), key=lambda x: x[0], reverse=True)
    if not best:
        return 0
    bf, bt = best[0]
    for f, t in best[1:]:
        if f != bf:
            break
        bt = oracle.lca(parents, bt, t)
    for t in path_to_root(parents, bt):
        if frac_below(t) >= threshold:
            return t
    return 0


@settings(max_examples=150, deadline=None)
@given(st.integers(0, 1000), st.integers(10, 200), st.floats(0, 1), st.floats(0, 1), st.integers(0, 2**31))
def test_resolve_tree_against_reference_model(tseed, kmers, invalid_frac, threshold, rseed):
    parents, _, _ = make_taxonomy(100, tseed)
    rng = np.random.default_rng(rseed)
    taxa = [t for t in range(1, len(parents)) if parents[t] != 0 or t == 1]
    pool = [int(x) for x in rng.choice(taxa, int(rng.integers(1, 8)))]
    invalid = int(kmers * invalid_frac)
    hits, left = [], kmers - invalid
    while left > 0:
        c = min(left, int(rng.integers(1, 11)))
        hits.append((pool[int(rng.integers(len(pool)))], c))
        left -= c
    while invalid > 0:
        c = min(invalid, int(rng.integers(1, 11)))
        hits.append((0, c))
        invalid -= c
    rng.shuffle(hits)
    hits = [(int(t), int(c)) for t, c in hits]
    want = correct_classification(parents, hits, threshold)
    got = oracle.resolve_tree(parents, hits, threshold)
    # the model compares fractions, resolveTree compares integer scores with ceil(): they can only disagree when
    # threshold * total is within float rounding of an integer score
    total = sum(c for _, c in hits)
    if abs(threshold * total - round(threshold * total)) > 1e-9:
        assert got == want


# ------------------------------------------------------------------ report (KrakenReportProps)
@settings(max_examples=40, deadline=None)
@given(st.integers(0, 1000), st.integers(0, 2**31))
def test_report_aggregation(tseed, rseed):
    parents, ranks, names = make_taxonomy(60, tseed)
    rng = np.random.default_rng(rseed)
    taxa = [t for t in range(0, len(parents)) if t < 2 or parents[t] != 0]
    picked = sorted(set(int(x) for x in rng.choice(taxa, int(rng.integers(1, 30)))))
    counts = [(t, int(rng.integers(1, 100))) for t in picked]
    text = oracle.kraken_report(parents, ranks, names, counts)
    rows = [l.split("\t") for l in text.splitlines()[1:]]
    own = {int(r[4]): int(r[2]) for r in rows}
    clade = {int(r[4]): int(r[1]) for r in rows}
    for t, c in counts:
        assert own[t] == c and clade[t] >= c
    if 1 in clade:
        assert clade[1] == sum(c for t, c in counts if t != 0)
    total = sum(c for _, c in counts)
    for r in rows:
        assert r[0] == oracle.java_format_fixed(100.0 * int(r[1]) / total, 2, 6)


def test_java_half_up_rounding():
    assert oracle.java_format_fixed(0.125, 2) == "0.13"      # C printf would give 0.12
    assert oracle.java_format_fixed(1.005, 2) == "1.01"      # shortest repr "1.005" rounds half up
    assert oracle.java_format_fixed(2.5, 0) == "3"
    assert oracle.java_format_fixed(100.0, 2, 6) == "100.00"
    assert oracle.java_format_fixed(0.0, 2, 6) == "  0.00"
