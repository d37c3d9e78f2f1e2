"""ctypes front-end of the CPU ORACLE (oracle/slk_oracle.c) plus the small text-level restatements
(per-read output line, kreport) that are easier to state in Python.

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py -- never from slacken_b200/.  Reference citations are relative to
/root/reference/src/main/scala/com/jnpersson/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from decimal import Decimal, ROUND_HALF_UP

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libslk_oracle.so")

DEFAULT_TOGGLE_MASK = 0xE37E28C4271B5A2D  # kmers/minimizer/package.scala:32
NONE, ROOT = 0, 1  # slacken/Taxonomy.scala:30-31
AMBIGUOUS_SPAN, MATE_PAIR_BORDER = -1, -2  # slacken/package.scala:28-29
SEQUENCE_FLAG, AMBIGUOUS_FLAG, MATE_PAIR_BORDER_FLAG = 1, 2, 3


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "slk_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("m", C.c_int32), ("spaces", C.c_int32), ("canonical", C.c_int32),
                ("ordering", C.c_int32), ("toggle_mask", C.c_uint64)]


def params(k=35, m=31, spaces=7, canonical=True, toggle_mask=DEFAULT_TOGGLE_MASK, ordering=0) -> Params:
    return Params(k, m, spaces, 1 if canonical else 0, ordering, toggle_mask)


class Result(C.Structure):
    _fields_ = [("taxon", C.c_int32), ("classified", C.c_uint8), ("has_span", C.c_uint8),
                ("num_distinct", C.c_int32), ("len1", C.c_int32), ("len2", C.c_int32), ("n_hits", C.c_int32)]


RESULT_DTYPE = np.dtype([("taxon", "<i4"), ("classified", "u1"), ("has_span", "u1"), ("_pad", "u2"),
                         ("num_distinct", "<i4"), ("len1", "<i4"), ("len2", "<i4"), ("n_hits", "<i4")])
HIT_DTYPE = np.dtype([("taxon", "<i4"), ("count", "<i4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        assert C.sizeof(Result) == RESULT_DTYPE.itemsize
        P = C.POINTER
        L.slko_priority.restype = C.c_uint64
        L.slko_priority.argtypes = [P(Params), C.c_uint64]
        L.slko_xor_mask.restype = C.c_uint64
        L.slko_xor_mask.argtypes = [P(Params)]
        L.slko_space_mask.restype = C.c_uint64
        L.slko_space_mask.argtypes = [P(Params)]
        L.slko_revcomp.restype = C.c_uint64
        L.slko_revcomp.argtypes = [C.c_uint64, C.c_int]
        L.slko_encode_window.restype = C.c_uint64
        L.slko_encode_window.argtypes = [C.c_char_p, C.c_int]
        L.slko_superkmers.restype = C.c_int64
        L.slko_superkmers.argtypes = [P(Params), C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.slko_all_matches.restype = C.c_int64
        L.slko_all_matches.argtypes = [P(Params), C.c_char_p, C.c_int64, C.c_void_p, C.c_void_p]
        L.slko_split_by_ambiguity.restype = C.c_int64
        L.slko_split_by_ambiguity.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.slko_spans.restype = C.c_int64
        L.slko_spans.argtypes = [P(Params), C.c_char_p, C.c_int64, C.c_char_p, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.slko_lca.restype = C.c_int32
        L.slko_lca.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.slko_has_ancestor.restype = C.c_int
        L.slko_has_ancestor.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.slko_resolve_tree.restype = C.c_int32
        L.slko_resolve_tree.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double]
        L.slko_lib_create.restype = C.c_void_p
        L.slko_lib_create.argtypes = [C.c_uint64]
        L.slko_lib_destroy.argtypes = [C.c_void_p]
        L.slko_lib_size.restype = C.c_uint64
        L.slko_lib_size.argtypes = [C.c_void_p]
        L.slko_lib_lookup.restype = C.c_int
        L.slko_lib_lookup.argtypes = [C.c_void_p, C.c_uint64, P(C.c_int32)]
        L.slko_lib_add_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.slko_lib_add_fragments.restype = C.c_int
        L.slko_lib_add_fragments.argtypes = [C.c_void_p, P(Params), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.slko_lib_add_sequences.restype = C.c_int
        L.slko_lib_add_sequences.argtypes = [C.c_void_p, P(Params), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.slko_lib_records.restype = C.c_uint64
        L.slko_lib_records.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.slko_classify_batch.restype = C.c_int
        L.slko_classify_batch.argtypes = [P(Params), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.slko_max_threads.restype = C.c_int
        L.slko_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.slko_synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
        L.slko_synth_mates.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
        L.slko_set_threads.argtypes = [C.c_int]
        L.slko_lib_clear_taxa.argtypes = [C.c_void_p]
        L.slko_lib_set_update_only.argtypes = [C.c_void_p, C.c_int]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _bytes(s) -> bytes:
    return s.encode("latin-1") if isinstance(s, str) else bytes(s)


# ----------------------------------------------------------------------------- scanner level
def priority(p: Params, window: int) -> int:
    return lib().slko_priority(C.byref(p), window)


def encode_window(s: str) -> int:
    b = _bytes(s)
    return lib().slko_encode_window(b, len(b))


def all_matches(p: Params, seq):
    b = _bytes(seq)
    pri = np.zeros(len(b) + 1, dtype=np.uint64)
    valid = np.zeros(len(b) + 1, dtype=np.uint8)
    n = lib().slko_all_matches(C.byref(p), b, len(b), _ptr(pri), _ptr(valid))
    if n < 0:
        raise ValueError("invalid nucleotide")
    return pri[:n], valid[:n].astype(bool)


def superkmers(p: Params, seq):
    """MinSplitter.superkmerPositions: list of (location, rank, length)."""
    b = _bytes(seq)
    cap = len(b) + 1
    loc = np.zeros(cap, dtype=np.int64)
    rank = np.zeros(cap, dtype=np.uint64)
    length = np.zeros(cap, dtype=np.int32)
    n = lib().slko_superkmers(C.byref(p), b, len(b), _ptr(loc), _ptr(rank), _ptr(length), cap)
    if n < 0:
        raise ValueError(f"superkmers failed ({n})")
    return [(int(loc[i]), int(rank[i]), int(length[i])) for i in range(n)]


def split_by_ambiguity(seq, k: int):
    b = _bytes(seq)
    cap = len(b) + 1
    st = np.zeros(cap, dtype=np.int64)
    ln = np.zeros(cap, dtype=np.int64)
    fl = np.zeros(cap, dtype=np.int32)
    n = lib().slko_split_by_ambiguity(b, len(b), k, _ptr(st), _ptr(ln), _ptr(fl), cap)
    return [(b[int(st[i]):int(st[i] + ln[i])].decode("latin-1"), int(fl[i]), int(st[i])) for i in range(n)]


def spans(p: Params, nt1, nt2=None):
    """Supermers.spans over Supermers.splitFragment: list of (minimizer, distinct, kmers, flag); ordinal = index."""
    b1 = _bytes(nt1)
    b2 = _bytes(nt2) if nt2 is not None else None
    cap = len(b1) + (len(b2) if b2 is not None else 0) + 4
    mn = np.zeros(cap, dtype=np.uint64)
    di = np.zeros(cap, dtype=np.uint8)
    km = np.zeros(cap, dtype=np.int32)
    fl = np.zeros(cap, dtype=np.uint8)
    n = lib().slko_spans(C.byref(p), b1, len(b1), b2, len(b2) if b2 is not None else 0,
                         _ptr(mn), _ptr(di), _ptr(km), _ptr(fl), cap)
    if n < 0:
        raise ValueError(f"spans failed ({n})")
    return [(int(mn[i]), bool(di[i]), int(km[i]), int(fl[i])) for i in range(n)]


# ----------------------------------------------------------------------------- taxonomy level
def lca(parents: np.ndarray, a: int, b: int) -> int:
    return lib().slko_lca(_ptr(parents), a, b)


def has_ancestor(parents: np.ndarray, t: int, anc: int) -> bool:
    return bool(lib().slko_has_ancestor(_ptr(parents), t, anc))


def resolve_tree(parents: np.ndarray, hits, confidence: float) -> int:
    """hits: iterable of (taxon, count) in read order (unmerged TaxonHits)."""
    taxa = np.array([h[0] for h in hits], dtype=np.int32)
    cnt = np.array([h[1] for h in hits], dtype=np.int32)
    return lib().slko_resolve_tree(_ptr(parents), _ptr(taxa), _ptr(cnt), len(taxa), float(confidence))


# ----------------------------------------------------------------------------- library + classify
def pack_sequences(seqs):
    """list of str/bytes -> (uint8 bases, int64 offsets[n+1])"""
    bs = [_bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs])
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, dtype=np.uint8)
    return bases, off


_VALID_BASES = None


def remove_invalid(seqs, taxa):
    """InputReader.removeInvalid (kmers/input/InputReader.scala:56-72): split every fragment around characters that
    are not ACGTU / newlines; each piece keeps the label of its fragment. Returns (pieces, labels)."""
    import re
    global _VALID_BASES
    if _VALID_BASES is None:
        _VALID_BASES = re.compile(rb"[ACTGUactgu][ACTGUactgu\n\r]*")
    out, lab = [], []
    for s, t in zip(seqs, taxa):
        for m in _VALID_BASES.finditer(_bytes(s)):
            out.append(m.group(0))
            lab.append(int(t))
    return out, np.array(lab, dtype=np.int32)


class Library:
    """The minimizer->LCA records of KeyValueIndex.makeRecords (slacken/KeyValueIndex.scala:85-122)."""

    def __init__(self, p: Params, parents: np.ndarray, expected_keys: int):
        self.p = p
        self.parents = np.ascontiguousarray(parents, dtype=np.int32)
        self.h = lib().slko_lib_create(int(expected_keys))
        if not self.h:
            raise MemoryError("oracle library allocation failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().slko_lib_destroy(self.h)
            self.h = None

    def add_fragments(self, bases: np.ndarray, off: np.ndarray, taxa: np.ndarray):
        off = np.ascontiguousarray(off, dtype=np.int64)
        taxa = np.ascontiguousarray(taxa, dtype=np.int32)
        rc = lib().slko_lib_add_fragments(self.h, C.byref(self.p), _ptr(self.parents), len(self.parents),
                                          _ptr(bases), _ptr(off), _ptr(taxa), len(taxa))
        if rc < 0:
            raise ValueError("invalid nucleotide in a genome fragment")

    def add_sequences(self, bases: np.ndarray, off: np.ndarray, taxa: np.ndarray):
        """Sequences that may still contain ambiguous characters: removeInvalid, then add_fragments."""
        off = np.ascontiguousarray(off, dtype=np.int64)
        taxa = np.ascontiguousarray(taxa, dtype=np.int32)
        rc = lib().slko_lib_add_sequences(self.h, C.byref(self.p), _ptr(self.parents), len(self.parents),
                                          _ptr(bases), _ptr(off), _ptr(taxa), len(taxa))
        if rc < 0:
            raise ValueError("add_sequences failed")

    def add_records(self, id1: np.ndarray, taxon: np.ndarray):
        id1 = np.ascontiguousarray(id1).view(np.uint64)
        taxon = np.ascontiguousarray(taxon, dtype=np.int32)
        lib().slko_lib_add_records(self.h, _ptr(self.parents), _ptr(id1), _ptr(taxon), len(id1))

    def clear_taxa(self):
        lib().slko_lib_clear_taxa(self.h)

    def set_update_only(self, on: bool):
        """Later add_* calls only update minimizers that are in the table already (see slko_lib_set_update_only)."""
        lib().slko_lib_set_update_only(self.h, 1 if on else 0)

    def __len__(self):
        return int(lib().slko_lib_size(self.h))

    def records(self):
        n = len(self)
        id1 = np.zeros(n, dtype=np.uint64)
        tx = np.zeros(n, dtype=np.int32)
        lib().slko_lib_records(self.h, _ptr(id1), _ptr(tx), n)
        o = np.argsort(id1, kind="stable")
        return id1[o], tx[o]

    def lookup(self, key: int):
        t = C.c_int32(0)
        return int(t.value) if lib().slko_lib_lookup(self.h, key, C.byref(t)) else None

    def classify(self, bases1, off1, bases2=None, off2=None, confidence=0.0, min_hit_groups=2, threads=0,
                 with_hits=True, lists=True):
        n = len(off1) - 1
        off1 = np.ascontiguousarray(off1, dtype=np.int64)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        if with_hits:
            l1 = np.diff(off1)
            ub = np.maximum(l1 - (self.p.k - 1), 0) + 1
            if bases2 is not None:
                off2 = np.ascontiguousarray(off2, dtype=np.int64)
                ub = ub + np.maximum(np.diff(off2) - (self.p.k - 1), 0) + 2
            hit_off = np.zeros(n + 1, dtype=np.int64)
            hit_off[1:] = np.cumsum(ub)
            hits = np.zeros(int(hit_off[-1]), dtype=HIT_DTYPE)
        else:
            hit_off = hits = None
            if bases2 is not None:
                off2 = np.ascontiguousarray(off2, dtype=np.int64)
        rc = lib().slko_classify_batch(C.byref(self.p), _ptr(self.parents), self.h, _ptr(bases1), _ptr(off1),
                                       _ptr(bases2), _ptr(off2), n, float(confidence), int(min_hit_groups),
                                       _ptr(res), _ptr(hit_off), _ptr(hits), int(threads))
        if rc < 0:
            raise ValueError(f"classify failed ({rc})")
        if with_hits:
            per_read = [hits[hit_off[i]:hit_off[i] + res["n_hits"][i]] for i in range(n)] if (lists and n <= 2_000_000) else None
            return res, hit_off, hits, per_read
        return res, None, None, None


def max_threads() -> int:
    return lib().slko_max_threads()


def host_threads() -> int:
    """The host cores this process may run on. torchrun exports OMP_NUM_THREADS=1, so the OpenMP default is not it."""
    import os
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def set_threads(n: int) -> int:
    """Fixes the OpenMP team size of every later oracle call (ignores OMP_NUM_THREADS)."""
    lib().slko_set_threads(int(n))
    return max_threads()


# ----------------------------------------------------------------------------- text level
def pairs_in_order_string(hits) -> str:
    """TaxonCounts.pairsInOrderString (slacken/TaxonCounts.scala:94-110) over MERGED hits."""
    out = []
    for t, c in hits:
        t, c = int(t), int(c)
        if t == MATE_PAIR_BORDER:
            out.append("|:|")
        elif t == AMBIGUOUS_SPAN:
            out.append(f"A:{c}")
        else:
            out.append(f"{t}:{c}")
    return " ".join(out)


def output_line(title: str, res, hits) -> str:
    """ClassifiedRead.outputLine (slacken/Classifier.scala:39-45) with lengthString (TaxonCounts.scala:114-121)."""
    flag = "C" if res["classified"] else "U"
    length = str(int(res["len1"])) if res["len2"] < 0 else f"{int(res['len1'])}|{int(res['len2'])}"
    return f"{flag}\t{title}\t{int(res['taxon'])}\t{length}\t{pairs_in_order_string(hits)}"


def java_format_fixed(x: float, decimals: int, width: int = 0) -> str:
    """java.util.Formatter %<width>.<decimals>f: HALF_UP on the shortest repr digits (Double.toString)."""
    q = Decimal(1).scaleb(-decimals)
    s = str(Decimal(repr(float(x))).quantize(q, rounding=ROUND_HALF_UP))
    return s.rjust(width)


RANK_CODES = {"unclassified": "U", "root": "R", "superkingdom": "D", "kingdom": "K", "phylum": "P", "class": "C",
              "order": "O", "family": "F", "genus": "G", "species": "S"}  # slacken/Taxonomy.scala:38-47


def kraken_report(parents: np.ndarray, ranks, names, counts) -> str:
    """KrakenReport.print (slacken/KrakenReport.scala:27-116). ranks[t]: rank title or None; names[t]: str or None;
    counts: iterable of (taxon, count)."""
    n = len(parents)
    counts = [(int(t), int(c)) for t, c in counts]
    taxon_counts: dict[int, int] = {}
    for t, c in counts:  # MMap.empty ++ counts : later duplicates overwrite
        taxon_counts[t] = c
    clade: dict[int, int] = {}
    for t, c in counts:
        x = t
        while x != NONE:
            clade[x] = clade.get(x, 0) + c
            x = int(parents[x])
        if t == NONE:
            clade[t] = c
    total = sum(c for _, c in counts)
    # Taxonomy.children (slacken/Taxonomy.scala:193-201): built by prepending in ascending taxid order
    children: dict[int, list[int]] = {}
    for taxid in range(n):
        if parents[taxid] != NONE or taxid == ROOT:
            children.setdefault(int(parents[taxid]), []).insert(0, taxid)
    lines = ["#Perc\tAggregate\tIn taxon\tRank\tTaxon\tName"]

    def line(taxid, code, rank_depth, depth):
        cc, tc = clade.get(taxid, 0), taxon_counts.get(taxid, 0)
        perc = java_format_fixed(100.0 * cc / total, 2, 6) if total else "   NaN"
        ds = "" if rank_depth == 0 else str(rank_depth)
        nm = names[taxid] if names is not None and names[taxid] is not None else ""
        return f"{perc}\t{cc}\t{tc}\t{code}{ds}\t{taxid}\t{'  ' * depth}{nm}"

    if taxon_counts.get(NONE, 0) != 0:
        lines.append(line(NONE, "U", 0, 0))
    stack = [(ROOT, "R", 0, 0)]
    while stack:
        taxid, code, rank_depth, depth = stack.pop()
        r = ranks[taxid] if ranks is not None else None
        if taxid == ROOT:
            r = "root"
        if r is not None and r in RANK_CODES:
            code_next, rd_next = RANK_CODES[r], 0
        else:
            code_next, rd_next = code, rank_depth + 1
        lines.append(line(taxid, code_next, rd_next, depth))
        ch = sorted(children.get(taxid, []), key=lambda c: -clade.get(c, 0))  # stable, descending clade count
        for c in reversed([c for c in ch if clade.get(c, 0) > 0]):
            stack.append((c, code_next, rd_next, depth + 1))
    return "\n".join(lines) + "\n"


def synth_genome(seed: int, start: int, n: int) -> np.ndarray:
    out = np.zeros(n, dtype=np.uint8)
    lib().slko_synth_genome(seed, start, n, _ptr(out))
    return out


def synth_reads(gseed: int, rseed: int, n_genomes: int, genome_len: int, first: int, n: int, L: int, mate: int = 0) -> np.ndarray:
    out = np.zeros(n * L, dtype=np.uint8)
    if mate:
        lib().slko_synth_mates(gseed, rseed, n_genomes, genome_len, first, n, L, mate, _ptr(out))
    else:
        lib().slko_synth_reads(gseed, rseed, n_genomes, genome_len, first, n, L, _ptr(out))
    return out


# ----------------------------------------------------------------------------- Bracken weights
# Restatement of slacken/BrackenWeights.scala:46-284,312-354. TEST INFRASTRUCTURE like the rest of this file; parity
# unpinned (the reference's known-answer test needs testData/slacken/slacken_tinydata.fna, which is not in the repo).
def split_to_max_length(seq: bytes, max_len: int, k: int):
    """TaxonFragment.splitToMaxLength (slacken/BrackenWeights.scala:152-164): (start, subsequence); consecutive pieces
    overlap by k-1 letters (k = the READ length here, as BrackenWeights.buildWeights calls it)."""
    if len(seq) <= max_len:
        return [(0, seq)]
    return [(s, seq[s:min(len(seq), s + max_len)]) for s in range(0, len(seq) - k + 1, max_len - (k - 1))]


def bracken_taxon_hits(p: Params, lookup, seq: bytes):
    """TaxonFragment.taxonHits (slacken/BrackenWeights.scala:199-236): [distinct, ordinal, taxon, count] per super-mer,
    plus the NONE quasi-hits that keep the k-mer positions of the fragment complete. `lookup(minimizer)` -> taxon or
    None. The filler after a sequence segment carries ordinal = seq.length - (k-1) WITHOUT the segment's position,
    exactly as the reference writes it."""
    k = p.k
    hits = []
    first, last = True, None
    for piece, flag, pos in split_by_ambiguity(seq, k):
        if flag == 1:   # SEQUENCE_FLAG
            for loc, rank, length in superkmers(p, piece):
                distinct = first or rank != last
                first, last = False, rank
                hits.append([distinct, loc + pos, lookup(rank) or 0, length - (k - 1)])
            hits.append([False, len(piece) - (k - 1), 0, k - 1])
        else:
            hits.append([False, pos, 0, len(piece)])
    return hits


def bracken_read_classifications(p: Params, parents: np.ndarray, lookup, seq: bytes, read_len: int):
    """TaxonFragment.readClassifications + FragmentWindow (slacken/BrackenWeights.scala:46-137,251-284): the destination
    taxon of every read of length read_len of the fragment, in order of its start position."""
    k = p.k
    K = read_len - (k - 1)
    hits = bracken_taxon_hits(p, lookup, seq)
    n_reads = len(seq) - read_len + 1
    if n_reads <= 0 or not hits:
        return []
    w_start, w_end = 0, K
    nxt = 0
    while nxt < len(hits) and hits[nxt][1] < w_end:     # hits.span(inWindow)
        nxt += 1
    window = list(range(nxt))                            # indices of the hits in the window
    groups = sum(1 for i in window if hits[i][0] and hits[i][2] != 0)
    last_in = window[-1]
    summary = {}
    for i in window:
        _, o, t, c = hits[i]
        for pos in range(o, o + c):
            if w_start <= pos < w_end:
                summary[t] = summary.get(t, 0) + 1
    out = []
    for start in range(n_reads):
        if start > 0:   # advance()
            rm = hits[window[0]]
            u = summary.get(rm[2], 0) - 1
            if u > 0:
                summary[rm[2]] = u
            else:
                summary.pop(rm[2], None)
            w_start += 1
            w_end += 1
            h0 = hits[window[0]]
            if h0[1] + (h0[3] - 1) < w_start:
                window.pop(0)
                if rm[0] and rm[2] != 0:
                    groups -= 1
            li = hits[last_in]
            if li[1] + li[3] < w_end and nxt < len(hits):
                window.append(nxt)
                last_in = nxt
                if hits[nxt][0] and hits[nxt][2] != 0:
                    groups += 1
                nxt += 1
            t = hits[last_in][2]
            summary[t] = summary.get(t, 0) + 1
        dest = resolve_tree(parents, list(summary.items()), 0.0) if groups >= 2 else 0
        out.append(dest)
    return out


def bracken_weights(p: Params, parents: np.ndarray, lookup, genomes, read_len: int, fragment_max: int = 1 << 20):
    """BrackenWeights.buildWeights (slacken/BrackenWeights.scala:312-354) for genomes = [(taxon, sequence)]: every genome
    sequence is one TaxonFragment, cut by splitToMaxLength(fragment_max, read_len). Returns {(dest, source): reads}."""
    out = {}
    for taxon, seq in genomes:
        seq = _bytes(seq)
        for _, piece in split_to_max_length(seq, fragment_max, read_len):
            for dest in bracken_read_classifications(p, parents, lookup, piece, read_len):
                out[(dest, int(taxon))] = out.get((dest, int(taxon)), 0) + 1
    return out


def kmer_distrib_lines(weights) -> list:
    """writeKmerDistrib (slacken/BrackenWeights.scala:418-430): one line per destination taxon,
    `dest \\t source:reads:total_reads_of_source ...` (the order of lines and of triples is Spark's, i.e. unspecified:
    here sorted)."""
    total = {}
    for (dest, src), c in weights.items():
        total[src] = total.get(src, 0) + c
    by_dest = {}
    for (dest, src), c in sorted(weights.items()):
        by_dest.setdefault(dest, []).append(f"{src}:{c}:{total[src]}")
    return ["mapped_taxid\tgenome_taxids:kmers_mapped:total_genome_kmers"] + [f"{d}\t{' '.join(v)}" for d, v in sorted(by_dest.items())]
