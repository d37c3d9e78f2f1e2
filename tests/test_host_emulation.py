"""The CUDA kernel bodies (slk_core.h), run thread by thread on the CPU by tests/host_emulation, against the oracle.
This is the no-GPU half of the parity suite; tests/test_gpu_parity.py runs the same comparisons on the device."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from tests import host_emulation as emu
from tests.util import chimeric_reads, leaf_taxa, make_taxonomy, random_dna, simulate_reads


def world(seed, k=35, m=31, s=7, canonical=True, n_genomes=16, glen=2500):
    rng = np.random.default_rng(seed)
    parents, ranks, names = make_taxonomy(100, seed)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, glen) for _ in range(n_genomes)]
    for i in range(n_genomes // 2, n_genomes):
        g = bytearray(genomes[i - n_genomes // 2])
        for j in range(0, len(g), 53):
            g[j] = ord("ACGT"[int(rng.integers(4))])
        genomes[i] = bytes(g)
    taxa = np.array([leaves[int(rng.integers(len(leaves)))] for _ in genomes], dtype=np.int32)
    p = oracle.params(k=k, m=m, spaces=s, canonical=canonical)
    lib = oracle.Library(p, parents, 1 << 17)
    b, off = oracle.pack_sequences(genomes)
    lib.add_fragments(b, off, taxa)
    sp = emu.scan_params(k, m, s, oracle.DEFAULT_TOGGLE_MASK, canonical)
    return rng, parents, genomes, taxa, p, lib, sp


def compare(lib, ix, rb, ro, rb2=None, ro2=None, confidence=0.0, k=35):
    for packed in (False, True):   # both input forms of the ABI must give the same answer
        _compare(lib, ix, rb, ro, rb2, ro2, confidence, k, packed, False)
    _compare(lib, ix, rb, ro, rb2, ro2, confidence, k, False, True)   # ... and so must the split (sharded-library) path
    _compare2(lib, ix, rb, ro, rb2, ro2, confidence, k)                # ... and the warp-cooperative kernel body (slk_group.h)


def _compare2(lib, ix, rb, ro, rb2, ro2, confidence, k):
    from slacken_b200.host import pack_reads
    res, _, _, per = lib.classify(rb, ro, rb2, ro2, confidence=confidence)
    r1 = pack_reads(np.frombuffer(bytes(rb), dtype=np.uint8), np.asarray(ro))
    r2 = pack_reads(np.frombuffer(bytes(rb2), dtype=np.uint8), np.asarray(ro2)) if rb2 is not None else None
    taxon, flags, detail, hits, probes, merged, counts = emu.classify2(ix, r1, r2, confidence=confidence)
    assert np.array_equal(res["taxon"], taxon)
    assert np.array_equal(res["classified"], flags & 1)
    assert np.array_equal(res["has_span"], (flags >> 1) & 1)
    hs = res["has_span"].astype(bool)
    assert np.array_equal(res["num_distinct"][hs], detail["num_distinct"][hs].astype(np.int32))
    assert np.array_equal(res["len1"][hs], detail["len1"][hs].astype(np.int32))
    if rb2 is not None:
        assert np.array_equal(res["len2"][hs], detail["len2"][hs].astype(np.int32))
    for i in range(len(per)):
        h = hits[int(detail["hit_off"][i]):int(detail["hit_off"][i]) + int(detail["hit_cnt"][i])]
        assert np.array_equal(h["taxon"], per[i]["taxon"]) and np.array_equal(h["count"], per[i]["count"]), i
    assert merged == sum(len(x) for x in per)
    rep = np.bincount(res["taxon"][hs], minlength=len(counts))
    assert np.array_equal(rep, counts.astype(np.int64)[:len(rep)])
    # without per-read hit lists (the --nodetailed mode): the same taxa
    t2, f2, _, _, _, _, _ = emu.classify2(ix, r1, r2, confidence=confidence, want_hits=False)
    assert np.array_equal(t2, taxon) and np.array_equal(f2, flags)
    # several thresholds in one pass (Classifier.scala:156-170): row t equals a separate call with threshold t
    thr = [confidence, 0.05, 0.3, 0.9]
    tm, fm, dm, hm, _, _, _ = emu.classify2(ix, r1, r2, confidence=thr)
    assert np.array_equal(tm[0], taxon) and np.array_equal(fm[0], flags) and np.array_equal(hm, hits)
    for t in range(1, len(thr)):
        rt, _, _, _ = lib.classify(rb, ro, rb2, ro2, confidence=thr[t])
        assert np.array_equal(tm[t], rt["taxon"]), thr[t]
        assert np.array_equal(fm[t] & 1, rt["classified"]) and np.array_equal((fm[t] >> 1) & 1, rt["has_span"])


def _compare(lib, ix, rb, ro, rb2, ro2, confidence, k, packed, split):
    res, _, _, per = lib.classify(rb, ro, rb2, ro2, confidence=confidence)
    eres, ehoff, ehits = ix.classify(rb, ro, rb2, ro2, confidence=confidence, packed=packed, split=split)
    assert np.array_equal(res["taxon"], eres["taxon"])
    assert np.array_equal(res["classified"], eres["flags"] & 1)
    assert np.array_equal(res["has_span"], (eres["flags"] >> 1) & 1)
    hs = res["has_span"].astype(bool)
    assert np.array_equal(res["num_distinct"][hs], eres["num_distinct"][hs].astype(np.int32))
    assert np.array_equal(res["len1"][hs], eres["kmers1"][hs].astype(np.int32) + (k - 1))
    if rb2 is not None:
        assert np.array_equal(res["len2"][hs], eres["kmers2"][hs].astype(np.int32) + (k - 1))
    for i in range(len(per)):
        h = ehits[int(ehoff[i]):int(ehoff[i + 1])]
        assert np.array_equal(h["taxon"], per[i]["taxon"]) and np.array_equal(h["count"], per[i]["count"]), i


@pytest.mark.parametrize("confidence", [0.0, 0.15, 0.7])
def test_classify_body_single_end(confidence):
    rng, parents, genomes, taxa, p, lib, sp = world(5)
    id1, tx = lib.records()
    ix = emu.EmuIndex(sp, parents, id1, tx)
    reads = simulate_reads(rng, genomes, 1200, (10, 230), n_rate=0.2) + [b"", b"ACGT", b"N" * 90, b"A" * 120]
    rb, ro = oracle.pack_sequences(reads)
    compare(lib, ix, rb, ro, confidence=confidence)


def test_classify_body_long_reads_with_many_hits():
    rng, parents, genomes, taxa, p, lib, sp = world(8)
    id1, tx = lib.records()
    ix = emu.EmuIndex(sp, parents, id1, tx)
    reads = chimeric_reads(rng, genomes, 12, 150) + chimeric_reads(rng, genomes, 20, 30) + simulate_reads(rng, genomes, 50, 100)
    rb, ro = oracle.pack_sequences(reads)
    res, _, _, per = lib.classify(rb, ro, confidence=0.02)
    assert max(len(h) for h in per) > 70          # beyond SLK_SHITS + SLK_XHITS: the spill path runs
    compare(lib, ix, rb, ro, confidence=0.02)
    compare(lib, ix, rb, ro, rb, ro, confidence=0.3)


def test_classify_body_paired_end():
    rng, parents, genomes, taxa, p, lib, sp = world(6)
    id1, tx = lib.records()
    ix = emu.EmuIndex(sp, parents, id1, tx)
    r1 = simulate_reads(rng, genomes, 600, (20, 160), n_rate=0.15)
    r2 = simulate_reads(rng, genomes, 600, (20, 160), n_rate=0.15)
    r1[0], r2[0] = b"AC", b"GT"
    b1, o1 = oracle.pack_sequences(r1)
    b2, o2 = oracle.pack_sequences(r2)
    for c in (0.0, 0.3):
        compare(lib, ix, b1, o1, b2, o2, confidence=c)


@pytest.mark.parametrize("k,m,s,canonical", [(31, 24, 0, True), (25, 20, 3, False), (28, 28, 4, True), (22, 15, 0, True),
                                             (12, 5, 1, True)])
def test_classify_and_emit_other_parameters(k, m, s, canonical):
    rng, parents, genomes, taxa, p, lib, sp = world(10 + k, k, m, s, canonical, n_genomes=8, glen=1500)
    id1, tx = lib.records()
    ix = emu.EmuIndex(sp, parents, id1, tx)
    reads = simulate_reads(rng, genomes, 400, (k - 2, 140), n_rate=0.1)
    rb, ro = oracle.pack_sequences(reads)
    compare(lib, ix, rb, ro, confidence=0.1, k=k)
    # build side: the cells the emit threads produce = the oracle's super-mer minimizers (as a set)
    for g in genomes[:3]:
        cells = emu.emit_cells(sp, g, 7)
        assert set(int(c) & 0xFFFF for c in cells) == {7}
        got = {emu.expand(sp, int(c) >> 16) for c in cells}
        want = {r for _, r, _ in oracle.superkmers(p, g)}
        assert got == want


def test_emit_cells_break_at_ambiguous_characters():
    rng, parents, genomes, taxa, p, lib, sp = world(3)
    g = bytearray(genomes[0])
    g[500:520] = b"N" * 20
    g[1200] = ord("x")
    pieces, _ = oracle.remove_invalid([bytes(g)], [1])
    want = set()
    for piece in pieces:
        want |= {r for _, r, _ in oracle.superkmers(p, piece)}
    for wpt in (96, 7, 1000):
        cells = emu.emit_cells(sp, bytes(g), 3, wpt)
        assert {emu.expand(sp, int(c) >> 16) for c in cells} == want


@settings(max_examples=200, deadline=None)
@given(st.integers(0, 2**64 - 1), st.sampled_from([(31, 7), (24, 0), (20, 3), (28, 4), (15, 0), (31, 15)]))
def test_key_compression_round_trip(x, ms):
    m, s = ms
    sp = emu.scan_params(m + 4, m, s, oracle.DEFAULT_TOGGLE_MASK, True)
    key = x & sp.sig_mask
    c = emu.compress(sp, key)
    assert c < (1 << sp.key_bits)
    assert emu.expand(sp, c) == key
    # order preserving
    y = (x * 0x9E3779B97F4A7C15 + 12345) & ((1 << 64) - 1) & sp.sig_mask
    assert (key < y) == (c < emu.compress(sp, y)) or key == y


def test_unsupported_parameters_are_rejected():
    for k, m, s in [(35, 31, 0), (40, 31, 7), (35, 32, 8), (30, 31, 7)]:
        with pytest.raises(ValueError):
            emu.scan_params(k, m, s, oracle.DEFAULT_TOGGLE_MASK, True)


def test_synthetic_generators_agree_with_the_oracle_copy():
    n = 200_000
    a = np.zeros(n, dtype=np.uint8)
    emu.lib().emu_synth_genome(7, 60_000, n, emu._p(a))
    assert np.array_equal(a, oracle.synth_genome(7, 60_000, n))
    assert (a == ord("N")).sum() > 0
    r = np.zeros(3000 * 150, dtype=np.uint8)
    emu.lib().emu_synth_reads(7, 8, 4, 70_000, 11, 3000, 150, emu._p(r))
    assert np.array_equal(r, oracle.synth_reads(7, 8, 4, 70_000, 11, 3000, 150))
    # mate 2 of a read pair: same genome 250 bases on, opposite strand (the paired leg of bench.py)
    r2 = np.zeros(3000 * 150, dtype=np.uint8)
    emu.lib().emu_synth_mates(7, 8, 4, 70_000, 11, 3000, 150, 1, emu._p(r2))
    assert np.array_equal(r2, oracle.synth_reads(7, 8, 4, 70_000, 11, 3000, 150, mate=1))
    assert not np.array_equal(r, r2)
    emu.lib().emu_synth_mates(7, 8, 4, 70_000, 11, 3000, 150, 0, emu._p(r2))
    assert np.array_equal(r, r2)


@pytest.mark.parametrize("read_len", [50, 100])
def test_bracken_window_body_matches_the_oracle(read_len):
    """hits of a genome fragment (incl. the NONE quasi-hits around ambiguous stretches) and the
    sliding window of FragmentWindow, read by read, against the oracle's restatement of slacken/BrackenWeights.scala."""
    rng, parents, genomes, taxa, p, lib, sp = world(11, n_genomes=6, glen=1800)
    id1, tx = lib.records()
    ix = emu.EmuIndex(sp, parents, id1, tx)
    cases = list(genomes[:3])
    g = bytearray(genomes[3]); g[300:350] = b"N" * 50; g[700] = ord("N"); g[720:740] = b"n" * 20; g[1790:] = b"N" * 10
    cases.append(bytes(g))
    cases.append(b"N" * 30 + genomes[4][:400] + b"NN" + genomes[4][400:430] + b"N" + genomes[4][430:900])
    cases += [genomes[5][:read_len], genomes[5][:read_len - 1], b"N" * 200, random_dna(rng, 600)]
    for seq in cases:
        want = oracle.bracken_read_classifications(p, parents, lib.lookup, seq, read_len)
        got = emu.bracken_dests(ix, seq, read_len)
        assert list(got) == want, (len(seq), read_len)


def test_four_bytes_at_a_time_coding_equals_the_scalar_coding():
    """slk_code4 (the stage-1 pack kernel's byte-parallel coding) against four slk_code calls: every byte value in every
    lane next to every accepted letter, plus random words."""
    L = emu.lib()
    letters = np.frombuffer(b"ACGTUacgtuNn\x00\xff@[`{SsVv", dtype=np.uint8).astype(np.uint32)
    words = []
    for lane in range(4):
        for other in letters:
            w = np.full(256, 0, dtype=np.uint32)
            for j in range(4):
                w |= (np.arange(256, dtype=np.uint32) if j == lane else other) << np.uint32(8 * j)
            words.append(w)
    words.append(np.random.default_rng(3).integers(0, 1 << 32, size=200000, dtype=np.uint64).astype(np.uint32))
    # words made of accepted letters only (the common case)
    acgt = np.frombuffer(b"ACGTacgtUu", dtype=np.uint8)
    pick = np.random.default_rng(4).integers(0, len(acgt), size=(50000, 4))
    words.append((acgt[pick].astype(np.uint32) << (np.arange(4, dtype=np.uint32) * 8)).sum(axis=1).astype(np.uint32))
    allw = np.ascontiguousarray(np.concatenate(words))
    assert L.emu_code4_mismatches(allw.ctypes.data, len(allw)) == 0


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_range_partition_is_one_virtual_table_cut_in_equal_ranges(world):
    """slk_shard_of / slk_bucket_of with mix_mul = world: the owner of a key and its line inside the
    owner's shard, read as one number (owner * lines + line), must be monotone in the table-line hash of the key -- which is
    what lets the distributed build ship cells ordered by that hash to their owners and insert them front to back -- and
    every shard must be used evenly over all of its lines."""
    L = emu.lib()
    rng = np.random.default_rng(world)
    keys = rng.integers(0, 1 << 48, size=20000, dtype=np.uint64)
    n_buckets = 4 * 1000                                  # 1000 lines per shard, not a power of two
    x = np.array([L.emu_key_mix(int(k)) for k in keys], dtype=np.int64)
    owner = np.array([L.emu_shard_of(int(k), world) for k in keys], dtype=np.int64)
    bucket = np.array([L.emu_bucket_of(int(k), n_buckets, world) for k in keys], dtype=np.int64)
    line = bucket >> 2
    assert owner.min() >= 0 and owner.max() < world and line.min() >= 0 and line.max() < 1000
    order = np.argsort(x, kind="stable")
    virtual = (owner * 1000 + line)[order]
    assert (np.diff(virtual) >= 0).all()
    assert np.array_equal(owner, (x * world) >> 32)
    # even use: every shard gets about 1/world of the keys, and every tenth of a shard's lines about a tenth of those
    for r in range(world):
        mine = line[owner == r]
        assert abs(len(mine) - len(keys) / world) < 6 * np.sqrt(len(keys) / world)
        h = np.bincount(mine // 100, minlength=10)
        assert h.min() > 0.8 * len(mine) / 10 and h.max() < 1.2 * len(mine) / 10
    # the sub-bucket inside the line does not depend on the cut
    assert np.array_equal(bucket & 3, np.array([L.emu_bucket_of(int(k), n_buckets, 1) for k in keys], dtype=np.int64) & 3)
