"""Parity of the CUDA path (through the C ABI) with the CPU oracle: bit-exact taxon per read, merged hit lists,
hit-group counts, library records and report counts."""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import Classifier, IndexParams, KeyValueIndex, KrakenReport, ReportCounts, Taxonomy
from slacken_b200.host import block_offsets, pack_reads, pack_sequences
from slacken_b200.report import output_line
from tests.util import chimeric_reads, leaf_taxa, make_taxonomy, random_dna, simulate_reads

pytestmark = pytest.mark.gpu


def make_world(seed, n_genomes=24, glen=5000, tax_size=120, related=True):
    rng = np.random.default_rng(seed)
    parents, ranks, names = make_taxonomy(tax_size, seed + 1)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, glen) for _ in range(n_genomes)]
    if related:  # strains: mutated copies share most minimizers, so LCAs land on inner nodes
        for i in range(n_genomes // 2, n_genomes):
            g = bytearray(genomes[i - n_genomes // 2])
            for j in range(0, len(g), 61):
                g[j] = ord("ACGT"[int(rng.integers(4))])
            genomes[i] = bytes(g)
    # ambiguous stretches inside genomes
    g0 = bytearray(genomes[0]); g0[100:140] = b"N" * 40; g0[900] = ord("n"); genomes[0] = bytes(g0)
    taxa = np.array([leaves[int(rng.integers(len(leaves)))] for _ in genomes], dtype=np.int32)
    return rng, parents, ranks, names, genomes, taxa


def oracle_lib(p, parents, genomes, taxa):
    lib = oracle.Library(p, parents, 1 << 18)
    pieces, labels = oracle.remove_invalid(genomes, taxa)   # the host-side split the reference applies first
    b, off = oracle.pack_sequences(pieces)
    lib.add_fragments(b, off, labels)
    return lib


def assert_batch_equal(res, per, got, k):
    assert np.array_equal(res["taxon"], got.taxon)
    assert np.array_equal(res["classified"].astype(bool), got.classified)
    assert np.array_equal(res["has_span"].astype(bool), got.has_span)
    hs = res["has_span"].astype(bool)
    assert np.array_equal(res["num_distinct"][hs], got.detail["num_distinct"][hs].astype(np.int32))
    assert np.array_equal(res["len1"][hs], got.detail["len1"][hs].astype(np.int32))
    l2 = got.detail["len2"].astype(np.int64)
    l2[l2 == 0xFFFFFFFF] = -1
    assert np.array_equal(res["len2"][hs], l2[hs])
    for i in range(len(per)):
        h = got.hits_of(i)
        assert np.array_equal(h["taxon"], per[i]["taxon"]), i
        assert np.array_equal(h["count"], per[i]["count"]), i


@pytest.mark.parametrize("confidence", [0.0, 0.15, 0.6])
def test_classify_single_end(gpu, confidence):
    rng, parents, ranks, names, genomes, taxa = make_world(21)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    assert len(index) == len(id1)
    reads = simulate_reads(rng, genomes, 3000, (10, 260), n_rate=0.15) + [b"", b"ACGT", b"N" * 80, b"A" * 200]
    rb, ro = pack_sequences(reads)
    cls = Classifier(index)
    got = cls.classify(rb, ro, confidence=confidence)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=confidence)
    assert_batch_equal(res, per, got, 35)
    # report-only mode gives the same taxa
    got2 = cls.classify(rb, ro, confidence=confidence, per_read_output=False)
    assert np.array_equal(got2.taxon, got.taxon) and np.array_equal(got2.flags, got.flags)
    # output lines, via the two independent text emitters
    for i in range(0, len(reads), 97):
        if res["has_span"][i]:
            a = oracle.output_line(f"r{i}", res[i], [(int(h["taxon"]), int(h["count"])) for h in per[i]])
            b = output_line(f"r{i}", got.taxon[i], got.classified[i], got.detail[i], got.hits_of(i))
            assert a == b
    cls.close(); index.close(); tax.close()


def test_classify_long_reads_with_many_hits(gpu):
    """Reads of several kb whose merged hit lists outgrow the per-thread buffers (spill to a worst-case global block),
    mixed with ordinary reads in the same warps."""
    rng, parents, ranks, names, genomes, taxa = make_world(44)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    reads = chimeric_reads(rng, genomes, 40, 150) + simulate_reads(rng, genomes, 300, (30, 150)) + chimeric_reads(rng, genomes, 60, 25)
    order = rng.permutation(len(reads))
    reads = [reads[i] for i in order]
    rb, ro = pack_sequences(reads)
    cls = Classifier(index)
    for conf in (0.0, 0.25):
        got = cls.classify(rb, ro, confidence=conf)
        res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=conf)
        assert max(len(h) for h in per) > 70
        assert_batch_equal(res, per, got, 35)
    got = cls.classify(rb, ro, rb, ro, confidence=0.1)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), rb, ro.astype(np.int64), confidence=0.1)
    assert_batch_equal(res, per, got, 35)
    cls.close(); index.close(); tax.close()


def test_packed_input_and_device_packer(gpu):
    """The packed input form (2-bit blocks + ambiguity mask) gives the same results as ASCII, single and paired, and
    the device-side stage-1 kernel packs exactly like the host packer."""
    rng, parents, ranks, names, genomes, taxa = make_world(55)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    r1 = simulate_reads(rng, genomes, 2000, (1, 200), n_rate=0.2) + [b"", b"N" * 70, b"acgu" * 20]
    r2 = simulate_reads(rng, genomes, len(r1), (20, 180), n_rate=0.1)
    b1, o1 = pack_sequences(r1)
    b2, o2 = pack_sequences(r2)
    p1, p2 = pack_reads(b1, o1), pack_reads(b2, o2)
    # device packer == host packer
    d_b, d_o = gpu.dev_alloc(max(len(b1), 16)), gpu.dev_alloc(o1.nbytes)
    gpu.h2d(d_b, b1); gpu.h2d(d_o, o1)
    boff = block_offsets(o1)
    d_boff, d_codes = gpu.dev_alloc(boff.nbytes), gpu.dev_alloc(max(p1.codes.nbytes, 16))
    d_mask, d_len = gpu.dev_alloc(max(p1.mask.nbytes, 16)), gpu.dev_alloc(p1.len.nbytes)
    gpu.h2d(d_boff, boff)
    gpu.pack_reads_dev(d_b, d_o, len(r1), d_boff, d_codes, d_mask, d_len)
    codes, mask, ln = np.zeros_like(p1.codes), np.zeros_like(p1.mask), np.zeros_like(p1.len)
    gpu.d2h(codes, d_codes); gpu.d2h(mask, d_mask); gpu.d2h(ln, d_len)
    assert np.array_equal(codes, p1.codes) and np.array_equal(mask, p1.mask) and np.array_equal(ln, p1.len)
    for d in (d_b, d_o, d_boff, d_codes, d_mask, d_len):
        gpu.dev_free(d)
    cls = Classifier(index)
    for conf in (0.0, 0.2):
        got = cls.classify_packed(p1, confidence=conf)
        res, _, _, per = olib.classify(b1, o1.astype(np.int64), confidence=conf)
        assert_batch_equal(res, per, got, 35)
        got = cls.classify_packed(p1, p2, confidence=conf)
        res, _, _, per = olib.classify(b1, o1.astype(np.int64), b2, o2.astype(np.int64), confidence=conf)
        assert_batch_equal(res, per, got, 35)
    got2 = cls.classify_packed(p1, p2, confidence=0.2, per_read_output=False)
    assert np.array_equal(got2.taxon, got.taxon) and np.array_equal(got2.flags, got.flags)
    cls.close(); index.close(); tax.close()


def test_classify_paired_end(gpu):
    rng, parents, ranks, names, genomes, taxa = make_world(33)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    r1 = simulate_reads(rng, genomes, 1500, (20, 180), n_rate=0.1)
    r2 = simulate_reads(rng, genomes, 1500, (20, 180), n_rate=0.1)
    r1[0], r2[0] = b"ACG", b"TTT"   # a pair without any span still yields U ... 34|34 |:|
    b1, o1 = pack_sequences(r1)
    b2, o2 = pack_sequences(r2)
    cls = Classifier(index)
    for conf in (0.0, 0.2):
        got = cls.classify(b1, o1, b2, o2, confidence=conf)
        res, _, _, per = olib.classify(b1, o1.astype(np.int64), b2, o2.astype(np.int64), confidence=conf)
        assert_batch_equal(res, per, got, 35)
    assert got.has_span.all()
    cls.close(); index.close(); tax.close()


@pytest.mark.parametrize("k,m,s,canonical", [(35, 31, 7, True), (31, 24, 0, True), (25, 20, 3, False), (28, 28, 4, True),
                                             (22, 15, 0, True)])
def test_build_and_classify_params(gpu, k, m, s, canonical):
    rng, parents, ranks, names, genomes, taxa = make_world(5 + k, n_genomes=16, glen=3000)
    p = oracle.params(k=k, m=m, spaces=s, canonical=canonical)
    olib = oracle_lib(p, parents, genomes, taxa)
    oid, otx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    params = IndexParams(k=k, m=m, spaces=s, canonical=canonical)
    gb, goff = pack_sequences(genomes)
    # two build batches, one of them with an undefined label that must be dropped
    half = len(genomes) // 2
    t2 = taxa.copy()
    extra = random_dna(rng, 500)
    b_a, o_a = pack_sequences(genomes[:half])
    b_b, o_b = pack_sequences(genomes[half:] + [extra])
    undefined = int(np.where((parents == 0) & (np.arange(len(parents)) > 1))[0][0]) if ((parents == 0) & (np.arange(len(parents)) > 1)).any() else 0
    index = KeyValueIndex.build(gpu, tax, params, [(b_a, o_a, t2[:half]), (b_b, o_b, np.append(t2[half:], undefined).astype(np.int32))],
                                expected_bases=len(gb))
    gid, gtx = index.records()
    assert np.array_equal(oid, gid)
    assert np.array_equal(otx, gtx)
    reads = simulate_reads(rng, genomes, 800, (k - 3, 150), n_rate=0.1)
    rb, ro = pack_sequences(reads)
    cls = Classifier(index)
    got = cls.classify(rb, ro, confidence=0.05)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), confidence=0.05)
    assert_batch_equal(res, per, got, k)
    cls.close(); index.close(); tax.close()


def test_report_counts_and_kreport(gpu):
    rng, parents, ranks, names, genomes, taxa = make_world(77)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    reads = simulate_reads(rng, genomes, 4000, (20, 150))
    rb, ro = pack_sequences(reads)
    cls = Classifier(index)
    counts = ReportCounts(gpu, tax, n_samples=2)
    cls.attach_counts(counts, 1)
    got = cls.classify(rb, ro, confidence=0.0)
    cls.attach_counts(None)
    res, _, _, _ = olib.classify(rb, ro.astype(np.int64), confidence=0.0, with_hits=False)
    hs = res["has_span"].astype(bool)
    ref = np.bincount(res["taxon"][hs], minlength=len(parents))
    assert np.array_equal(counts.fetch(1), ref)
    assert counts.fetch(0).sum() == 0
    # the host-fed path (slk_counts_add) gives the same vector
    counts.add(got.taxon, got.flags, np.zeros(len(reads), dtype=np.int32))
    assert np.array_equal(counts.fetch(0), ref)
    a = oracle.kraken_report(parents, ranks, names, [(int(t), int(c)) for t, c in enumerate(ref) if c])
    b = KrakenReport(parents, ranks, names, counts.pairs(1)).text()
    assert a == b
    counts.close(); cls.close(); index.close(); tax.close()


def test_synthetic_generators_match_oracle(gpu):
    n = 300000
    d = gpu.dev_alloc(n)
    import ctypes as C
    from slacken_b200._lib import check
    check(gpu._L.slk_synth_genome_dev(gpu.h, 42, 65000, n, C.c_void_p(d)))
    out = np.zeros(n, dtype=np.uint8)
    gpu.d2h(out, d)
    assert np.array_equal(out, oracle.synth_genome(42, 65000, n))
    gpu.dev_free(d)
    nr, L = 2000, 150
    d = gpu.dev_alloc(nr * L)
    check(gpu._L.slk_synth_reads_dev(gpu.h, 42, 43, 5, 100000, 7, nr, L, C.c_void_p(d)))
    out = np.zeros(nr * L, dtype=np.uint8)
    gpu.d2h(out, d)
    assert np.array_equal(out, oracle.synth_reads(42, 43, 5, 100000, 7, nr, L))
    check(gpu._L.slk_synth_mates_dev(gpu.h, 42, 43, 5, 100000, 7, nr, L, 1, C.c_void_p(d)))
    gpu.d2h(out, d)
    assert np.array_equal(out, oracle.synth_reads(42, 43, 5, 100000, 7, nr, L, mate=1))
    gpu.dev_free(d)


def test_chunked_pipeline_many_reads(gpu):
    """More reads than one internal chunk (2^19) so the 3-slot H2D / kernel / D2H pipeline wraps around."""
    rng, parents, ranks, names, genomes, taxa = make_world(91, n_genomes=8, glen=4000)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    base_reads = simulate_reads(rng, genomes, 5000, (40, 120))
    reads = base_reads * 260   # 1.3 M reads
    rb, ro = pack_sequences(reads)
    cls = Classifier(index)
    got = cls.classify(rb, ro, confidence=0.1)
    res, _, _, per = olib.classify(rb[:int(ro[5000])], ro[:5001].astype(np.int64), confidence=0.1)
    n = 5000
    for rep in (0, 129, 259):
        sl = slice(rep * n, (rep + 1) * n)
        assert np.array_equal(got.taxon[sl], res["taxon"])
        assert np.array_equal(got.flags[sl] & 1, res["classified"])
    for i in list(range(0, n, 501)):
        for rep in (0, 104, 105, 259):
            h = got.hits_of(rep * n + i)
            assert np.array_equal(h["taxon"], per[i]["taxon"]) and np.array_equal(h["count"], per[i]["count"])
    cls.close(); index.close(); tax.close()


@pytest.mark.parametrize("n", [0, 1, 31, 4096, 4097, 100_003, 3_000_000])
def test_radix_sort_matches_numpy(gpu, n):
    """K3a: the hand-written LSD radix sort against numpy, including stability on a partial bit range."""
    import ctypes as C
    from slacken_b200._lib import check
    rng = np.random.default_rng(n)
    keys = rng.integers(0, 2**63, n, dtype=np.int64).astype(np.uint64) ^ (rng.integers(0, 2, n).astype(np.uint64) << np.uint64(63))
    if n > 10:
        keys[: n // 3] = keys[n // 3: 2 * (n // 3)]          # duplicates
        keys[-5:] = [0, 2**64 - 1, 0, 2**64 - 1, 1 << 16]
    a = keys.copy()
    check(gpu._L.slk_debug_sort_u64(gpu.h, a.ctypes.data_as(C.c_void_p), n, 0, 64))
    assert np.array_equal(a, np.sort(keys))
    b = keys.copy()
    check(gpu._L.slk_debug_sort_u64(gpu.h, b.ctypes.data_as(C.c_void_p), n, 16, 64))
    want = keys[np.argsort(keys >> np.uint64(16), kind="stable")]   # low 16 bits keep their input order
    assert np.array_equal(b, want)
    c = keys.copy()
    check(gpu._L.slk_debug_sort_u64(gpu.h, c.ctypes.data_as(C.c_void_p), n, 16, 56))
    want = keys[np.argsort((keys >> np.uint64(16)) & np.uint64((1 << 40) - 1), kind="stable")]
    assert np.array_equal(c, want)


@pytest.mark.parametrize("n", [0, 1, 31, 33, 129])
def test_tiny_batches(gpu, n):
    """Batches that do not fill a warp / a block, and the empty batch, single and paired, both input forms."""
    rng, parents, ranks, names, genomes, taxa = make_world(71, n_genomes=6, glen=3000)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    cls = Classifier(index)
    r1 = simulate_reads(rng, genomes, n, (34, 160), n_rate=0.2)
    r2 = simulate_reads(rng, genomes, n, (1, 160), n_rate=0.2)
    if n >= 31:
        r1[0], r2[0] = b"", b""          # a pair of empty mates still has its border span
        r1[5], r2[7] = genomes[0][:35], genomes[1][:34]   # exactly k, exactly k-1
    b1, o1 = pack_sequences(r1)
    b2, o2 = pack_sequences(r2)
    for paired in (False, True):
        args = (b1, o1, b2, o2) if paired else (b1, o1)
        got = cls.classify(*args, confidence=0.1)
        oargs = (b1, o1.astype(np.int64), b2, o2.astype(np.int64)) if paired else (b1, o1.astype(np.int64))
        res, _, _, per = olib.classify(*oargs, confidence=0.1)
        assert len(got.taxon) == n
        if n:
            assert_batch_equal(res, per, got, 35)
            pk = (pack_reads(b1, o1), pack_reads(b2, o2)) if paired else (pack_reads(b1, o1),)
            got = cls.classify_packed(*pk, confidence=0.1)
            assert_batch_equal(res, per, got, 35)
    cls.close(); index.close(); tax.close()


def test_fp64_pipe_minimum_is_the_integer_minimum(gpu):
    """slk_min62 (slk_core.h): the scanner compares right-aligned priorities (< 2^62) as doubles on the device. Every
    bit pattern below 2^62 is a non-negative finite double (zero, subnormal or normal), so the result must equal the
    unsigned integer minimum, bit for bit: edge values, neighbours, and random pairs over every magnitude."""
    import ctypes as C
    from slacken_b200._lib import check
    rng = np.random.default_rng(5)
    edge = [0, 1, 2, 3, (1 << 52) - 1, 1 << 52, (1 << 52) + 1, (1 << 53) - 1, 1 << 53, (1 << 61) - 1, 1 << 61, (1 << 62) - 1,
            0x000fffffffffffff, 0x0010000000000000, 0x3fefffffffffffff, 0x3ff0000000000000, 0x3fffffffffffffff]
    a = [x for x in edge for _ in edge]
    b = [y for _ in edge for y in edge]
    for bits in range(1, 63):   # random pairs of every magnitude, and pairs that differ only in low / only in high bits
        x = rng.integers(0, 1 << bits, size=400, dtype=np.uint64)
        y = rng.integers(0, 1 << bits, size=400, dtype=np.uint64)
        a += x.tolist(); b += y.tolist()
        a += x.tolist(); b += (x ^ np.uint64(1)).tolist()
        a += x.tolist(); b += (x ^ (np.uint64(1) << np.uint64(bits - 1))).tolist()
    a, b = np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64)
    out = np.zeros(len(a), dtype=np.uint64)
    v = lambda z: z.ctypes.data_as(C.c_void_p)
    check(gpu._L.slk_debug_min62(gpu.h, v(a), v(b), len(a), v(out)))
    assert np.array_equal(out, np.minimum(a, b))


def test_line_breaks_inside_sequences(gpu):
    """kmers/minimizer/ShiftScanner.scala:113-120: line breaks inside a sequence (multi-line FASTA records) are skipped by the
    reference's scanner. The oracle restates that and sees the raw reads and genomes; on this side pack_sequences drops the
    line breaks while it assembles the device buffers, in the library build as in the classification."""
    rng, parents, ranks, names, genomes, taxa = make_world(27)
    p = oracle.params()

    def wrap(s, width):
        return b"\n".join(s[i:i + width] for i in range(0, len(s), width)) + (b"\r\n" if len(s) % 2 else b"")
    raw_genomes = [wrap(g, 60) for g in genomes]
    olib = oracle_lib(p, parents, raw_genomes, taxa)
    tax = Taxonomy(gpu, parents, ranks, names)
    gb, go = pack_sequences(raw_genomes)
    assert int(go[-1]) == sum(len(g) for g in genomes)
    index = KeyValueIndex.build(gpu, tax, IndexParams(), [(gb, go, taxa)], expected_bases=len(gb))
    oid, otx = olib.records()
    gid, gtx = index.records()
    assert np.array_equal(oid, gid) and np.array_equal(otx, gtx)
    # no ambiguous stretches in these reads: a line break INSIDE a run of N splits the run for the reference (each part
    # then counts its own k-mers, and parts shorter than k count none), whereas dropping the line break joins it -- the one
    # place where the two differ, see DESIGN.md section 9
    reads = simulate_reads(rng, genomes, 1500, (40, 260), n_rate=0.0)
    raw_reads = [wrap(r, int(rng.integers(7, 80))) for r in reads]
    rb, ro = pack_sequences(raw_reads)
    ob, oo = oracle.pack_sequences(raw_reads)
    cls = Classifier(index)
    got = cls.classify(rb, ro, confidence=0.05)
    res, _, _, per = olib.classify(ob, oo, confidence=0.05)
    assert_batch_equal(res, per, got, 35)
    cls.close(); index.close(); tax.close()
