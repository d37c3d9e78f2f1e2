"""The warp-cooperative classify kernel (slacken_b200/csrc/slk_group.h, classify2_kernel) through the packed entry points of
the C ABI, against the oracle: every per-read field, the merged hit lists and the report counters, over the shapes that
stress its pieces (chunk borders, buffer closes, mates without windows, long reads, many hits, ambiguous stretches, other
k / m / spaced-seed widths)."""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import Classifier, IndexParams, KeyValueIndex, ReportCounts, Taxonomy
from slacken_b200.host import pack_reads, pack_sequences
from tests.test_gpu_parity import assert_batch_equal, make_world, oracle_lib
from tests.util import chimeric_reads, random_dna, simulate_reads

pytestmark = pytest.mark.gpu


def _setup(gpu, seed, **params):
    rng, parents, ranks, names, genomes, taxa = make_world(seed)
    p = oracle.params(**params)
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(**{"k": p.k, "m": p.m, "spaces": params.get("spaces", 7),
                                                                 "canonical": params.get("canonical", True)}), id1, tx)
    return rng, parents, genomes, olib, tax, index


def _check(cls, olib, reads, mates=None, conf=0.0, k=35, counts=None, n_tax=0):
    rb, ro = pack_sequences(reads)
    mb = mo = None
    pk = [pack_reads(rb, ro)]
    if mates is not None:
        mb, mo = pack_sequences(mates)
        pk.append(pack_reads(mb, mo))
    if counts is not None:
        counts.reset()
    got = cls.classify_packed(*pk, confidence=conf)
    res, _, _, per = olib.classify(rb, ro.astype(np.int64), mb, mo.astype(np.int64) if mo is not None else None, confidence=conf)
    assert_batch_equal(res, per, got, k)
    if counts is not None:
        ref = np.bincount(res["taxon"][res["has_span"].astype(bool)], minlength=n_tax)
        assert np.array_equal(counts.fetch(0), ref)
    lean = cls.classify_packed(*pk, confidence=conf, per_read_output=False)
    assert np.array_equal(lean.taxon, got.taxon) and np.array_equal(lean.flags, got.flags)


@pytest.mark.parametrize("conf", [0.0, 0.15, 0.6])
def test_single_end_all_shapes(gpu, conf):
    rng, parents, genomes, olib, tax, index = _setup(gpu, 61)
    cls = Classifier(index)
    counts = ReportCounts(gpu, tax)
    cls.attach_counts(counts)
    reads = simulate_reads(rng, genomes, 4000, (1, 400), n_rate=0.15)
    reads += [b"", b"A", b"ACGT" * 8 + b"AC", b"ACGT" * 8 + b"ACG", b"N" * 34, b"N" * 35, b"N" * 200, b"A" * 500, b"AC" * 300,
              b"ACGTN" * 60, genomes[0][:35], genomes[0][:36], genomes[1][:50] + b"N" * 40 + genomes[1][90:150],
              genomes[2][100:134] + b"N" + genomes[2][135:400]]
    reads += [genomes[3][i:i + 150] for i in range(0, 64)]                      # every chunk phase of a run border
    reads += [genomes[4][:L] for L in range(30, 100)]                           # every tail length
    _check(cls, olib, reads, conf=conf, counts=counts, n_tax=len(parents))
    counts.close(); cls.close(); index.close(); tax.close()


def test_paired_end_including_mates_without_windows(gpu):
    rng, parents, genomes, olib, tax, index = _setup(gpu, 62)
    cls = Classifier(index)
    r1 = simulate_reads(rng, genomes, 2500, (1, 300), n_rate=0.1)
    r2 = simulate_reads(rng, genomes, 2500, (1, 300), n_rate=0.1)
    r1 += [b"", b"", genomes[0][:150], b"ACGT", b"N" * 100, genomes[1][:35]]
    r2 += [b"", genomes[0][200:350], b"", b"N" * 50, genomes[2][:150], genomes[1][35:70]]
    for conf in (0.0, 0.05, 0.3):
        _check(cls, olib, r1, r2, conf=conf)
    cls.close(); index.close(); tax.close()


def test_long_reads_and_many_hits(gpu):
    rng, parents, genomes, olib, tax, index = _setup(gpu, 63)
    cls = Classifier(index)
    reads = chimeric_reads(rng, genomes, 300, 40) + chimeric_reads(rng, genomes, 40, 200) + [genomes[0], genomes[1] + genomes[2]]
    reads += [random_dna(rng, 20000), b"A" * 3000 + genomes[3][:2000] + b"N" * 700 + genomes[4][:1500]]
    _check(cls, olib, reads, conf=0.1)
    _check(cls, olib, reads[:170], reads[170:340], conf=0.1)
    cls.close(); index.close(); tax.close()


@pytest.mark.parametrize("k,m,s,canonical", [(35, 31, 7, False), (31, 24, 0, True), (20, 15, 3, True), (28, 21, 5, True),
                                             (24, 24, 2, True), (38, 31, 9, True), (17, 16, 0, False), (12, 5, 1, True)])
def test_other_parameters(gpu, k, m, s, canonical):
    rng, parents, genomes, olib, tax, index = _setup(gpu, 64, k=k, m=m, spaces=s, canonical=canonical)
    cls = Classifier(index)
    reads = simulate_reads(rng, genomes, 1500, (1, 260), n_rate=0.1)
    mates = simulate_reads(rng, genomes, 1500, (1, 260), n_rate=0.1)
    _check(cls, olib, reads, conf=0.1, k=k)
    _check(cls, olib, reads, mates, conf=0.0, k=k)
    cls.close(); index.close(); tax.close()


def test_several_thresholds_in_one_pass_equal_separate_calls(gpu):
    """slk_classify_batch_packed_multi (Classifier.scala:156-170): row t == a call with threshold t; hits are shared."""
    rng, parents, genomes, olib, tax, index = _setup(gpu, 65)
    cls = Classifier(index)
    reads = simulate_reads(rng, genomes, 3000, (20, 300), n_rate=0.1)
    mates = simulate_reads(rng, genomes, 3000, (20, 300), n_rate=0.1)
    thr = [0.0, 0.05, 0.15, 0.3, 0.6, 1.0]
    for pk in ([pack_reads(*pack_sequences(reads))], [pack_reads(*pack_sequences(reads)), pack_reads(*pack_sequences(mates))]):
        r2 = pk[1] if len(pk) > 1 else None
        taxon, flags, detail, hits, used = cls.classify_packed_thresholds(pk[0], r2, thr)
        for t, conf in enumerate(thr):
            one = cls.classify_packed(*pk, confidence=conf)
            assert np.array_equal(taxon[t], one.taxon) and np.array_equal(flags[t], one.flags), conf
            assert np.array_equal(detail["hit_cnt"], one.detail["hit_cnt"]) and used == one.hits_used
    cls.close(); index.close(); tax.close()


def test_compact_boundary_equals_the_oracle(gpu):
    """slk_classify_batch_compact: codes + lengths + sparse ambiguity list in, 16-byte results + hits in read order out."""
    from slacken_b200.host import compact_reads
    rng, parents, genomes, olib, tax, index = _setup(gpu, 66)
    cls = Classifier(index)
    reads = simulate_reads(rng, genomes, 5000, (1, 400), n_rate=0.2) + [b"", b"N" * 90, genomes[0][:200]] + chimeric_reads(rng, genomes, 50, 120)
    mates = simulate_reads(rng, genomes, len(reads), (1, 400), n_rate=0.2)
    for paired in (False, True):
        rb, ro = pack_sequences(reads)
        r1 = compact_reads(pack_reads(rb, ro))
        mb = mo = r2 = None
        if paired:
            mb, mo = pack_sequences(mates)
            r2 = compact_reads(pack_reads(mb, mo))
        thr = [0.1, 0.0, 0.5]
        got = cls.classify_compact(r1, r2, thresholds=thr)
        for t, conf in enumerate(thr):
            res, _, _, per = olib.classify(rb, ro.astype(np.int64), mb, mo.astype(np.int64) if mo is not None else None, confidence=conf)
            tx = got.taxon if t == 0 else got.taxon_more[t - 1]
            fl = got.flags if t == 0 else got.flags_more[t - 1]
            assert np.array_equal(res["taxon"], tx), conf
            assert np.array_equal(res["classified"], fl & 1) and np.array_equal(res["has_span"], (fl >> 1) & 1)
        hs = res["has_span"].astype(bool)
        assert np.array_equal(res["len1"][hs], got.results["len1"][hs].astype(np.int32))
        if paired:
            assert np.array_equal(res["len2"][hs], got.results["len2"][hs].astype(np.int32))
        # hits in read order: read i's hits follow read i - 1's
        cnt = got.hit_cnt.astype(np.int64)
        assert np.array_equal(cnt, res["n_hits"].astype(np.int64)) and got.hits_used == int(cnt.sum())
        off = np.concatenate([[0], np.cumsum(cnt)])
        for i in range(len(per)):
            h = got.hits[off[i]:off[i + 1]]
            assert np.array_equal(h["taxon"], per[i]["taxon"]) and np.array_equal(h["count"], per[i]["count"]), i
        lean = cls.classify_compact(r1, r2, thresholds=thr[:1], per_read_output=False)
        assert np.array_equal(lean.taxon, got.taxon) and np.array_equal(lean.flags, got.flags)
        # the same with 4-byte hits (slk_classify_batch_compact_short)
        short = cls.classify_compact(r1, r2, thresholds=thr[:1], short_hits=True)
        assert short.hits.dtype == np.uint32 and short.hits_used == got.hits_used and short.results.itemsize == 8
        assert np.array_equal(short.results["taxon"], got.results["taxon"])
        assert np.array_equal(short.results["hits_flags"], got.results["hits_flags"])
        dec = short.decode_short_hits(index.taxa(), 35)
        l1, l2 = short.lengths_from_hits(dec, 35, paired)   # the lengths the 8-byte results leave out
        assert np.array_equal(l1[hs], got.results["len1"][hs]) and np.array_equal(l2[hs], got.results["len2"][hs])
        assert np.array_equal(dec["taxon"], got.hits[:got.hits_used]["taxon"])
        assert np.array_equal(dec["count"], got.hits[:got.hits_used]["count"])
        if paired:
            assert (dec["taxon"] == -2).sum() > 0 and (dec["taxon"] == -1).sum() > 0
    cls.close(); index.close(); tax.close()
