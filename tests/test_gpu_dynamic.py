"""SURVEY section 8 row f3 on the GPU: the 2-step (dynamic library) mode of slacken/Dynamic.scala:250-374 over the
kernels, against the oracle: the three taxon-set heuristics, the taxon-set file, the rebuilt library and the second
classification."""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import Classifier, IndexParams, KeyValueIndex, Taxonomy
from slacken_b200.dynamic import (ClassifiedReadCount, Dynamic, MinimizerDistinctCount, MinimizerTotalCount, TaxonomyTree,
                                  count_filter)
from slacken_b200.host import pack_sequences
from tests.test_gpu_parity import assert_batch_equal, make_world, oracle_lib
from tests.util import simulate_reads

pytestmark = pytest.mark.gpu


def _setup(gpu):
    rng, parents, ranks, names, genomes, taxa = make_world(53, n_genomes=16, related=False)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    base = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    present = [0, 3, 7]                                           # the sample only holds reads of three genomes
    reads = simulate_reads(rng, [genomes[i] for i in present], 900, (60, 200))
    rb, ro = pack_sequences(reads)
    return rng, parents, genomes, taxa, p, olib, tax, base, rb, ro


def _oracle_span_hits(p, olib, reads_b, reads_o):
    """(taxon, minimizer) of every sequence span with a record: what findHitsWithMinimizers yields."""
    out = []
    for i in range(len(reads_o) - 1):
        for minimizer, _distinct, _kmers, flag in oracle.spans(p, bytes(reads_b[int(reads_o[i]):int(reads_o[i + 1])])):
            if flag == 1:   # SEQUENCE_FLAG
                t = olib.lookup(minimizer)
                if t:
                    out.append((t, minimizer))
    return out


@pytest.mark.parametrize("criteria", [ClassifiedReadCount(20, 0.05), MinimizerTotalCount(200), MinimizerDistinctCount(150)])
def test_two_step_dynamic_library(gpu, tmp_path, criteria):
    rng, parents, genomes, taxa, p, olib, tax, base, rb, ro = _setup(gpu)
    tree = TaxonomyTree(tax)
    rank = "species"
    dyn = Dynamic(gpu, base, list(zip(taxa.tolist(), genomes)), rank=rank, criteria=criteria)
    loc = str(tmp_path / "out_taxonSet.txt")
    taxon_set, index = dyn.make_index(rb, ro, write_location=loc)
    # the same heuristic from the oracle's results
    if isinstance(criteria, ClassifiedReadCount):
        res, _, _, _ = olib.classify(rb, ro.astype(np.int64), confidence=criteria.confidence, with_hits=False)
        t = res["taxon"][res["classified"].astype(bool)]
        u, c = np.unique(t, return_counts=True)
        counts = list(zip(u.tolist(), c.tolist()))
    else:
        hits = [(t, m) for t, m in _oracle_span_hits(p, olib, rb, ro) if tree.depth(t) >= 8]
        if isinstance(criteria, MinimizerDistinctCount):
            hits = list(set(hits))
        u, c = np.unique([t for t, _ in hits], return_counts=True)
        counts = list(zip(u.tolist(), c.tolist()))
    want = count_filter(tree, counts, rank, criteria.threshold)
    assert len(want) > 0
    assert [int(x) for x in open(loc).read().split()] == sorted(want)
    assert taxon_set == tree.with_descendants(want)
    # the dynamic library = the oracle's library over the genomes of the set
    chosen = [i for i in range(len(genomes)) if int(taxa[i]) in taxon_set]
    assert 0 < len(chosen) < len(genomes)
    dlib = oracle_lib(p, parents, [genomes[i] for i in chosen], taxa[chosen])
    oid, otx = dlib.records()
    gid, gtx = index.records()
    assert np.array_equal(oid, gid) and np.array_equal(otx, gtx)
    # ... and the second pass classifies like the oracle does with that library
    got = Classifier(index).classify(rb, ro, confidence=0.1)
    res, _, _, per = dlib.classify(rb, ro.astype(np.int64), confidence=0.1)
    assert_batch_equal(res, per, got, 35)
    index.close(); base.close(); tax.close()


def test_gold_set_library_and_support_reports(gpu, tmp_path):
    """The gold-set variant of makeRecords (slacken/Dynamic.scala:362-370) and the four Kraken-style support reports of
    reportDynamicIndexSupport (slacken/Dynamic.scala:146-180,210-230), against the oracle."""
    from slacken_b200.dynamic import GoldSetOptions
    from slacken_b200.report import KrakenReport
    rng, parents, genomes, taxa, p, olib, tax, base, rb, ro = _setup(gpu)
    tree = TaxonomyTree(tax)
    gold = [int(taxa[0]), int(taxa[3]), int(parents[int(taxa[7])])]
    dyn = Dynamic(gpu, base, list(zip(taxa.tolist(), genomes)), rank="species", criteria=ClassifiedReadCount(20, 0.05),
                  gold_set_opts=GoldSetOptions(gold, classify_with=True))
    want_set = tree.with_descendants(dyn.read_gold_set())
    taxon_set, index = dyn.make_index(rb, ro)
    assert taxon_set == want_set
    chosen = [i for i in range(len(genomes)) if int(taxa[i]) in taxon_set]
    assert len(chosen) >= 2
    dlib = oracle_lib(p, parents, [genomes[i] for i in chosen], taxa[chosen])
    oid, otx = dlib.records()
    gid, gtx = index.records()
    assert np.array_equal(oid, gid) and np.array_equal(otx, gtx)
    index.close()
    # compare-only gold set: the detected set is unchanged, the statistics are logged
    dyn2 = Dynamic(gpu, base, list(zip(taxa.tolist(), genomes)), rank="species", criteria=ClassifiedReadCount(20, 0.05),
                   gold_set_opts=GoldSetOptions(gold))
    dyn2.find_taxon_set(rb, ro)
    assert any(m.startswith("Comparing detected set with supplied gold set.") for m in dyn2.log)
    # support reports
    out = str(tmp_path / "dyn")
    files = dyn2.report_dynamic_index_support(out, rb, ro)
    assert [f[len(out):] for f in files] == ["_support_report_totalKmerCount.txt", "_support_report_distinctMinimizerCount.txt",
                                             "_support_report_totalMinimizerCount.txt", "_support_report_classifiedReadCount.txt",
                                             "_support_report_minimizerCoverage", "_support_report_minimizerDistinctCoverage"]
    # minimizerCoverage (IndexStatistics.showTaxonFullCoverageStats): super-mers of every library genome per depth of the
    # record's taxon, from the oracle's superkmerPositions and records
    cov_all, cov_dist = {}, {}
    pieces, labels = oracle.remove_invalid(genomes, taxa)
    seen = set()
    for piece, t in zip(pieces, labels.tolist()):
        if len(piece) < p.k:
            continue
        for _loc, rank_, _len in oracle.superkmers(p, piece):
            lca = olib.lookup(rank_)
            if lca:
                d = tree.depth(lca)
                cov_all[(t, d)] = cov_all.get((t, d), 0) + 1
                if (t, rank_) not in seen:
                    seen.add((t, rank_))
                    cov_dist[(t, d)] = cov_dist.get((t, d), 0) + 1
    for f, cov in zip(files[4:], (cov_all, cov_dist)):
        want = "".join(f"{t}  " + "|".join(f"{d}:{cov[(tt, d)]}" for tt, d in sorted(cov) if tt == t) + "\n"
                       for t in sorted({tt for tt, _ in cov}))
        assert open(f + "/part-00000.txt").read() == want and len(want) > 0, f
    files = files[:4]
    kmers, mins, dist = {}, {}, {}
    for i in range(len(ro) - 1):
        for minimizer, _distinct, n_kmers, flag in oracle.spans(p, bytes(rb[int(ro[i]):int(ro[i + 1])])):
            if flag == 1:
                t = olib.lookup(minimizer)
                if t and tree.depth(t) >= 8:
                    kmers[t] = kmers.get(t, 0) + n_kmers
                    mins[t] = mins.get(t, 0) + 1
                    dist.setdefault(t, set()).add(minimizer)
    res, _, _, _ = olib.classify(rb, ro.astype(np.int64), confidence=0.0, with_hits=False)
    u, c = np.unique(res["taxon"][res["classified"].astype(bool)], return_counts=True)
    for f, counts in zip(files, [sorted(kmers.items()), sorted((t, len(s)) for t, s in dist.items()), sorted(mins.items()),
                                 list(zip(u.tolist(), c.tolist()))]):
        assert open(f).read() == KrakenReport(tax.parents, tax.ranks, tax.names, counts).text(), f
        assert len(counts) > 0
    base.close(); tax.close()
