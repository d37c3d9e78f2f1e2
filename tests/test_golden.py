"""Committed golden fixture (tests/golden/, made by make_golden.py from excerpts of the reference's own testData):
per-read output lines, kreports and library records for one single-end and one paired-end batch.
The CPU test keeps the oracle pinned to the fixture; the GPU test compares the CUDA path with the fixture directly."""
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load():
    man = json.load(open(os.path.join(GOLD, "manifest.json")))
    genomes = [l.strip() for l in open(os.path.join(GOLD, "genomes.fna")) if not l.startswith(">")]
    reads = [l.rstrip("\n").split("\t") for l in open(os.path.join(GOLD, "reads.tsv"))]
    return man, genomes, reads


def expected(name):
    lines = open(os.path.join(GOLD, f"{name}.lines.txt")).read().splitlines()
    return lines, open(os.path.join(GOLD, f"{name}_kreport.txt")).read()


def test_oracle_reproduces_golden_fixture():
    from oracle import oracle
    man, genomes, reads = load()
    parents = np.array(man["taxonomy"]["parents"], dtype=np.int32)
    ranks, names = man["taxonomy"]["ranks"], man["taxonomy"]["names"]
    lib = oracle.Library(oracle.params(**man["params"]), parents, 1 << 18)
    pieces, labels = oracle.remove_invalid(genomes, man["genome_taxa"])
    b, off = oracle.pack_sequences(pieces)
    lib.add_fragments(b, off, labels)
    id1, tx = lib.records()
    assert np.array_equal(id1, np.load(os.path.join(GOLD, "records_id1.npy")))
    assert np.array_equal(tx, np.load(os.path.join(GOLD, "records_taxon.npy")))
    rb, ro = oracle.pack_sequences([r[1] for r in reads])
    rb2, ro2 = oracle.pack_sequences([r[2] for r in reads])
    for name, run in man["runs"].items():
        res, _, _, per = lib.classify(rb, ro, rb2 if run["paired"] else None, ro2 if run["paired"] else None,
                                      confidence=run["confidence"])
        lines = [oracle.output_line(reads[i][0], res[i], [(int(h["taxon"]), int(h["count"])) for h in per[i]])
                 for i in range(len(reads)) if res["has_span"][i]]
        counts = np.bincount(res["taxon"][res["has_span"].astype(bool)], minlength=len(parents))
        rep = oracle.kraken_report(parents, ranks, names, [(int(t), int(c)) for t, c in enumerate(counts) if c])
        want_lines, want_rep = expected(name)
        assert lines == want_lines and rep == want_rep


@pytest.mark.gpu
def test_cuda_path_matches_golden_fixture(gpu):
    from slacken_b200 import Classifier, IndexParams, KeyValueIndex, KrakenReport, ReportCounts, Taxonomy
    from slacken_b200.host import pack_sequences
    from slacken_b200.report import output_line
    man, genomes, reads = load()
    parents = np.array(man["taxonomy"]["parents"], dtype=np.int32)
    ranks, names = man["taxonomy"]["ranks"], man["taxonomy"]["names"]
    tax = Taxonomy(gpu, parents, ranks, names)
    params = IndexParams(**man["params"])
    gb, goff = pack_sequences(genomes)   # the GPU build takes the genomes as they are (N runs included)
    index = KeyValueIndex.build(gpu, tax, params, [(gb, goff, np.array(man["genome_taxa"], dtype=np.int32))])
    id1, tx = index.records()
    assert np.array_equal(id1, np.load(os.path.join(GOLD, "records_id1.npy")))
    assert np.array_equal(tx, np.load(os.path.join(GOLD, "records_taxon.npy")))
    # the same library loaded from its records (the Parquet route) must behave identically
    index2 = KeyValueIndex.from_records(gpu, tax, params, id1, tx)
    rb, ro = pack_sequences([r[1] for r in reads])
    rb2, ro2 = pack_sequences([r[2] for r in reads])
    for ix in (index, index2):
        cls = Classifier(ix)
        counts = ReportCounts(gpu, tax)
        for name, run in man["runs"].items():
            counts.reset()
            cls.attach_counts(counts, 0)
            got = cls.classify(rb, ro, rb2 if run["paired"] else None, ro2 if run["paired"] else None,
                               confidence=run["confidence"])
            lines = [output_line(reads[i][0], got.taxon[i], got.classified[i], got.detail[i], got.hits_of(i))
                     for i in range(len(reads)) if got.has_span[i]]
            rep = KrakenReport(parents, ranks, names, counts.pairs(0)).text()
            want_lines, want_rep = expected(name)
            assert lines == want_lines
            assert rep == want_rep
        counts.close(); cls.close()
    index.close(); index2.close(); tax.close()
