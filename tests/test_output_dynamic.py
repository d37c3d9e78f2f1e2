"""Host logic of SURVEY section 8 rows f2 (output layout) and f3 (taxon-set step of the 2-step mode): no GPU needed."""
import gzip
import os

import numpy as np

from slacken_b200.dynamic import RANK_DEPTH, TaxonomyTree, count_filter
from slacken_b200.host import DETAIL_DTYPE, HIT_DTYPE, ClassifiedBatch
from slacken_b200.output import ClassificationWriter, sample_ids, threshold_string


class _Tax:   # what ClassificationWriter / TaxonomyTree need of a Taxonomy, without a device
    def __init__(self, parents, ranks, names):
        self.parents, self.ranks, self.names = np.asarray(parents, dtype=np.int32), ranks, names


def _tree():
    #            1 root
    #        2 superkingdom
    #     3 genus          10 species
    #   4 species  5 species
    #  6,7,9 strain  8 strain
    parents = [0, 0, 1, 2, 3, 3, 4, 4, 5, 4, 2]
    ranks = [None, "root", "superkingdom", "genus", "species", "species", None, None, None, None, "species"]
    names = [f"t{i}" for i in range(11)]
    names[0] = "unclassified"
    return _Tax(parents, ranks, names)


def test_threshold_string_uses_the_longest_decimal_part():
    assert threshold_string(0.0, [0.0]) == "0.0"
    assert threshold_string(0.0, [0.0, 0.15]) == "0.00"
    assert threshold_string(0.15, [0.0, 0.15]) == "0.15"
    assert threshold_string(0.5, [0.05, 0.5, 0.125]) == "0.500"


def test_sample_ids():
    assert sample_ids(["a", "b"], None) == ["all", "all"]
    assert sample_ids(["S1_read7", "x", "S22_r"], r"(S[0-9]+)_") == ["S1", "other", "S22"]


def test_depth_descendants_and_count_filter():
    tree = TaxonomyTree(_tree())
    assert [tree.depth(t) for t in (1, 2, 3, 4, 6, 8, 10)] == [0, 1, 7, 8, 8, 8, 8]   # unranked strains inherit 'species'
    assert tree.depth(0) == -1
    assert tree.with_descendants([4]) == {4, 6, 7, 9}
    assert tree.with_descendants([3, 10]) == {3, 4, 5, 6, 7, 8, 9, 10}
    counts = [(6, 40), (7, 30), (8, 5), (3, 100), (10, 99), (0, 1000)]
    # species and below with clade total >= 50: 6 and 7 sit below species 4 but are keys with clade totals 40 / 30 only
    assert count_filter(tree, counts, "species", 50) == {10}
    assert count_filter(tree, counts, "species", 30) == {6, 7, 10}
    assert count_filter(tree, counts, "genus", 100) == {3}          # clade(3) = 40 + 30 + 5 + 100
    assert RANK_DEPTH["species"] == 8


def test_writer_layout_and_report(tmp_path):
    tax = _tree()
    n = 5
    taxon = np.array([6, 0, 4, 6, 0], dtype=np.int32)
    flags = np.array([3, 2, 3, 3, 0], dtype=np.uint8)          # last read has no span: it vanishes
    detail = np.zeros(n, dtype=DETAIL_DTYPE)
    detail["len1"] = 150
    detail["len2"] = 0xFFFFFFFF
    detail["hit_off"] = [0, 2, 3, 4, 6]
    detail["hit_cnt"] = [2, 1, 1, 2, 0]
    hits = np.array([(6, 100), (0, 16), (0, 116), (4, 116), (-1, 3), (6, 113)], dtype=HIT_DTYPE)
    batch = ClassifiedBatch(taxon, flags, detail, hits, len(hits))
    titles = ["A_r1", "A_r2", "B_r1", "zzz", "A_r5"]
    w = ClassificationWriter(tax, str(tmp_path / "out"), 0.15, [0.0, 0.15], sample_regex=r"^([AB])_")
    w.add(titles, batch)
    assert w.close() == ["A", "B", "other"]
    loc = str(tmp_path / "out") + "_c0.15"
    lines = gzip.open(os.path.join(loc, "sample=A", "part-00000.txt.gz"), "rt").read().splitlines()
    assert lines == ["C\tA_r1\t6\t150\t6:100 0:16", "U\tA_r2\t0\t150\t0:116"]
    assert gzip.open(os.path.join(loc, "sample=other", "part-00000.txt.gz"), "rt").read() == "C\tzzz\t6\t150\tA:3 6:113\n"
    rep = open(os.path.join(loc, "A_kreport.txt")).read().splitlines()
    assert rep[0] == "#Perc\tAggregate\tIn taxon\tRank\tTaxon\tName"
    assert rep[1] == " 50.00\t1\t1\tU\t0\tunclassified"
    assert rep[2] == " 50.00\t1\t0\tR\t1\tt1"
    assert rep[-1] == " 50.00\t1\t1\tS1\t6\t        t6"
    # --nounclassified drops the U rows from lines and reports
    w2 = ClassificationWriter(tax, str(tmp_path / "o2"), 0.0, [0.0], with_unclassified=False)
    w2.add(titles, batch)
    assert w2.close() == ["all"]
    rep2 = open(str(tmp_path / "o2") + "_c0.0/all_kreport.txt").read()
    assert "unclassified" not in rep2 and rep2.splitlines()[1].startswith("100.00\t3\t0\tR\t1")


def test_gold_set_promotion_filter_and_comparison():
    """readGoldSet + the comparison of findTaxonSet (slacken/Dynamic.scala:265-275,282-318), host logic only."""
    from slacken_b200.dynamic import Dynamic, GoldSetOptions, format_perc

    class _Base:
        taxonomy = _tree()
        params = None

    # the library has sequence for strains 6 and 8 and species 10 (plus, implicitly, all their ancestors)
    genomes = [(6, b""), (8, b""), (10, b"")]
    gold = [6, 7, 9, 10, 3]   # 7 and 9 have no sequence: both are promoted to species 4; 3 (genus) is filtered at species
    d = Dynamic(None, _Base(), genomes, rank="species", gold_set_opts=GoldSetOptions(gold))
    assert d.taxon_set_in_library() == {1, 2, 3, 4, 5, 6, 8, 10}
    assert d.read_gold_set() == {4, 6, 7, 9, 10}
    assert d.log[:3] == ["Gold set contained 5 taxa", "2 taxa from gold set not found in library, promoted to 1 taxa.",
                         "Initial adjusted gold set size 6, filtered at species to 5"]
    # at rank genus nothing is filtered; promote_rank keeps promoted taxa at that rank or below even if the filter drops them
    assert Dynamic(None, _Base(), genomes, rank="genus", gold_set_opts=GoldSetOptions(gold)).read_gold_set() == {3, 4, 6, 7, 9, 10}
    d2 = Dynamic(None, _Base(), [(3, b"")], rank="species", gold_set_opts=GoldSetOptions([6, 10], promote_rank="genus"))
    assert d2.read_gold_set() == {6, 10, 3}        # 6 -> promoted to genus 3 (kept by promote_rank), 10 -> superkingdom 2 (dropped)
    d3 = Dynamic(None, _Base(), [(3, b"")], rank="species", gold_set_opts=GoldSetOptions([6, 10]))
    assert d3.read_gold_set() == {6, 10}
    # merged.dmp mapping and a gold-set file
    import tempfile
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        f.write("9\n10,extra\n")
    primary = list(range(11))
    primary[9] = 6
    d4 = Dynamic(None, _Base(), genomes, rank="species", gold_set_opts=GoldSetOptions(f.name), primary=primary)
    assert d4.read_gold_set() == {6, 10}
    # comparison: detected {6, 5} against gold {4, 6, 7, 9, 10}
    st = d.compare_with_gold_set({6, 5})
    assert (st["tp"], st["fp"], st["fn"]) == (1, 1, 4)
    assert d.log[-1].endswith("True Positives: 1, False Positives: 1, False Negatives: 4, Precision: 50.00%, Recall: 20.00%")
    assert format_perc(1 / 3) == "33.33%"


def test_threshold_directory_names_follow_java_double_to_string():
    """Classifier.scala:189-190 takes the decimal count from Double.toString, which switches to scientific notation below 1e-3:
    1.0E-5 -> "0E-5" -> four decimals."""
    from slacken_b200.output import java_double_to_string, threshold_string
    cases = {0.0: "0.0", 0.15: "0.15", 1e-5: "1.0E-5", 1e-4: "1.0E-4", 0.001: "0.001", 1.5e-7: "1.5E-7", 1e7: "1.0E7",
             1234567.0: "1234567.0", 100.0: "100.0", 12345678.9: "1.23456789E7", 0.05: "0.05"}
    for x, want in cases.items():
        assert java_double_to_string(x) == want
    assert threshold_string(1e-5, [1e-5]) == "0.0000"
    assert threshold_string(0.0, [0.0, 0.15]) == "0.00" and threshold_string(0.15, [0.0, 0.15]) == "0.15"
    assert threshold_string(0.1, [0.1, 1e-4]) == "0.1000"


def test_taxonomy_loader_returns_the_merged_mapping(tmp_path):
    from slacken_b200 import library_io as lio
    d = tmp_path / "tax"
    d.mkdir()
    (d / "nodes.dmp").write_text("1\t|\t1\t|\tno rank\t|\n2\t|\t1\t|\tsuperkingdom\t|\n7\t|\t2\t|\tspecies\t|\n")
    (d / "names.dmp").write_text("1\t|\troot\t|\t\t|\tscientific name\t|\n7\t|\tSeven\t|\t\t|\tscientific name\t|\n99\t|\tGhost\t|\t\t|\tscientific name\t|\n")
    (d / "merged.dmp").write_text("12\t|\t7\t|\n")
    parents, ranks, names, primary = lio.load_taxonomy_primary(str(d))
    assert len(parents) == 13 and parents[7] == 2 and names[7] == "Seven"      # id 99 of names.dmp is unknown: ignored
    assert primary[12] == 7 and primary[7] == 7 and primary[2] == 2
    assert lio.load_taxonomy_dmp(str(d))[0].tolist() == parents.tolist()


def test_short_hit_words_and_the_lengths_they_imply():
    """The 4-byte hit format of slk_classify_batch_compact_short (include/slacken_gpu.h) and the two length fields that its
    8-byte results leave out: decoded on the host, they must give back the (taxon, count) records and len1 / len2."""
    import numpy as np
    from slacken_b200.host import HIT_DTYPE, RESULT_SHORT_DTYPE, CompactBatch
    k = 35
    taxa = np.array([1, 7, 12, 40], dtype=np.int32)                     # slk_index_taxa: label i + 1 -> taxa[i]
    # read 0: single mate; read 1: no hits; read 2: pair with a border; read 3: ambiguous only
    want = np.zeros(8, dtype=HIT_DTYPE)
    want["taxon"] = [7, 0, 12, -2, 0, 40, -1, -1]
    want["count"] = [10, 6, 20, -(k - 1), 3, 4, 2, 65534]
    words = np.zeros(8, dtype=np.uint32)
    label = {0: 0, 1: 1, 7: 2, 12: 3, 40: 4, -1: 0xFFFF}
    for i, (t, c) in enumerate(zip(want["taxon"], want["count"])):
        words[i] = 0xFFFFFFFF if t == -2 else (label[int(t)] << 16) | int(c)
    res = np.zeros(4, dtype=RESULT_SHORT_DTYPE)
    res["hits_flags"] = np.array([2, 0, 4, 2], dtype=np.uint32) << 2
    b = CompactBatch(res, None, None, words, 8)
    got = b.decode_short_hits(taxa, k)
    assert np.array_equal(got["taxon"], want["taxon"]) and np.array_equal(got["count"], want["count"])
    l1, l2 = b.lengths_from_hits(got, k, True)
    assert l1.tolist() == [16 + k - 1, k - 1, 20 + k - 1, 65536 + k - 1] and l2.tolist() == [k - 1, k - 1, 7 + k - 1, k - 1]
    l1s, l2s = b.lengths_from_hits(got, k, False)
    assert l2s.tolist() == [0xFFFFFFFF] * 4 and l1s.tolist() == l1.tolist()


def test_pack_sequences_drops_line_breaks():
    """host.pack_sequences assembles the device input: one sequence per offset range, line breaks inside a sequence dropped
    (the reference's scanner skips them, kmers/minimizer/ShiftScanner.scala:113-120)."""
    import numpy as np
    from slacken_b200.host import pack_sequences
    b, off = pack_sequences([b"ACGT\nACG\r\nT", "NN\nAC", b"", b"\n"])
    assert bytes(b) == b"ACGTACGTNNAC" and off.tolist() == [0, 8, 12, 12, 12] and off.dtype == np.uint64
