"""The two reference-held invariants of the build/classify path, transcribed onto the CUDA path (through the C ABI):

* src/test/scala/com/jnpersson/slacken/ClassifierTest.scala:75-130 ("Classify with random genomes"): reads simulated from
  random genomes on the leaves of a random taxonomy classify to the source taxon or to one of its ancestors (or stay
  unclassified), for random k, m and spaced-seed widths, minHitGroups = 1, confidence 0;
* src/test/scala/com/jnpersson/slacken/KeyValueIndexTest.scala:55-70 ("Insert random genomes and check index contents"):
  the number of records built from one random genome equals its number of distinct minimizers, all labelled with its taxon.

The reference draws k from [15, 158] and m from [15, 128]; the CUDA path supports m <= 31, k - m + 1 <= 8 and at most 48
significant minimizer bits (limits), so the draws are restricted to that envelope. These are properties the
reference's own tests state; nothing here depends on the oracle except the distinct-minimizer count of the second test,
which uses the oracle's restatement of MinSplitter.superkmerPositions (kmers/minimizer/MinSplitter.scala:133-216).
"""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import Classifier, IndexParams, KeyValueIndex, Taxonomy
from slacken_b200.host import pack_sequences
from tests.util import leaf_taxa, make_taxonomy, random_dna

pytestmark = pytest.mark.gpu


def _supported(k, m, s):
    bits = 2 * m - 2 * s if s > 0 else 2 * m
    return 1 <= m <= 31 and m <= k and k - m + 1 <= 8 and 0 <= s <= m // 2 and bits <= 48


def _has_ancestor(parents, tax, anc):
    """Taxonomy.hasAncestor (slacken/Taxonomy.scala:236-244): true also when tax == anc."""
    while tax != 0:
        if tax == anc:
            return True
        tax = int(parents[tax])
    return False


def test_simulated_reads_classify_to_their_taxon_or_an_ancestor(gpu):
    rng = np.random.default_rng(2024)
    n_genomes = 100
    parents, ranks, names = make_taxonomy(n_genomes * 8, 77)          # Testing.taxonomies(numberOfGenomes * 8)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, int(rng.integers(1000, 10001))) for _ in leaves]   # dnaStrings(1000, 10000), one per leaf
    taxa = np.array(leaves, dtype=np.int32)
    # simulateReads(200, 1000): exact substrings of length 200, the source taxon kept aside (the reference keeps it in the title)
    src = rng.integers(0, len(genomes), size=1000)
    reads = []
    for g in src:
        start = int(rng.integers(0, len(genomes[g]) - 200))
        reads.append(genomes[g][start:start + 200])
    expected = taxa[src]
    rb, ro = pack_sequences(reads)
    gb, go = pack_sequences(genomes)
    tax = Taxonomy(gpu, parents, ranks, names)
    tried = 0
    while tried < 10:                                                # minSuccessful(5) x minSuccessful(2)
        m = int(rng.integers(15, 32))
        k = int(rng.integers(m, m + 8))
        s = int(rng.integers(0, m // 2 + 1))                         # seedMaskSpaces(m)
        if not _supported(k, m, s):
            continue
        tried += 1
        params = IndexParams(k=k, m=m, spaces=s)
        index = KeyValueIndex.build(gpu, tax, params, [(gb, go, taxa)], expected_bases=len(gb))
        cls = Classifier(index)
        got = cls.classify(rb, ro, confidence=0.0, min_hit_groups=1)  # ClassifyParams(1, withUnclassified = true)
        bad = [i for i in range(len(reads))
               if got.classified[i] and not _has_ancestor(parents, int(expected[i]), int(got.taxon[i]))]
        assert not bad, (k, m, s, bad[:5])
        # exact substrings of a library genome: every k-mer is known, so (unlike the reference, which only rules out wrong
        # answers) nearly every read must in fact classify
        assert got.classified.mean() > 0.95, (k, m, s, float(got.classified.mean()))
        cls.close()
        index.close()
    tax.close()


def test_record_count_equals_distinct_minimizers_of_a_random_genome(gpu):
    rng = np.random.default_rng(5)
    parents = np.array([0, 0], dtype=np.int32)                        # the genome is labelled with taxon 1 (ROOT)
    tax = Taxonomy(gpu, parents)
    done = 0
    while done < 40:
        k = int(rng.integers(1, 92)) | 1                              # ks: odd, 1..91
        m = int(rng.integers(1, k + 1))                               # ms(k)
        s = m // 3
        if not _supported(k, m, s):
            continue
        done += 1
        x = random_dna(rng, int(rng.integers(k, 1001)))               # dnaStrings(k, 1000)
        p = oracle.params(k=k, m=m, spaces=s)
        ranks_ = {int(r) for _, r, _ in oracle.superkmers(p, x)}      # superkmerPositions(x).map(_.rank).distinct
        gb, go = pack_sequences([x])
        index = KeyValueIndex.build(gpu, tax, IndexParams(k=k, m=m, spaces=s), [(gb, go, np.array([1], dtype=np.int32))],
                                    expected_bases=len(gb))
        id1, tx = index.records()
        assert len(id1) == len(ranks_), (k, m, s, len(x))
        assert set(int(v) for v in id1) == ranks_
        assert (tx == 1).all()
        index.close()
    tax.close()
