"""SURVEY section 8 row f4 on the GPU: Bracken weights through the C ABI against the oracle's restatement of
slacken/BrackenWeights.scala (genomes with ambiguous stretches, pieces cut with read_len - 1 overlap)."""
import numpy as np
import pytest

from oracle import oracle
from slacken_b200 import IndexParams, KeyValueIndex, Taxonomy
from slacken_b200.bracken import BrackenWeights, kmer_distrib_lines
from tests.test_gpu_parity import make_world, oracle_lib

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("read_len,fragment_max", [(100, 1 << 20), (60, 2000)])
def test_bracken_weights_match_the_oracle(gpu, read_len, fragment_max):
    rng, parents, ranks, names, genomes, taxa = make_world(61, n_genomes=12, glen=6000)
    g = bytearray(genomes[5]); g[2000:2100] = b"N" * 100; g[4000] = ord("N"); genomes[5] = bytes(g)
    p = oracle.params()
    olib = oracle_lib(p, parents, genomes, taxa)
    id1, tx = olib.records()
    tax = Taxonomy(gpu, parents, ranks, names)
    index = KeyValueIndex.from_records(gpu, tax, IndexParams(), id1, tx)
    lib = list(zip(taxa.tolist(), genomes))
    got = BrackenWeights(index, read_len).build(lib, fragment_max=fragment_max)
    want = oracle.bracken_weights(p, parents, olib.lookup, lib, read_len, fragment_max=fragment_max)
    assert got == want
    assert sum(got.values()) == sum(len(s) - read_len + 1 for s in genomes)
    assert kmer_distrib_lines(got) == oracle.kmer_distrib_lines(want)
    index.close(); tax.close()
