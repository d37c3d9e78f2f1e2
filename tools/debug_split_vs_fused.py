"""Debug helper: fused (hits on / off) vs split path vs oracle on a mid-size synthetic workload, one GPU."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import bench_workload as bw
from oracle import oracle
from slacken_b200 import Classifier, GpuContext, IndexParams, KeyValueIndex, Taxonomy
from slacken_b200._lib import check
from slacken_b200.sharded import ShardedClassifier, ShardedKeyValueIndex

w = bw.Workload(); w.n_genomes, w.n_reads = 40, 200000
ctx = GpuContext(0)
parents, ranks, names, genome_taxa = bw.taxonomy(w)
tax = Taxonomy(ctx, parents, ranks, names)
params = IndexParams()
def batches():
    per = 16
    for g0 in range(0, w.n_genomes, per):
        g1 = min(w.n_genomes, g0 + per)
        yield oracle.synth_genome(w.gseed, g0 * w.genome_len, (g1 - g0) * w.genome_len), np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len), genome_taxa[g0:g1]
full = KeyValueIndex.build(ctx, tax, params, batches(), expected_bases=w.total_bases)
reads = oracle.synth_reads(w.gseed, w.rseed, w.n_genomes, w.genome_len, 0, w.n_reads, w.read_len)
off = np.arange(w.n_reads + 1, dtype=np.uint64) * np.uint64(w.read_len)
cls = Classifier(full)
a = cls.classify(reads, off, confidence=0.15, per_read_output=True)
b = cls.classify(reads, off, confidence=0.15, per_read_output=False)
sh = ShardedKeyValueIndex(full, 0, 1)
sc = ShardedClassifier(sh)
c = sc.classify(reads, off, confidence=0.15, per_read_output=False)
id1, tx = full.records()
olib = oracle.Library(oracle.params(), parents, len(id1)); olib.add_records(id1, tx)
res, _, _, _ = olib.classify(reads, off.astype(np.int64), confidence=0.15, with_hits=False)
for name, g in (("fused+hits", a), ("fused", b), ("split", c)):
    bad = np.nonzero(res["taxon"] != g.taxon)[0]
    badf = np.nonzero((res["classified"].astype(bool) != g.classified) | (res["has_span"].astype(bool) != g.has_span))[0]
    print(name, "taxon mismatches", len(bad), "flag mismatches", len(badf), "first", bad[:5], [(int(res["taxon"][i]), int(g.taxon[i])) for i in bad[:5]])
