"""Timing of BrackenWeights.build on synthetic genomes (one GPU)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import bench_workload as bw
from slacken_b200 import GpuContext, IndexParams, KeyValueIndex, LibraryBuilder, Taxonomy
from slacken_b200._lib import check
from slacken_b200.bracken import BrackenWeights

n_genomes = int(sys.argv[1]) if len(sys.argv) > 1 else 50
w = bw.Workload(); w.n_genomes = n_genomes
ctx = GpuContext(0)
parents, ranks, names, genome_taxa = bw.taxonomy(w)
tax = Taxonomy(ctx, parents, ranks, names)
genomes = []
d = ctx.dev_alloc(w.genome_len)
for g in range(n_genomes):
    check(ctx._L.slk_synth_genome_dev(ctx.h, w.gseed, g * w.genome_len, w.genome_len, C.c_void_p(d)))
    b = np.zeros(w.genome_len, dtype=np.uint8); ctx.d2h(b, d); genomes.append(b.tobytes())
def batches():
    for g in range(n_genomes):
        yield np.frombuffer(genomes[g], dtype=np.uint8), np.array([0, w.genome_len], dtype=np.uint64), genome_taxa[g:g + 1]
index = KeyValueIndex.build(ctx, tax, IndexParams(), batches(), expected_bases=w.total_bases)
t0 = time.perf_counter()
weights = BrackenWeights(index, 100).build(list(zip(genome_taxa.tolist(), genomes)))
dt = time.perf_counter() - t0
print(f"{n_genomes} genomes, {w.total_bases/1e6:.0f} Mbp, {sum(weights.values())} reads of 100 bp classified in {dt:.2f} s "
      f"({w.total_bases/dt/1e6:.0f} Mbases/s), {len(weights)} (dest, source) pairs")
