"""Do an ALU-bound scan kernel and a DRAM-bound probe kernel overlap when launched from two streams?
Two host threads, one GpuContext (= one stream) each, synchronous C calls (ctypes releases the GIL)."""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
import bench_workload as bw
from slacken_b200 import GpuContext, IndexParams, KeyValueIndex, Taxonomy
from slacken_b200._lib import check
from slacken_b200.sharded import GpuSplitOps

w = bw.Workload(); w.n_genomes = 1000; n = 4_000_000
ctx1, ctx2 = GpuContext(0), GpuContext(0)
parents, ranks, names, genome_taxa = bw.taxonomy(w)
tax1, tax2 = Taxonomy(ctx1, parents, ranks, names), Taxonomy(ctx2, parents, ranks, names)
params = IndexParams()
def build(ctx, tax):
    from slacken_b200 import LibraryBuilder
    per = 64
    d_bases, d_off, d_tax = ctx.dev_alloc(per * w.genome_len), ctx.dev_alloc((per + 1) * 8), ctx.dev_alloc(per * 4)
    b = LibraryBuilder(ctx, tax, params, expected_bases=w.total_bases)
    for g0 in range(0, w.n_genomes, per):
        g1 = min(w.n_genomes, g0 + per); nb = (g1 - g0) * w.genome_len
        check(ctx._L.slk_synth_genome_dev(ctx.h, w.gseed, g0 * w.genome_len, nb, C.c_void_p(d_bases)))
        ctx.h2d(d_off, np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len)); ctx.h2d(d_tax, genome_taxa[g0:g1])
        b.add_dev(d_bases, d_off, d_tax, g1 - g0, nb)
    ix = b.finish(); b.close(); return ix
index = build(ctx2, tax2)
opsA = GpuSplitOps(KeyValueIndex(ctx1, tax1, params, index.h), np.unique(genome_taxa))   # scan side on ctx1's stream
opsA.ctx = ctx1
opsB = GpuSplitOps(index, np.unique(genome_taxa))                                          # probe side on ctx2's stream
d = ctx1.dev_alloc(n * 150)
check(ctx1._L.slk_synth_reads_dev(ctx1.h, w.gseed, w.rseed, w.n_genomes, w.genome_len, 0, n, 150, C.c_void_p(d)))
reads = np.zeros(n * 150, dtype=np.uint8); ctx1.d2h(reads, d)
d_reads, d_off = opsA.upload(reads), opsA.upload((np.arange(n + 1, dtype=np.uint64) * np.uint64(150)).view(np.int64))
span_off, spans, n_spans = opsA.scan_spans(d_reads, d_off, None, None, n)
keys, idx, counts = opsA.route(spans, n_spans, 1)
print("spans", n_spans, "keys", keys.numel())
def scan_loop(k, out):
    t0 = time.perf_counter()
    for _ in range(k): opsA.scan_spans(d_reads, d_off, None, None, n)
    out["scan"] = (time.perf_counter() - t0) / k
def probe_loop(k, out):
    t0 = time.perf_counter()
    for _ in range(k): opsB.probe(keys)
    out["probe"] = (time.perf_counter() - t0) / k
o = {}
scan_loop(3, o); probe_loop(3, o)
scan_loop(10, o); probe_loop(10, o)
print("alone   : scan %.2f ms, probe %.2f ms" % (1e3 * o["scan"], 1e3 * o["probe"]))
o2 = {}
ta, tb = threading.Thread(target=scan_loop, args=(10, o2)), threading.Thread(target=probe_loop, args=(17, o2))
t0 = time.perf_counter(); ta.start(); tb.start(); ta.join(); tb.join(); wall = time.perf_counter() - t0
print("together: scan %.2f ms, probe %.2f ms per call; wall %.1f ms for 10 scans + 17 probes (serial would be %.1f ms)" %
      (1e3 * o2["scan"], 1e3 * o2["probe"], 1e3 * wall, 1e3 * (10 * o["scan"] + 17 * o["probe"])))
