"""Device timeline of the compact entry point's chunk pipeline (SLK_TRACE=1): python tools/trace_compact.py [n_reads]"""
import os, sys, numpy as np, ctypes as C
os.environ["SLK_TRACE"] = "1"
sys.path.insert(0, '.')
import bench_workload as bw, bench
from slacken_b200 import Classifier, GpuContext, IndexParams, Taxonomy
from slacken_b200._lib import check
from slacken_b200.host import pack_reads, CompactReads, CompactBatch, RESULT_DTYPE, HIT_DTYPE, block_offsets, PackedReads
w = bw.Workload(); w.n_genomes = int(os.environ.get("GENOMES", "1000"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = GpuContext(0)
parents, ranks, names, genome_taxa = bw.taxonomy(w)
tax = Taxonomy(ctx, parents, ranks, names)
index, _ = bench.build_gpu_library(ctx, tax, IndexParams(), w, genome_taxa)
L = 150
d = ctx.dev_alloc(n * L)
check(ctx._L.slk_synth_reads_dev(ctx.h, w.gseed, w.rseed, w.n_genomes, w.genome_len, 0, n, L, C.c_void_p(d)))
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
boff = block_offsets(off); nb = int(boff[-1])
d_off, d_boff = ctx.dev_alloc(off.nbytes), ctx.dev_alloc(boff.nbytes); ctx.h2d(d_off, off); ctx.h2d(d_boff, boff)
codes, mask, ln = ctx.dev_alloc(nb * 8), ctx.dev_alloc(nb * 4), ctx.dev_alloc(n * 4)
ctx.pack_reads_dev(d, d_off, n, d_boff, codes, mask, ln)
hc, hl = ctx.pinned(nb, np.uint64), ctx.pinned(n, np.uint32)
ctx.d2h(hc, codes); ctx.d2h(hl, ln)
e = np.zeros(0, dtype=np.uint32)
cr = CompactReads(hc, hl, e, e)
cls = Classifier(index)
out = CompactBatch(ctx.pinned(n, RESULT_DTYPE), np.zeros((0, n), np.int32), np.zeros((0, n), np.uint8), ctx.pinned(16 * n, HIT_DTYPE))
import time
for i in range(3):
    t0 = time.perf_counter(); cls.classify_compact(cr, None, thresholds=[0.0], out=out); print("call ms", 1e3 * (time.perf_counter() - t0), file=sys.stderr)
