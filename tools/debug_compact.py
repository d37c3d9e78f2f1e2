import sys, numpy as np, ctypes as C
sys.path.insert(0, '.')
import bench_workload as bw
from slacken_b200 import Classifier, GpuContext, IndexParams, Taxonomy, KeyValueIndex
from slacken_b200._lib import check
from slacken_b200.host import pack_reads, compact_reads, block_offsets
import bench
w = bw.Workload(); w.n_genomes = 20; w.genome_len = 1_000_000
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
ctx = GpuContext(0)
parents, ranks, names, genome_taxa = bw.taxonomy(w)
tax = Taxonomy(ctx, parents, ranks, names)
params = IndexParams()
index, _ = bench.build_gpu_library(ctx, tax, params, w, genome_taxa)
L = 150
d = ctx.dev_alloc(n * L)
check(ctx._L.slk_synth_reads_dev(ctx.h, w.gseed, w.rseed, w.n_genomes, w.genome_len, 0, n, L, C.c_void_p(d)))
reads = np.zeros(n * L, dtype=np.uint8); ctx.d2h(reads, d)
off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
p1 = pack_reads(reads, off)
c1 = compact_reads(p1)
cls = Classifier(index)
a = cls.classify_packed(p1, confidence=0.0)
b = cls.classify_compact(c1, None, thresholds=[0.0])
print('taxon', np.array_equal(a.taxon, b.taxon), 'flags', np.array_equal(a.flags & 3, b.flags), 'len1', np.array_equal(a.detail['len1'], b.results['len1']),
      'cnt', np.array_equal(a.detail['hit_cnt'], b.hit_cnt), 'used', a.hits_used, b.hits_used, int(b.hit_cnt.sum()))
bad = np.nonzero(a.taxon != b.taxon)[0]
print('first bad', bad[:10], len(bad))
cnt = b.hit_cnt.astype(np.int64)
within = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
gi = np.repeat(a.detail['hit_off'].astype(np.int64), cnt) + within
eq = (b.hits[:b.hits_used]['taxon'] == a.hits[gi]['taxon']) & (b.hits[:b.hits_used]['count'] == a.hits[gi]['count'])
print('hits equal', eq.all(), 'first bad hit idx', np.nonzero(~eq)[0][:5])
