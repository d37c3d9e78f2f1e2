#!/usr/bin/env python
"""bench_build.py -- BASELINE.json configs[2]: library build (minimizer scan + sort + LCA reduce + hash-table construction)
from a synthetic genome set over the synthetic 50k-node taxonomy, on 1/2/4/8 GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_build.py --gpus N [--gbp-per-gpu 8.75] [--genome-len 4000000]

Weak scaling: every rank scans, sorts and LCA-reduces 8.75 Gbp of genomes (70 Gbp on 8 GPUs, the size of the standard
library) generated in its own HBM, then the reduced records travel to the owner of their minimizer in one all-to-all and
the owner's insert merges equal minimizers by LCA (slacken_b200/sharded.py, ShardedKeyValueIndex.from_builder). The result
is the library sharded by minimizer hash range, ready for bench_sharded.py's classifier. Prints one JSON line. Not the
driver's bench (that is bench.py)."""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_workload as bw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gbp-per-gpu", type=float, default=8.75)
    ap.add_argument("--genome-len", type=int, default=4_000_000)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from slacken_b200 import GpuContext, IndexParams, LibraryBuilder, Taxonomy
    from slacken_b200._lib import check
    from slacken_b200.dist import shard_bounds
    from slacken_b200.sharded import ShardedKeyValueIndex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = bw.Workload()
    w.genome_len = args.genome_len
    w.n_genomes = max(world, int(round(args.gbp_per_gpu * 1e9 * world / w.genome_len)))
    ctx = GpuContext(local)
    parents, ranks, names, genome_taxa = bw.taxonomy(w)
    tax = Taxonomy(ctx, parents, ranks, names)
    params = IndexParams(k=w.k, m=w.m, spaces=w.spaces)
    g_lo, g_hi = shard_bounds(w.n_genomes, rank, world)
    per = max(1, min(g_hi - g_lo, (256 << 20) // w.genome_len))
    d_bases = ctx.dev_alloc(per * w.genome_len)
    d_off = ctx.dev_alloc((per + 1) * 8)
    d_tax = ctx.dev_alloc(per * 4)
    if world > 1:   # NCCL sets its communicators up lazily: not part of the build
        warm = torch.zeros(world * 4, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(torch.empty_like(warm), warm)
        dist.barrier()
    ctx.sync()
    t0 = time.perf_counter()
    b = LibraryBuilder(ctx, tax, params, expected_bases=(g_hi - g_lo) * w.genome_len)
    for g0 in range(g_lo, g_hi, per):
        g1 = min(g_hi, g0 + per)
        n = (g1 - g0) * w.genome_len
        check(ctx._L.slk_synth_genome_dev(ctx.h, w.gseed, g0 * w.genome_len, n, C.c_void_p(d_bases)))
        ctx.h2d(d_off, np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len))
        ctx.h2d(d_tax, genome_taxa[g0:g1])
        b.add_dev(d_bases, d_off, d_tax, g1 - g0, n)
    for p in (d_bases, d_off, d_tax):
        ctx.dev_free(p)
    ctx.sync()
    t_local = time.perf_counter() - t0
    shard = ShardedKeyValueIndex.from_builder(b)
    b.close()
    ctx.sync()
    torch.cuda.synchronize()
    t_all = time.perf_counter() - t0
    n_local = int(sum(ShardedKeyValueIndex.last_build_counts))
    t = torch.tensor([t_local, t_all], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([n_local, len(shard)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    t_local, t_all = float(t[0]), float(t[1])
    if rank == 0:
        print(json.dumps({
            "metric": "library build throughput (scan + sort + LCA reduce + hash-table construction)",
            "value": w.total_bases / t_all / 1e9, "unit": "Gbases/s", "n_gpus": world, "seconds": t_all, "scaling": "weak",
            "timing": "host wall clock from the first genome batch to the finished sharded table, genome generation on the "
                      "device included, max over ranks",
            "phases_s": {"local scan (minimizers of this rank's genomes -> cells)": t_local,
                         "sort + LCA reduce, cells to their owners (all-to-all), insert on the owner": t_all - t_local,
                         "exchange_breakdown_rank0": ShardedKeyValueIndex.last_build_times},
            "config": {"workload": f"{w.n_genomes} synthetic genomes x {w.genome_len} bp = {w.total_bases / 1e9:.2f} Gbp, "
                                   f"{len(parents)}-node taxonomy, k{w.k}/m{w.m}/s{w.spaces}",
                       "records_before_exchange": int(cnt[0]), "library_records": int(cnt[1]), "records_on_rank0": len(shard)}}),
            flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
