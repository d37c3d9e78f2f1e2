#!/usr/bin/env python
"""bench_sharded.py -- BASELINE.json configs[4]: a library sharded over the GPUs of one box by minimizer hash range,
query minimizers routed to their owners by NCCL all-to-all (slacken_b200/sharded.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench_sharded.py --gpus N [--reads R] [--genomes G] [--steps K] [--warmup W] [--check]

Every rank builds the records of ITS genomes on its GPU, the records travel to the owner of their minimizer (the
distributed build of SURVEY.md section 8e), and every rank classifies its own reads against all shards. Not the
driver's bench (that is bench.py, the replicated-library configuration); prints one JSON line of its own.
--check compares rank 0's results with the fused kernel on a replicated copy of the library (small workloads only).
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--reads", type=int, default=2_000_000, help="reads per GPU per step")
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--genome-len", type=int, default=4_000_000)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--mailbox", action="store_true",
                    help="exchange keys and taxa through the NVLink mailbox (peer-memory stores from the kernels) instead of NCCL")
    ap.add_argument("--pipeline", action="store_true",
                    help="with --mailbox: scan batch e+1 while the exchange of batch e is in flight (classify_pipelined)")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from slacken_b200 import Classifier, GpuContext, IndexParams, KeyValueIndex, Taxonomy
    from slacken_b200.dist import shard_bounds
    from slacken_b200.sharded import ShardedClassifier, ShardedKeyValueIndex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = bw.Workload()
    w.n_genomes, w.genome_len, w.n_reads = args.genomes, args.genome_len, args.reads
    ctx = GpuContext(local)
    parents, ranks, names, genome_taxa = bw.taxonomy(w)
    tax = Taxonomy(ctx, parents, ranks, names)
    params = IndexParams(k=w.k, m=w.m, spaces=w.spaces)
    import ctypes as C
    from slacken_b200._lib import check

    def synth_genome(start: int, n_bases: int) -> np.ndarray:   # the library's device generator, copied to the host
        d = ctx.dev_alloc(n_bases)
        check(ctx._L.slk_synth_genome_dev(ctx.h, w.gseed, start, n_bases, C.c_void_p(d)))
        out = np.zeros(n_bases, dtype=np.uint8)
        ctx.d2h(out, d)
        ctx.dev_free(d)
        return out

    def synth_reads(first: int, n_reads: int) -> np.ndarray:
        d = ctx.dev_alloc(n_reads * w.read_len)
        check(ctx._L.slk_synth_reads_dev(ctx.h, w.gseed, w.rseed, w.n_genomes, w.genome_len, first, n_reads, w.read_len, C.c_void_p(d)))
        out = np.zeros(n_reads * w.read_len, dtype=np.uint8)
        ctx.d2h(out, d)
        ctx.dev_free(d)
        return out

    # distributed build: this rank's genomes only
    g_lo, g_hi = shard_bounds(w.n_genomes, rank, world)

    def batches():
        per = max(1, (64 << 20) // w.genome_len)
        for g0 in range(g_lo, g_hi, per):
            g1 = min(g_hi, g0 + per)
            bases = synth_genome(g0 * w.genome_len, (g1 - g0) * w.genome_len)
            yield bases, (np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len)), genome_taxa[g0:g1]

    t0 = time.perf_counter()
    shard = ShardedKeyValueIndex.build(ctx, tax, params, batches(), expected_bases=(g_hi - g_lo) * w.genome_len)
    t_build = time.perf_counter() - t0
    n_local = len(shard)
    tot = torch.tensor([n_local], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    # mailbox capacity per (asker, owner) pair: ~40 spans per 150-base read spread evenly over the owners, plus 25 % slack
    cap = int(args.reads * 44 / world * 1.25) + 65536
    cls = ShardedClassifier(shard, mailbox_cap=cap if args.mailbox else 0)

    n, L = w.n_reads, w.read_len
    reads = synth_reads(rank * n, n)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)

    d_reads, d_off = cls.ops.upload(reads), cls.ops.upload(off.view(np.int64))   # resident in HBM, like bench.py's `value`

    def step():
        return cls.classify_uploaded(d_reads, d_off, None, None, n, confidence=0.15, min_hit_groups=w.min_hit_groups,
                                     per_read_output=False)

    def run(k):
        if not (args.pipeline and args.mailbox):
            for _ in range(k):
                out = step()
            return out
        out = None
        for out in cls.classify_pipelined([(d_reads, d_off, None, None, n)] * k, confidence=0.15, min_hit_groups=w.min_hit_groups,
                                          per_read_output=False):
            pass
        return out

    got = run(args.warmup)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = run(args.steps)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    if world > 1:
        dist.barrier()
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())

    same = None
    if args.check and rank == 0:
        # the whole library on one GPU, fused kernel: must give the same taxa
        def all_batches():
            per = max(1, (64 << 20) // w.genome_len)
            for g0 in range(0, w.n_genomes, per):
                g1 = min(w.n_genomes, g0 + per)
                bases = synth_genome(g0 * w.genome_len, (g1 - g0) * w.genome_len)
                yield bases, (np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len)), genome_taxa[g0:g1]
        full = KeyValueIndex.build(ctx, tax, params, all_batches(), expected_bases=w.total_bases)
        ref = Classifier(full).classify(reads, off, confidence=0.15, min_hit_groups=w.min_hit_groups, per_read_output=False)
        same = bool(np.array_equal(ref.taxon, got.taxon) and np.array_equal(ref.flags, got.flags) and len(full) == int(tot.item()))
    kb, tb = cls.last_exchange_bytes
    if rank == 0:
        print(json.dumps({
            "metric": "reads/sec classified (150bp), library sharded by minimizer hash range", "value": world * n * args.steps / wall,
            "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
            "exchange": "NVLink mailbox (peer-memory stores fused into the route and lookup kernels)" if args.mailbox else "NCCL all-to-all",
            "pipelined": bool(args.pipeline and args.mailbox),
            "timing": "host wall clock around the collective classify_uploaded() calls (reads resident in HBM: scan, route, two "
                      "exchanges, probe, resolve, D2H of taxon and flags), max over ranks",
            "step_breakdown_rank0_s": cls.last_times,
            "config": {"workload": f"synthetic {n} x {L}bp reads per GPU vs {w.total_bases/1e9:.2f} Gbp library in {world} shards",
                       "library_records": int(tot.item()), "records_on_rank0": n_local, "confidence": 0.15},
            "exchange_bytes_per_step_rank0": {"keys_out": kb, "taxa_back": tb},
            "build": {"seconds": t_build, "what": "local scan + sort + LCA reduce, all-to-all of the reduced records, insert on the owner"},
            "equal_to_replicated_fused_kernel": same}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
