#!/usr/bin/env python
"""Phase timings of the NVLink-mailbox exchange with one process driving every rank (ranks on the GPUs named by --devices;
several ranks may share a device): all ranks scan, then route, then probe, then resolve, with a device synchronisation
between the phases so that each phase can be timed on its own. Not a benchmark of the overlapped path.

    python tools/mailbox_timing.py --devices 0,1 --reads 4000000 --genomes 200
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_workload as bw  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", default="0,0")
    ap.add_argument("--reads", type=int, default=2_000_000)
    ap.add_argument("--genomes", type=int, default=200)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    from slacken_b200 import GpuContext, IndexParams, KeyValueIndex, Taxonomy
    from slacken_b200._lib import check
    from slacken_b200.sharded import Mailbox, ShardedClassifier, ShardedKeyValueIndex, shard_of_records
    devices = [int(x) for x in args.devices.split(",")]
    world = len(devices)
    w = bw.Workload()
    w.n_genomes, w.n_reads = args.genomes, args.reads
    parents, ranks, names, genome_taxa = bw.taxonomy(w)
    params = IndexParams(k=w.k, m=w.m, spaces=w.spaces)
    ctxs = {d: GpuContext(d) for d in sorted(set(devices))}
    taxs = {d: Taxonomy(c, parents, ranks, names) for d, c in ctxs.items()}
    ctx0 = ctxs[devices[0]]

    def synth(fn, n_bytes, *a):
        d = ctx0.dev_alloc(n_bytes)
        check(fn(ctx0.h, *a, C.c_void_p(d)))
        out = np.zeros(n_bytes, dtype=np.uint8)
        ctx0.d2h(out, d)
        ctx0.dev_free(d)
        return out

    def batches():
        per = max(1, (64 << 20) // w.genome_len)
        for g0 in range(0, w.n_genomes, per):
            g1 = min(w.n_genomes, g0 + per)
            bases = synth(ctx0._L.slk_synth_genome_dev, (g1 - g0) * w.genome_len, w.gseed, g0 * w.genome_len, (g1 - g0) * w.genome_len)
            yield bases, (np.arange(g1 - g0 + 1, dtype=np.uint64) * np.uint64(w.genome_len)), genome_taxa[g0:g1]
    full = KeyValueIndex.build(ctx0, taxs[devices[0]], params, batches(), expected_bases=w.total_bases)
    id1, tx = full.records(sort=False)
    full.close()
    owner = shard_of_records(params, id1, world)
    union = np.unique(tx)
    cap = int(args.reads * 44 / world * 1.25) + 65536
    cls = []
    for r, d in enumerate(devices):
        sh = ShardedKeyValueIndex(KeyValueIndex.from_records(ctxs[d], taxs[d], params, id1[owner == r], tx[owner == r]), r, world)
        cls.append(ShardedClassifier(sh, mailbox=Mailbox(ctxs[d], r, world, cap, connect=False), taxa_union=union))
    Mailbox.connect_local([c.mailbox for c in cls])
    n, L = args.reads, w.read_len
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(L)
    dev_reads = []
    for r in range(world):
        reads = synth(ctx0._L.slk_synth_reads_dev, n * L, w.gseed, w.rseed, w.n_genomes, w.genome_len, r * n, n, L)
        dev_reads.append((cls[r].ops.upload(reads), cls[r].ops.upload(off.view(np.int64))))

    def sync_all():
        for c in ctxs.values():
            c.sync()

    for step in range(args.steps):
        t = {}
        t0 = time.perf_counter()
        st = [cls[r].ops.scan_spans(dev_reads[r][0], dev_reads[r][1], None, None, n) for r in range(world)]
        sync_all(); t["scan"] = time.perf_counter() - t0; t0 = time.perf_counter()
        for r in range(world):
            cls[r].mailbox_route(st[r][1], st[r][2])
        sync_all(); t["route"] = time.perf_counter() - t0; t0 = time.perf_counter()
        for r in range(world):
            cls[r].mailbox_probe()
        sync_all(); t["probe"] = time.perf_counter() - t0; t0 = time.perf_counter()
        for r in range(world):
            cls[r].mailbox_resolve(st[r][1], st[r][0], st[r][2], n, False, 0.15, 2, False)
        sync_all(); t["resolve+download"] = time.perf_counter() - t0
        print(f"step {step}: world {world} devices {devices} spans/rank {st[0][2]}: " +
              "  ".join(f"{k} {1e3 * v:.2f} ms" for k, v in t.items()), flush=True)
    for c in cls:
        c.close()


if __name__ == "__main__":
    main()
