"""World-size-2 test (gloo, CPU) of the host-side multi-GPU logic: reads are sharded over ranks, every rank counts
the taxa of its shard (here with the CPU oracle standing in for the device), and one all-reduce gives the report
counts of the whole job."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from slacken_b200.dist import shard_bounds


def test_shard_bounds_partition():
    for n in (0, 1, 7, 10, 1000003):
        for world in (1, 2, 3, 8):
            parts = [shard_bounds(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    import torch
    from oracle import oracle
    from slacken_b200.dist import allreduce_counts, max_over_ranks
    from tests.util import leaf_taxa, make_taxonomy, random_dna, simulate_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)   # same world on every rank
    parents, _, _ = make_taxonomy(80, 5)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, 2000) for _ in range(6)]
    taxa = np.array([leaves[int(rng.integers(len(leaves)))] for _ in genomes], dtype=np.int32)
    reads = simulate_reads(rng, genomes, 401, (30, 120))
    lib = oracle.Library(oracle.params(), parents, 1 << 16)   # replicated library
    b, off = oracle.pack_sequences(genomes)
    lib.add_fragments(b, off, taxa)
    lo, hi = shard_bounds(len(reads), rank, world)
    rb, ro = oracle.pack_sequences(reads[lo:hi])
    res, _, _, _ = lib.classify(rb, ro, confidence=0.05, with_hits=False, threads=1)
    counts = torch.from_numpy(np.bincount(res["taxon"][res["has_span"].astype(bool)], minlength=len(parents)).astype(np.int64))
    allreduce_counts(counts)
    slowest = max_over_ranks(float(rank + 1))
    if rank == 0:
        rb, ro = oracle.pack_sequences(reads)
        full, _, _, _ = lib.classify(rb, ro, confidence=0.05, with_hits=False, threads=1)
        want = np.bincount(full["taxon"][full["has_span"].astype(bool)], minlength=len(parents))
        ok = np.array_equal(counts.numpy(), want) and slowest == float(world)
        open(os.path.join(out_dir, "ok"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_counts_allreduce_gloo(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
