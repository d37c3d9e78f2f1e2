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


# ---------------------------------------------------------------------------------------------- sharded library
class EmuSplitOps:
    """Stand-in for slacken_b200.sharded.GpuSplitOps on CPU tensors: the same per-thread kernel bodies (slk_core.h),
    run by tests/host_emulation, so that ShardedClassifier's routing and exchanges can be driven over gloo."""

    def __init__(self, sp, parents, shard_id1, shard_taxon, taxa_union, k=35, world=1):
        import ctypes as C
        import torch
        from tests import host_emulation as emu
        self.C, self.torch, self.emu, self.sp, self.k = C, torch, emu, sp, k
        self.shard = emu.EmuIndex(sp, parents, shard_id1, shard_taxon, world=world)   # this rank's table
        self.dt = emu.DenseTax(parents, taxa_union)                              # the query side's taxonomy
        self.parent, self.depth, self.raw = self.dt.arrays()

    def upload(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a).copy())

    def scan_spans(self, b1, o1, b2, o2, n):
        emu, p = self.emu, self.emu._p
        nb1, no1 = b1.numpy(), o1.numpy().view(np.uint64)
        nb2 = b2.numpy() if b2 is not None else None
        no2 = o2.numpy().view(np.uint64) if o2 is not None else None
        span_off = np.zeros(n + 1, dtype=np.uint64)
        total = emu.lib().emu_scan_spans(self.C.byref(self.sp), p(nb1), p(no1), p(nb2), p(no2), n, p(span_off), None, 0)
        spans = np.zeros(max(total, 1), dtype=np.uint64)
        assert emu.lib().emu_scan_spans(self.C.byref(self.sp), p(nb1), p(no1), p(nb2), p(no2), n, p(span_off), p(spans), total) == total
        return self.torch.from_numpy(span_off.view(np.int64)), self.torch.from_numpy(spans.view(np.int64)), int(total)

    def route(self, spans, n_spans, world):
        w = spans.numpy().view(np.uint64)[:n_spans]
        seq = np.nonzero(((w >> np.uint64(14)) & np.uint64(3)) == 0)[0]
        dest = np.array([self.emu.lib().emu_shard_of(int(x) >> 16, world) for x in w[seq]], dtype=np.int64)
        order = np.argsort(dest, kind="stable")
        keys = (w[seq][order] >> np.uint64(16)).astype(np.uint64).view(np.int64)
        idx = seq[order].astype(np.int32)
        return self.torch.from_numpy(keys.copy()), self.torch.from_numpy(idx.copy()), np.bincount(dest, minlength=world).tolist()

    def probe(self, keys):
        k = np.ascontiguousarray(keys.numpy()).view(np.uint64)
        taxa = np.zeros(max(len(k), 1), dtype=np.int32)
        p = self.emu._p
        self.emu.lib().emu_probe_keys(p(self.shard.cells), self.shard.n_buckets, p(self.shard.raw), p(k), len(k), p(taxa),
                                      self.shard.world)
        return self.torch.from_numpy(taxa[:len(k)].copy())

    def resolve(self, spans, span_off, n_spans, n_reads, paired, send_idx, taxa, confidence, min_hit_groups, want_hits):
        from slacken_b200.host import DETAIL_DTYPE, HIT_DTYPE, ClassifiedBatch
        emu, p = self.emu, self.emu._p
        dense = np.zeros(max(n_spans, 1), dtype=np.uint16)
        dense[send_idx.numpy()] = [self.dt.to_dense[int(t)] for t in taxa.numpy()]
        res = np.zeros(n_reads, dtype=emu.RESULT_DTYPE)
        hits = np.zeros(max(n_spans, 1), dtype=HIT_DTYPE)
        so = np.ascontiguousarray(span_off.numpy()).view(np.uint64)
        sw = np.ascontiguousarray(spans.numpy()).view(np.uint64)
        emu.lib().emu_resolve_spans(p(self.parent), p(self.depth), p(self.raw), len(self.raw), self.dt.root, self.k, p(sw), p(so),
                                    p(dense), n_reads, float(confidence), int(min_hit_groups), p(res), p(hits))
        detail = np.zeros(n_reads, dtype=DETAIL_DTYPE)
        detail["hit_off"] = so[:-1]
        detail["hit_cnt"] = res["n_hits"]
        detail["num_distinct"] = res["num_distinct"]
        detail["len1"] = res["kmers1"] + (self.k - 1)
        detail["len2"] = res["kmers2"] + (self.k - 1) if paired else 0xFFFFFFFF
        return ClassifiedBatch(res["taxon"].copy(), res["flags"].astype(np.uint8), detail, hits, n_spans)


def _sharded_worker(rank, world, port, out_dir):
    from oracle import oracle
    from slacken_b200.sharded import ShardedClassifier
    from tests import host_emulation as emu
    from tests.util import leaf_taxa, make_taxonomy, random_dna, simulate_reads
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)   # same world on every rank
    parents, _, _ = make_taxonomy(90, 9)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, 3000) for _ in range(8)]
    taxa = np.array([leaves[int(rng.integers(len(leaves)))] for _ in genomes], dtype=np.int32)
    reads = simulate_reads(rng, genomes, 300, (20, 200), n_rate=0.1)
    mates = simulate_reads(rng, genomes, 300, (20, 200), n_rate=0.1)
    lib = oracle.Library(oracle.params(), parents, 1 << 16)
    b, off = oracle.pack_sequences(genomes)
    lib.add_fragments(b, off, taxa)
    id1, tx = lib.records()
    sp = emu.scan_params(35, 31, 7, oracle.DEFAULT_TOGGLE_MASK, True)
    # this rank's shard of the records, by the library's own owner function
    owner = np.array([emu.lib().emu_shard_of(emu.lib().emu_compress(sp, int(k)), world) for k in id1.view(np.uint64)])
    mine = owner == rank
    ops = lambda union: EmuSplitOps(sp, parents, id1[mine], tx[mine], union, world=world)
    shard_taxa = np.unique(tx[mine])
    cls = ShardedClassifier(None, ops=ops, local_taxa=shard_taxa)
    ok = 0 < mine.sum() < len(id1)
    for paired in (False, True):
        lo, hi = shard_bounds(len(reads), rank, world)   # every rank classifies its own reads against ALL shards
        rb, ro = oracle.pack_sequences(reads[lo:hi])
        mb, mo = oracle.pack_sequences(mates[lo:hi]) if paired else (None, None)
        got = cls.classify(rb, ro.astype(np.uint64), mb, mo.astype(np.uint64) if paired else None, confidence=0.1)
        res, _, _, per = lib.classify(rb, ro, mb, mo, confidence=0.1)   # the oracle with the WHOLE library
        ok = ok and np.array_equal(res["taxon"], got.taxon) and np.array_equal(res["classified"].astype(bool), got.classified)
        ok = ok and np.array_equal(res["has_span"].astype(bool), got.has_span)
        for i in range(hi - lo):
            h = got.hits_of(i)
            ok = ok and np.array_equal(h["taxon"], per[i]["taxon"]) and np.array_equal(h["count"], per[i]["count"])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_sharded_library_classify_gloo(tmp_path):
    """configs[4] on CPU: the library split over two ranks by minimizer hash, span keys routed to their owner and taxa
    routed back (point-to-point over gloo), results equal to the oracle that holds the whole library."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"


# ---------------------------------------------------------------------------------------------- distributed build
def _build_worker(rank, world, port, out_dir):
    """The distributed build by hash range (slacken_b200.sharded.ShardedKeyValueIndex.from_builder, C ABI section
    "Distributed build") with the CPU bodies standing in for the kernels: this rank's genomes -> cells in ITS dense
    numbering -> ordered by the table-line hash and LCA-reduced -> already grouped by owner -> exchange() of the 8-byte cells
    and gather of the senders' taxon lists -> the owner renumbers every run and inserts it into its shard's table."""
    import torch
    from oracle import oracle
    from slacken_b200.sharded import exchange
    from tests import host_emulation as emu
    from tests.util import leaf_taxa, make_taxonomy, random_dna
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(91)   # same world on every rank
    parents, _, _ = make_taxonomy(90, 11)
    leaves = leaf_taxa(parents)
    genomes = [random_dna(rng, 3000) for _ in range(6)]
    for i in range(3, 6):             # relatives of the first three: the same minimizers arrive from both ranks
        g = bytearray(genomes[i - 3])
        for j in range(0, len(g), 53):
            g[j] = ord("ACGT"[int(rng.integers(4))])
        genomes[i] = bytes(g)
    g0 = bytearray(genomes[0]); g0[400:430] = b"N" * 30; genomes[0] = bytes(g0)
    taxa = np.array([leaves[int(rng.integers(len(leaves)))] for _ in genomes], dtype=np.int32)
    p = oracle.params()
    sp = emu.scan_params(35, 31, 7, oracle.DEFAULT_TOGGLE_MASK, True)
    L = emu.lib()
    # 1) this rank's genomes -> cells in its own dense numbering
    lo, hi = shard_bounds(len(genomes), rank, world)
    mine = emu.DenseTax(parents, taxa[lo:hi])
    cells = np.concatenate([emu.emit_cells(sp, genomes[i], mine.to_dense[int(taxa[i])]) for i in range(lo, hi)])
    # 2) ordered by the line hash, LCA-reduced, grouped by owner (slk_build_reduce)
    def lca(a, b):
        while a != b:
            if mine.depth[a] >= mine.depth[b]:
                a = mine.parent[a]
            else:
                b = mine.parent[b]
        return a if a else mine.root
    mix = np.array([L.emu_key_mix(int(c) >> 16) for c in cells], dtype=np.int64)
    order = np.lexsort((cells >> np.uint64(16), mix))
    cells, mix = cells[order], mix[order]
    red = []
    for c in cells:
        if red and (red[-1] >> 16) == (int(c) >> 16):
            red[-1] = (red[-1] & ~0xffff) | lca(red[-1] & 0xffff, int(c) & 0xffff)
        else:
            red.append(int(c))
    red = np.array(red, dtype=np.uint64)
    owner = np.array([L.emu_shard_of(int(c) >> 16, world) for c in red], dtype=np.int64)
    ok = bool((np.diff(owner) >= 0).all())                # ordered by the line hash = grouped by owner
    counts = np.bincount(owner, minlength=world).tolist()
    # 3) the cells to their owners, the senders' dense -> raw lists to everybody
    recv, run_cells = exchange(torch.from_numpy(red.view(np.int64).copy()), counts)
    lists = [None] * world
    dist.all_gather_object(lists, [int(x) for x in mine.raw])
    # 4) the owner: union of the lists in rank order, one translation table per run, insert
    shard = emu.EmuIndex(sp, parents, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.int32), world=world)
    shard.dt = emu.DenseTax(parents, [])
    for lst in lists:
        for t in lst:
            if t:
                shard.dt.add(t)
    shard.parent, shard.depth, shard.raw = shard.dt.arrays()
    shard.n_buckets = int(L.emu_buckets_for(int(sum(run_cells))))
    shard.cells = np.zeros(shard.n_buckets * 4, dtype=np.uint64)
    got = recv.numpy().view(np.uint64)
    at = 0
    for lst, n in zip(lists, run_cells):
        run = got[at:at + n]
        table = np.array([shard.dt.to_dense[t] for t in lst], dtype=np.uint64)
        shard.insert((run & ~np.uint64(0xffff)) | table[(run & np.uint64(0xffff)).astype(np.int64)])
        at += n
    # the shard's records == the records of the whole library that this rank owns
    lib = oracle.Library(p, parents, 1 << 16)
    pieces, labels = oracle.remove_invalid(genomes, taxa)
    b, off = oracle.pack_sequences(pieces)
    lib.add_fragments(b, off, labels)
    id1, tx = lib.records()
    own = np.array([L.emu_shard_of(L.emu_compress(sp, int(k)), world) for k in id1.view(np.uint64)]) == rank
    sid, stx = shard.records()
    ok = ok and 0 < own.sum() < len(id1) and np.array_equal(sid, id1.view(np.uint64)[own]) and np.array_equal(stx, tx[own])
    # and the shard answers lookups through the shard-aware line function
    keys = np.array([L.emu_compress(sp, int(k)) for k in id1.view(np.uint64)[own][:200]], dtype=np.uint64)
    ans = np.zeros(len(keys), dtype=np.int32)
    L.emu_probe_keys(emu._p(shard.cells), shard.n_buckets, emu._p(shard.raw), emu._p(keys), len(keys), emu._p(ans), world)
    ok = ok and np.array_equal(ans, tx[own][:200])
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("1" if ok else "0")
    dist.destroy_process_group()


def test_distributed_build_by_hash_range_gloo(tmp_path):
    """configs[2] on CPU, world size 2: reduced cells travel to the rank that owns their range of the table-line hash and
    are merged there by LCA; every shard ends up with exactly the whole library's records it owns."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_build_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok0").read() == "1" and open(tmp_path / "ok1").read() == "1"
