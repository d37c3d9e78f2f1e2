"""A library sharded over several GPUs by minimizer hash range (BASELINE.json configs[4]).

Slacken joins spans and records with a Spark shuffle (slacken/Classifier.scala:77-96); here every rank keeps the
records whose minimizer it owns in its own HBM table, classifies its own reads, and the join is two all-to-alls over
NVLink per batch: 8-byte span keys to their owners, 4-byte taxa back. One process per GPU, `torch.distributed` for
the exchange (NCCL on the GPU box; the CPU tests drive the same code over gloo with a stand-in for the device ops).

    index = ShardedKeyValueIndex.from_records(ctx, tax, params, id1, taxon)      # keeps this rank's shard
    cls = ShardedClassifier(index)                                               # collective: all ranks
    batch = cls.classify(bases, off, confidence=0.15)                            # collective: all ranks, own reads each
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import ClassifyOpts, check
from .host import DETAIL_DTYPE, HIT_DTYPE, ClassifiedBatch, GpuContext, IndexParams, KeyValueIndex, Taxonomy


def _dist():
    import torch.distributed as dist
    return dist


def world_of(group=None) -> Tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_of_records(params: IndexParams, id1: np.ndarray, world: int) -> np.ndarray:
    """Owner rank of every record (id1 = the Parquet column: the left-aligned minimizer priority)."""
    id1 = np.ascontiguousarray(id1).view(np.int64)
    out = np.zeros(len(id1), dtype=np.uint8)
    p = params.c_params()
    check(_lib.load().slk_shard_of_records(C.byref(p), id1.ctypes.data_as(C.c_void_p), len(id1), world,
                                           out.ctypes.data_as(C.c_void_p)))
    return out


# ------------------------------------------------------------------------------------------------ the exchange
def exchange(send, send_counts: Sequence[int], group=None):
    """All-to-all of variable-sized slices: rank r receives send[offsets[r] : offsets[r] + send_counts[r]] of every rank,
    concatenated in rank order. Returns (received tensor, receive counts). NCCL: one all_to_all_single; gloo (CPU tests)
    has no all-to-all, so the same exchange is written as point-to-point sends."""
    import torch
    dist = _dist()
    rank, world = world_of(group)
    if world == 1:
        return send, list(send_counts)
    cnt = torch.tensor(list(send_counts), dtype=torch.int64, device=send.device)
    rcnt = torch.empty_like(cnt)
    backend = dist.get_backend(group)
    if backend == "nccl":
        dist.all_to_all_single(rcnt, cnt, group=group)
        recv_counts = [int(x) for x in rcnt.tolist()]
        recv = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv, send, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
        # the library's kernels run on the library's own stream: the received data must be complete before they start
        torch.cuda.current_stream(send.device).synchronize()
        return recv, recv_counts
    gathered = [torch.empty_like(cnt) for _ in range(world)]
    dist.all_gather(gathered, cnt, group=group)
    recv_counts = [int(g[rank]) for g in gathered]
    recv = torch.empty(sum(recv_counts), dtype=send.dtype, device=send.device)
    so = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    ro = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if send_counts[peer]:
            ops.append(dist.P2POp(dist.isend, send[so[peer]:so[peer + 1]].contiguous(), peer, group=group))
        if recv_counts[peer]:
            ops.append(dist.P2POp(dist.irecv, recv[ro[peer]:ro[peer + 1]], peer, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv, recv_counts


def union_of_taxa(local: np.ndarray, group=None) -> np.ndarray:
    """Union (sorted, unique) of the taxa all shards can answer with."""
    import torch
    dist = _dist()
    rank, world = world_of(group)
    local = np.unique(np.asarray(local, dtype=np.int32))
    if world == 1:
        return local
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    n = torch.tensor([len(local)], dtype=torch.int64, device=dev)
    ns = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    cap = max(int(x) for x in ns)
    buf = torch.zeros(max(cap, 1), dtype=torch.int32, device=dev)
    buf[:len(local)] = torch.from_numpy(local).to(dev)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    parts = [b[:int(k)].cpu().numpy() for b, k in zip(bufs, ns)]
    return np.unique(np.concatenate(parts)).astype(np.int32)


# ------------------------------------------------------------------------------------------------ device ops
class GpuSplitOps:
    """The four device steps of the split path through the C ABI, on torch CUDA tensors (torch owns the buffers so that
    NCCL can send them; the kernels are the library's)."""

    def __init__(self, index: KeyValueIndex, taxa_union: np.ndarray):
        import torch
        self.torch = torch
        self.index, self.ctx, self.params = index, index.ctx, index.params
        self.device = torch.device("cuda", self.ctx.device)
        self._L = self.ctx._L
        self._p = self.params.c_params()
        taxa_union = np.ascontiguousarray(taxa_union, dtype=np.int32)
        h = C.c_void_p()
        check(self._L.slk_resolver_create(self.ctx.h, index.taxonomy.h, C.byref(self._p), taxa_union.ctypes.data_as(C.c_void_p),
                                          len(taxa_union), C.byref(h)))
        self.resolver = h

    def close(self):
        if getattr(self, "resolver", None):
            self._L.slk_resolver_destroy(self.resolver)
            self.resolver = None
        for buf in getattr(self, "_pinned", {}).values():
            self.ctx.free_pinned(buf)
        self._pinned = {}

    @staticmethod
    def _v(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None

    def pinned_out(self, name: str, dtype, n: int) -> np.ndarray:
        """The pinned host array of one result kind (grow-only), as n elements of dtype."""
        dtype = np.dtype(dtype)
        nbytes = n * dtype.itemsize
        if not hasattr(self, "_pinned"):
            self._pinned = {}
        buf = self._pinned.get(name)
        if buf is None or buf.nbytes < nbytes:
            if buf is not None:
                self.ctx.free_pinned(buf)
            buf = self.ctx.pinned((max(nbytes + nbytes // 4, 4096),), np.uint8)
            self._pinned[name] = buf
        return buf[:nbytes].view(dtype)

    def download(self, name: str, t, dtype, n: int) -> np.ndarray:
        """Device tensor -> pinned host array (grow-only, one per result kind; pageable copies run at a few GB/s). The
        returned view is valid until the next download of the same kind."""
        out = self.pinned_out(name, dtype, n)
        if out.nbytes:
            self.ctx.d2h(out.view(np.uint8), t.data_ptr())
        return out

    def _results(self, taxon, flags, detail, hits, n_reads: int, n_spans: int, want_hits: bool) -> ClassifiedBatch:
        """The arrays of the batch live in pinned buffers that the next batch overwrites: copy what must outlive it."""
        d = self.download("detail", detail, DETAIL_DTYPE, n_reads) if want_hits else None
        h = self.download("hits", hits, HIT_DTYPE, n_spans) if want_hits else None
        return ClassifiedBatch(self.download("taxon", taxon, np.int32, n_reads), self.download("flags", flags, np.uint8, n_reads),
                               d, h, n_spans if want_hits else 0)

    def upload(self, a: np.ndarray):
        t = self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
        self.torch.cuda.current_stream(self.device).synchronize()   # torch's stream is not the library's
        return t

    def scan_spans(self, bases1, off1, bases2, off2, n_reads: int):
        t = self.torch
        span_off = t.empty(n_reads + 1, dtype=t.int64, device=self.device)   # the library zeroes it itself
        n = C.c_uint64(0)
        args = (self._v(bases1), self._v(off1), self._v(bases2) if bases2 is not None else None,
                self._v(off2) if off2 is not None else None, n_reads, self._v(span_off))
        check(self._L.slk_scan_spans_dev(self.ctx.h, C.byref(self._p), *args, None, 0, C.byref(n)))
        spans = t.empty(max(int(n.value), 1), dtype=t.int64, device=self.device)
        check(self._L.slk_emit_spans_dev(self.ctx.h, C.byref(self._p), *args, self._v(spans)))
        return span_off, spans, int(n.value)

    def route(self, spans, n_spans: int, world: int):
        t = self.torch
        counts = (C.c_uint64 * world)()
        check(self._L.slk_route_spans_dev(self.ctx.h, self._v(spans), n_spans, world, None, None, 0, counts))
        total = sum(counts)
        keys = t.empty(max(total, 1), dtype=t.int64, device=self.device)
        idx = t.empty(max(total, 1), dtype=t.int32, device=self.device)
        check(self._L.slk_route_spans_dev(self.ctx.h, self._v(spans), n_spans, world, self._v(keys), self._v(idx), total, counts))
        return keys[:total], idx[:total], [int(c) for c in counts]

    def probe(self, keys):
        t = self.torch
        taxa = t.empty(max(keys.numel(), 1), dtype=t.int32, device=self.device)
        check(self._L.slk_probe_keys_dev(self.index.h, self._v(keys), keys.numel(), self._v(taxa)))
        return taxa[:keys.numel()]

    def resolve(self, spans, span_off, n_spans: int, n_reads: int, paired: bool, send_idx, taxa, confidence: float,
                min_hit_groups: int, want_hits: bool) -> ClassifiedBatch:
        t = self.torch
        taxon = t.empty(max(n_reads, 1), dtype=t.int32, device=self.device)
        flags = t.empty(max(n_reads, 1), dtype=t.uint8, device=self.device)
        # detail and hits only when the per-read output is wanted (they are 24 B per read + 8 B per hit to bring back)
        detail = t.empty(max(n_reads, 1) * DETAIL_DTYPE.itemsize, dtype=t.uint8, device=self.device) if want_hits else None
        hits = t.empty(max(n_spans, 1) * HIT_DTYPE.itemsize, dtype=t.uint8, device=self.device) if want_hits else None
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        check(self._L.slk_resolve_spans_dev(self.resolver, C.byref(opts), self._v(spans), self._v(span_off), n_spans, n_reads,
                                            1 if paired else 0, self._v(send_idx), self._v(taxa), send_idx.numel(),
                                            self._v(taxon), self._v(flags), self._v(detail), self._v(hits)))
        return self._results(taxon, flags, detail, hits, n_reads, n_spans, want_hits)


# ------------------------------------------------------------------------------------------------ NVLink mailbox
class Mailbox:
    """The two exchanges of the split path as stores into peer memory (include/slacken_gpu.h, "NVLink mailbox"): keys go
    straight from the routing kernel into their owner's inbox, taxa straight from the owner's lookup kernel into the
    asker's reply area, completion flags are awaited on the device. torch.distributed is only used once, to hand the
    CUDA IPC handles around. cap = room for the keys one rank sends to one owner per batch."""

    def __init__(self, ctx: GpuContext, rank: int, world: int, cap: int, group=None, connect: bool = True):
        self.ctx, self.rank, self.world, self.cap = ctx, rank, world, int(cap)
        self._L = ctx._L
        h = C.c_void_p()
        handle = (C.c_uint8 * _lib.IPC_HANDLE_BYTES)()
        check(self._L.slk_mailbox_create(ctx.h, rank, world, self.cap, C.byref(h), handle))
        self.h = h
        self.handle = bytes(handle)
        self._group, self._collective = group, False
        if connect:
            self.connect(group)

    def connect(self, group=None):
        """Collective: every rank's process learns every other rank's mailbox."""
        import torch
        dist = _dist()
        if self.world == 1:
            check(self._L.slk_mailbox_connect(self.h, self.handle))
            return
        dev = torch.device("cuda", self.ctx.device)
        mine = torch.frombuffer(bytearray(self.handle), dtype=torch.uint8).to(dev)
        every = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(every, mine, group=group)
        blob = b"".join(bytes(t.cpu().numpy().tobytes()) for t in every)
        check(self._L.slk_mailbox_connect(self.h, blob))
        dist.barrier(group=group)
        self._group, self._collective = group, True

    @staticmethod
    def connect_local(boxes: Sequence["Mailbox"]):
        """One process driving all ranks (tests): connects by pointer."""
        arr = (C.c_void_p * len(boxes))(*[b.h for b in boxes])
        check(boxes[0]._L.slk_mailbox_connect_local(arr, len(boxes)))

    def close(self):
        """Collective when the mailbox was connected across processes: nobody frees memory a peer has mapped before every
        rank is through with its last batch (the device-side protocol already guarantees that no store is in flight)."""
        if getattr(self, "h", None):
            if self._collective and self.world > 1:
                dist = _dist()
                if dist.is_initialized():
                    dist.barrier(group=self._group)
            self._L.slk_mailbox_destroy(self.h)
            self.h = None


# ------------------------------------------------------------------------------------------------ index + classifier
class ShardedKeyValueIndex:
    """This rank's shard of a KeyValueIndex: the records whose minimizer hashes into the rank's range."""

    def __init__(self, index: KeyValueIndex, rank: int, world: int):
        self.index, self.rank, self.world = index, rank, world
        self.ctx, self.taxonomy, self.params = index.ctx, index.taxonomy, index.params

    @classmethod
    def from_records(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, id1: np.ndarray, taxon: np.ndarray,
                     group=None, rank: Optional[int] = None, world: Optional[int] = None):
        """Every rank is handed the same (or any superset of its share of the) records and keeps its own shard
        (KeyValueIndex.loadRecords, slacken/KeyValueIndex.scala:150-159, per shard)."""
        r, w = world_of(group)
        rank = r if rank is None else rank
        world = w if world is None else world
        mine = shard_of_records(params, id1, world) == rank
        return cls(KeyValueIndex.from_records(ctx, taxonomy, params, np.ascontiguousarray(id1)[mine],
                                              np.ascontiguousarray(taxon)[mine], world=world), rank, world)

    @classmethod
    def build(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, local_batches, expected_bases: int = 0,
              group=None):
        """Distributed KeyValueIndex.makeRecords (slacken/KeyValueIndex.scala:85-122): every rank scans, sorts and
        LCA-reduces ITS genomes on its GPU, the reduced records travel to the owner of their key (one all-to-all;
        LCA is associative and commutative, slacken/LowestCommonAncestor.scala:152-170), and the owner's insert merges
        records of the same minimizer by LCA again."""
        from .host import LibraryBuilder
        b = LibraryBuilder(ctx, taxonomy, params, expected_bases)
        try:
            for bases, off, taxa in local_batches:
                b.add(bases, off, taxa)
            return cls.from_builder(b, group)
        finally:
            b.close()

    @classmethod
    def from_builder(cls, b, group=None):
        """The exchange half of the distributed build, from a LibraryBuilder that has been fed this rank's genomes (it is
        consumed). The owner of a minimizer is a RANGE of the hash that orders the table's lines, and the builder's
        reduced cells are ordered by that hash: they are already grouped by owner, travel as they are (8 bytes each, one
        all-to-all), and every owner inserts the `world` ordered runs it receives front to back."""
        import torch
        ctx, taxonomy, params = b.ctx, b.taxonomy, b.params
        rank, world = world_of(group)
        if world == 1:
            index = b.finish()
            cls.last_build_counts = [len(index)]
            return cls(index, 0, 1)
        if _dist().get_backend(group) != "nccl":   # CPU tests (gloo): through a local table and host memory
            return cls.from_local(b.finish(), group)
        import time
        tm, t0 = {}, time.perf_counter()

        def lap(name):
            nonlocal t0
            torch.cuda.synchronize()
            now = time.perf_counter()
            tm[name] = now - t0
            t0 = now
        dev = torch.device("cuda", ctx.device)
        counts = b.reduce(world)
        cls.last_build_counts = counts
        lap("sort_and_lca_reduce")
        send = b.cells_tensor()     # the builder's own buffer: nothing is copied
        raw = b.dense_taxa()
        # the senders' dense -> raw lists (at most 65 535 ids each): padded all-gather
        dist = _dist()
        n_raw = torch.tensor([len(raw)], dtype=torch.int64, device=dev)
        all_n = [torch.empty_like(n_raw) for _ in range(world)]
        dist.all_gather(all_n, n_raw, group=group)
        run_dense = [int(x) for x in torch.cat(all_n).tolist()]
        pad = torch.zeros(max(run_dense), dtype=torch.int32, device=dev)
        pad[:len(raw)] = torch.from_numpy(raw).to(dev)
        all_raw = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(all_raw, pad, group=group)
        dense_raw = np.concatenate([t[:n].cpu().numpy() for t, n in zip(all_raw, run_dense)])
        recv, run_cells = exchange(send, counts, group)
        del send
        lap("all_to_all")
        n_recv = int(recv.numel())
        if torch.cuda.mem_get_info(dev)[0] < 24 * n_recv + (4 << 30):
            b.close()               # the table needs the room of the builder's buffers; otherwise their release can wait
        out = cls(KeyValueIndex.from_cell_runs(ctx, taxonomy, params, world, recv.data_ptr(), run_cells, dense_raw, run_dense),
                  rank, world)
        del recv
        torch.cuda.empty_cache()   # the staging tensors must not keep HBM that the classifier's buffers will want
        lap("insert_on_owner")
        b.close()
        cls.last_build_times = tm
        return out

    @classmethod
    def from_local(cls, local: KeyValueIndex, group=None):
        """The exchange half of the distributed build: `local` holds the LCA-reduced records of this rank's genomes (it is
        consumed); the result holds the records this rank owns, merged over all ranks."""
        import torch
        ctx, taxonomy, params = local.ctx, local.taxonomy, local.params
        rank, world = world_of(group)
        if world == 1:
            return cls(local, 0, 1)
        if _dist().get_backend(group) != "nccl":   # CPU tests (gloo): the records travel through host memory
            id1, taxon = local.records(sort=False)
            local.close()
            owner = shard_of_records(params, id1, world)
            order = np.argsort(owner, kind="stable")
            counts = np.bincount(owner, minlength=world).tolist()
            rid, _ = exchange(torch.from_numpy(id1.view(np.int64)[order]), counts, group)
            rtx, _ = exchange(torch.from_numpy(taxon[order]), counts, group)
            return cls(KeyValueIndex.from_records(ctx, taxonomy, params, rid.numpy(), rtx.numpy(), world=world), rank, world)
        # the records never leave HBM: dump the local table, group the rows by owner, all-to-all, insert on the owner
        import time
        tm, t0 = {}, time.perf_counter()

        def lap(name):
            nonlocal t0
            torch.cuda.synchronize()
            now = time.perf_counter()
            tm[name] = now - t0
            t0 = now
        n = len(local)
        dev = torch.device("cuda", ctx.device)
        sid = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        stx = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        cnt = (C.c_uint64 * world)()
        check(ctx._L.slk_index_records_by_owner_dev(local.h, world, C.c_void_p(sid.data_ptr()), C.c_void_p(stx.data_ptr()), n, cnt))
        counts = [int(c) for c in cnt]
        sid, stx = sid[:n], stx[:n]
        local.close()
        lap("dump_local_table_by_owner")
        rid, _ = exchange(sid, counts, group)
        rtx, _ = exchange(stx, counts, group)
        del sid, stx
        lap("all_to_all")
        out = cls(KeyValueIndex.from_records_dev(ctx, taxonomy, params, rid, rtx, world=world), rank, world)
        del rid, rtx
        torch.cuda.empty_cache()   # the staging tensors must not keep HBM that the classifier's buffers will want
        lap("insert_on_owner")
        cls.last_build_times = tm
        return out

    last_build_times: dict = {}
    last_build_counts: list = []

    def taxa(self) -> np.ndarray:
        n = C.c_uint32(0)
        check(self.ctx._L.slk_index_taxa(self.index.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        check(self.ctx._L.slk_index_taxa(self.index.h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return out

    def __len__(self) -> int:
        return len(self.index)

    def close(self):
        self.index.close()


class ShardedClassifier:
    """Classifier over a sharded library. Construction and classify() are collective: every rank of the group calls them
    (with its own reads, possibly none)."""

    def __init__(self, shard: Optional[ShardedKeyValueIndex], group=None, ops=None, local_taxa: Optional[np.ndarray] = None,
                 mailbox_cap: int = 0, mailbox: Optional[Mailbox] = None, taxa_union: Optional[np.ndarray] = None):
        """mailbox_cap > 0 (or a ready-made mailbox): the exchanges go through the NVLink mailbox instead of NCCL;
        mailbox_cap = room for the keys this rank sends to ONE owner per batch (about spans per batch / world, plus slack)."""
        self.group = group
        # a ready-made mailbox means the caller drives the ranks itself (one process, several ranks: the tests): the
        # shard knows its place; otherwise the process group does
        self.rank, self.world = (mailbox.rank, mailbox.world) if mailbox is not None else world_of(group)
        taxa = taxa_union if taxa_union is not None else union_of_taxa(shard.taxa() if shard is not None else local_taxa, group)
        self.ops = ops(taxa) if ops is not None else GpuSplitOps(shard.index, taxa)
        self.mailbox = mailbox
        if mailbox is None and mailbox_cap > 0:
            self.mailbox = Mailbox(shard.ctx, self.rank, self.world, mailbox_cap, group)
        self.last_exchange_bytes = (0, 0)
        self.last_times: dict = {}

    def classify(self, bases1: np.ndarray, off1: np.ndarray, bases2: Optional[np.ndarray] = None,
                 off2: Optional[np.ndarray] = None, confidence: float = 0.0, min_hit_groups: int = 2,
                 per_read_output: bool = True) -> ClassifiedBatch:
        """Host arrays in, host arrays out (the reads are uploaded first)."""
        import time
        ops = self.ops
        paired = bases2 is not None
        t0 = time.perf_counter()
        d_b1, d_o1 = ops.upload(bases1 if len(bases1) else np.zeros(16, np.uint8)), ops.upload(off1.astype(np.uint64).view(np.int64))
        d_b2 = ops.upload(bases2 if len(bases2) else np.zeros(16, np.uint8)) if paired else None
        d_o2 = ops.upload(off2.astype(np.uint64).view(np.int64)) if paired else None
        t_up = time.perf_counter() - t0
        out = self.classify_uploaded(d_b1, d_o1, d_b2, d_o2, len(off1) - 1, confidence, min_hit_groups, per_read_output)
        self.last_times["upload"] = t_up
        return out

    def classify_uploaded(self, d_b1, d_o1, d_b2, d_o2, n: int, confidence: float = 0.0, min_hit_groups: int = 2,
                          per_read_output: bool = True) -> ClassifiedBatch:
        """The same for reads already on the device (tensors made by ops.upload): scan, route, exchange, probe, exchange,
        resolve. `last_times` holds the wall time of every step of the last call, `last_exchange_bytes` what went over
        the wire from this rank."""
        import time
        ops = self.ops
        paired = d_b2 is not None
        tm = {}
        t = time.perf_counter()

        def lap(name):
            nonlocal t
            now = time.perf_counter()
            tm[name] = now - t
            t = now
        span_off, spans, n_spans = ops.scan_spans(d_b1, d_o1, d_b2, d_o2, n)
        lap("scan")
        if self.mailbox is not None:
            check(ops._L.slk_mailbox_set_blocks_per_sm(self.mailbox.h, 8))
            self.mailbox_route(spans, n_spans)
            self.mailbox_probe()
            out = self.mailbox_resolve(spans, span_off, n_spans, n, paired, confidence, min_hit_groups, per_read_output)
            lap("route_probe_resolve_and_download")
            self.last_times = tm
            return out
        keys, idx, counts = ops.route(spans, n_spans, self.world)
        lap("route")
        recv_keys, recv_counts = exchange(keys, counts, self.group)          # keys -> owners
        lap("keys_all_to_all")
        taxa_here = ops.probe(recv_keys)
        lap("probe")
        taxa, _ = exchange(taxa_here, recv_counts, self.group)                 # taxa -> askers, in send order
        lap("taxa_all_to_all")
        self.last_exchange_bytes = (8 * int(keys.numel()), 4 * int(taxa.numel()))
        out = ops.resolve(spans, span_off, n_spans, n, paired, idx, taxa, confidence, min_hit_groups, per_read_output)
        lap("resolve_and_download")
        self.last_times = tm
        return out

    def classify_pipelined(self, batches, confidence: float = 0.0, min_hit_groups: int = 2, per_read_output: bool = True):
        """Generator over the results of a sequence of uploaded batches (tuples d_b1, d_o1, d_b2, d_o2, n), mailbox only.
        The span scan of batch e+1 runs (on the library's scan stream) while the mailbox kernels of batch e are in flight:
        the scan is bound by the integer pipe, the exchange by HBM and NVLink. Collective: every rank passes the same
        number of batches. A result is valid until the next one is produced (pinned buffers)."""
        assert self.mailbox is not None, "the pipelined path needs the NVLink mailbox"
        ops = self.ops
        check(ops._L.slk_mailbox_set_blocks_per_sm(self.mailbox.h, 4))   # leave half of every SM to the overlapping scan
        import os
        import time
        trace = [] if os.environ.get("SLK_TRACE_SHARDED") else None   # host-side call boundaries of every step (ms since the first)
        t0 = time.perf_counter()

        def mark(name):
            if trace is not None:
                trace.append((name, round((time.perf_counter() - t0) * 1e3, 3)))
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        scanned = ops.scan_spans(*cur)
        self.mailbox_route(scanned[1], scanned[2])
        self.mailbox_probe()
        while cur is not None:
            span_off, spans, n_spans = scanned
            mark("scan_next>")
            nxt = next(it, None)
            scanned_next = ops.scan_spans(*nxt) if nxt is not None else None   # overlaps with the exchange of `cur`
            mark("resolve>")
            pending = self._resolve_async(spans, span_off, n_spans, cur[4], cur[2] is not None, confidence, min_hit_groups,
                                          per_read_output)
            if nxt is not None:             # queued behind the resolve kernels: the GPU goes on while the results leave
                self.mailbox_route(scanned_next[1], scanned_next[2])
                self.mailbox_probe()
            mark("wait>")
            out = self._resolve_wait(pending)
            pending = None                  # the batch's device tensors go back to the allocator before the next scan
            mark("done")
            self.last_trace = trace
            yield out
            cur, scanned = nxt, scanned_next

    def _resolve_async(self, spans, span_off, n_spans: int, n_reads: int, paired: bool, confidence: float, min_hit_groups: int,
                       want_hits: bool):
        ops, t = self.ops, self.ops.torch
        taxon = t.empty(max(n_reads, 1), dtype=t.int32, device=ops.device)
        flags = t.empty(max(n_reads, 1), dtype=t.uint8, device=ops.device)
        detail = t.empty(max(n_reads, 1) * DETAIL_DTYPE.itemsize, dtype=t.uint8, device=ops.device) if want_hits else None
        hits = t.empty(max(n_spans, 1) * HIT_DTYPE.itemsize, dtype=t.uint8, device=ops.device) if want_hits else None
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        check(ops._L.slk_mailbox_resolve_async(self.mailbox.h, ops.resolver, C.byref(opts), ops._v(spans), ops._v(span_off), n_spans,
                                               n_reads, 1 if paired else 0, ops._v(taxon), ops._v(flags), ops._v(detail),
                                               ops._v(hits)))
        return (spans, span_off, taxon, flags, detail, hits, n_reads, n_spans, want_hits)

    def _resolve_wait(self, pending) -> ClassifiedBatch:
        spans, span_off, taxon, flags, detail, hits, n_reads, n_spans, want_hits = pending
        ops = self.ops
        h_taxon = ops.pinned_out("taxon", np.int32, n_reads)
        h_flags = ops.pinned_out("flags", np.uint8, n_reads)
        check(ops._L.slk_mailbox_resolve_wait(self.mailbox.h, ops.resolver, n_reads, ops._v(taxon), ops._v(flags),
                                              h_taxon.ctypes.data_as(C.c_void_p), h_flags.ctypes.data_as(C.c_void_p)))
        d = ops.download("detail", detail, DETAIL_DTYPE, n_reads) if want_hits else None
        h = ops.download("hits", hits, HIT_DTYPE, n_spans) if want_hits else None
        return ClassifiedBatch(h_taxon, h_flags, d, h, n_spans if want_hits else 0)

    # the three mailbox steps, separately (a single process driving several ranks interleaves them: tests)
    def mailbox_route(self, spans, n_spans: int):
        check(self.ops._L.slk_mailbox_route(self.mailbox.h, self.ops._v(spans), n_spans))

    def mailbox_probe(self):
        check(self.ops._L.slk_mailbox_probe(self.mailbox.h, self.ops.index.h))

    def mailbox_resolve(self, spans, span_off, n_spans: int, n_reads: int, paired: bool, confidence: float, min_hit_groups: int,
                        want_hits: bool) -> ClassifiedBatch:
        ops, t = self.ops, self.ops.torch
        taxon = t.empty(max(n_reads, 1), dtype=t.int32, device=ops.device)
        flags = t.empty(max(n_reads, 1), dtype=t.uint8, device=ops.device)
        detail = t.empty(max(n_reads, 1) * DETAIL_DTYPE.itemsize, dtype=t.uint8, device=ops.device) if want_hits else None
        hits = t.empty(max(n_spans, 1) * HIT_DTYPE.itemsize, dtype=t.uint8, device=ops.device) if want_hits else None
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        check(ops._L.slk_mailbox_resolve(self.mailbox.h, ops.resolver, C.byref(opts), ops._v(spans), ops._v(span_off), n_spans,
                                         n_reads, 1 if paired else 0, ops._v(taxon), ops._v(flags), ops._v(detail), ops._v(hits)))
        return ops._results(taxon, flags, detail, hits, n_reads, n_spans, want_hits)

    def close(self):
        if self.mailbox is not None:
            self.mailbox.close()
            self.mailbox = None
        if hasattr(self.ops, "close"):
            self.ops.close()
