"""Host-side mirror of the reference's classes for the hot path (the part of the Scala driver that would call the
C ABI). Citations are relative to /root/reference/src/main/scala/com/jnpersson/."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import ClassifyMultiOpts, ClassifyOpts, Params, check

DEFAULT_TOGGLE_MASK = 0xE37E28C4271B5A2D  # kmers/minimizer/package.scala:32
NONE, ROOT = 0, 1  # slacken/Taxonomy.scala:30-31
AMBIGUOUS_SPAN, MATE_PAIR_BORDER = -1, -2  # slacken/package.scala:28-29

HIT_DTYPE = np.dtype([("taxon", "<i4"), ("count", "<i4")])
DETAIL_DTYPE = np.dtype([("hit_off", "<u8"), ("hit_cnt", "<u4"), ("len1", "<u4"), ("len2", "<u4"), ("num_distinct", "<u4")])
assert DETAIL_DTYPE.itemsize == 24
RESULT_DTYPE = np.dtype([("taxon", "<i4"), ("len1", "<u4"), ("len2", "<u4"), ("hits_flags", "<u4")])   # slk_read_result
RESULT_SHORT_DTYPE = np.dtype([("taxon", "<i4"), ("hits_flags", "<u4")])                                # slk_read_result_short


def _ptr(a) -> Optional[C.c_void_p]:
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


class GpuContext:
    """One CUDA device (slk_ctx). One process per GPU is the deployment model."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        h = C.c_void_p()
        check(self._L.slk_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self._L.slk_ctx_destroy(self.h)
            self.h = None

    def sync(self):
        check(self._L.slk_ctx_sync(self.h))

    # pinned host arrays ------------------------------------------------------------
    def pinned(self, shape, dtype) -> np.ndarray:
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        check(self._L.slk_host_alloc(max(n, 1), C.byref(p)))
        buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=np.uint8, count=n).view(dtype).reshape(shape)
        _PINNED[arr.ctypes.data] = (p, self._L)
        return arr

    def free_pinned(self, arr: np.ndarray):
        ent = _PINNED.pop(arr.ctypes.data, None)
        if ent:
            ent[1].slk_host_free(ent[0])

    # raw device memory (for device-resident benchmarks) ------------------------------
    def dev_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self._L.slk_dev_alloc(self.h, nbytes, C.byref(p)))
        return p.value

    def dev_free(self, p: int):
        self._L.slk_dev_free(self.h, C.c_void_p(p))

    def h2d(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        check(self._L.slk_memcpy_h2d(self.h, C.c_void_p(dst), _ptr(src), src.nbytes))

    def pack_reads_dev(self, bases_dev: int, off_dev: int, n_reads: int, boff_dev: int, codes_dev: int, mask_dev: int,
                       len_dev: int):
        """Stage 1 on the device: ASCII reads in HBM -> packed blocks in HBM."""
        v = C.c_void_p
        check(self._L.slk_pack_reads_dev(self.h, v(bases_dev), v(off_dev), n_reads, v(boff_dev), v(codes_dev), v(mask_dev),
                                         v(len_dev)))

    def d2h(self, dst: np.ndarray, src: int):
        assert dst.flags["C_CONTIGUOUS"]
        check(self._L.slk_memcpy_d2h(self.h, _ptr(dst), C.c_void_p(src), dst.nbytes))


_PINNED: dict = {}


@dataclass
class IndexParams:
    """kmers/IndexParams.scala:63-91 + kmers/SplitterFormat.scala:55-77 (splitter = randomXOR)."""
    k: int = 35
    m: int = 31
    spaces: int = 7
    canonical: bool = True
    toggle_mask: int = DEFAULT_TOGGLE_MASK
    buckets: int = 200

    def c_params(self) -> Params:
        p = Params()
        check(_lib.load().slk_params_init(self.k, self.m, self.spaces, self.toggle_mask, 1 if self.canonical else 0,
                                          C.byref(p)))
        return p


def significant_mask(params: "IndexParams") -> int:
    """The bits of a minimizer (the left-aligned 64-bit priority word stored in id1) that can be non-zero: the m-mer's bits,
    without the positions a spaced seed blanks out (SpacedSeed.spaceMask, kmers/minimizer/MinimizerPriorities.scala:287-301)."""
    m, spaces = params.m, params.spaces
    full = (1 << 64) - 1
    r = full
    if m % 32 != 0:
        r &= (full << (64 - (m % 32) * 2)) & full
    if spaces > 0:
        final_bits = 3 << ((64 - (m % 32) * 2) & 63)
        for _ in range(spaces):
            r = ((r << 4) & full) | final_bits
    return r


def expand_keys(params: "IndexParams", ckeys: np.ndarray) -> np.ndarray:
    """Compressed keys (the significant bits of a minimizer gathered at the low end, as the table cells and the span words of
    the split path hold them) -> the id1 values of the Parquet column."""
    mask = significant_mask(params)
    ck = np.ascontiguousarray(ckeys).astype(np.uint64)
    out = np.zeros(len(ck), dtype=np.uint64)
    j = 0
    for b in range(64):
        if (mask >> b) & 1:
            out |= ((ck >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
            j += 1
    return out


class Taxonomy:
    """slacken/Taxonomy.scala:159-160: parents[] indexed by raw taxon id (NONE = 0, ROOT = 1), plus the rank titles
    and scientific names the report needs (host only)."""

    def __init__(self, ctx: GpuContext, parents: np.ndarray, ranks: Optional[Sequence[Optional[str]]] = None,
                 names: Optional[Sequence[Optional[str]]] = None):
        self.ctx = ctx
        self.parents = np.ascontiguousarray(parents, dtype=np.int32).copy()
        self.parents[ROOT] = NONE  # Taxonomy.fromNodesAndNames, slacken/Taxonomy.scala:105
        self.ranks = list(ranks) if ranks is not None else [None] * len(self.parents)
        self.names = list(names) if names is not None else [None] * len(self.parents)
        h = C.c_void_p()
        check(ctx._L.slk_taxonomy_create(ctx.h, _ptr(self.parents), len(self.parents), C.byref(h)))
        self.h = h

    @property
    def size(self) -> int:
        return len(self.parents)

    def is_defined(self, t: int) -> bool:  # slacken/Taxonomy.scala:175-176
        return bool(self.parents[t] != NONE or t == ROOT)

    def close(self):
        if getattr(self, "h", None):
            self.ctx._L.slk_taxonomy_destroy(self.h)
            self.h = None


class KeyValueIndex:
    """slacken/KeyValueIndex.scala: the minimizer -> LCA taxon library, resident in HBM as an open-addressing table."""

    def __init__(self, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, handle):
        self.ctx, self.taxonomy, self.params, self.h = ctx, taxonomy, params, handle

    # KeyValueIndex.loadRecords (slacken/KeyValueIndex.scala:150-159)
    @classmethod
    def from_records(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, id1: np.ndarray, taxon: np.ndarray,
                     world: int = 1):
        """world > 1: the records are one shard of a library cut for `world` GPUs (sharded.shard_of_records); the table then
        spreads the shard's range of the line hash over all of its lines."""
        id1 = np.ascontiguousarray(id1).view(np.int64)
        taxon = np.ascontiguousarray(taxon, dtype=np.int32)
        assert len(id1) == len(taxon)
        p = params.c_params()
        h = C.c_void_p()
        check(ctx._L.slk_index_from_records_shard(ctx.h, taxonomy.h, C.byref(p), _ptr(id1), _ptr(taxon), len(id1), int(world),
                                                  C.byref(h)))
        return cls(ctx, taxonomy, params, h)

    # KeyValueIndex.makeRecords (slacken/KeyValueIndex.scala:85-122): genomes -> records, on the GPU
    @classmethod
    def build(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, batches, expected_bases: int = 0):
        """batches: iterable of (bases uint8[], frag_off uint64[n+1], frag_taxon int32[n])."""
        b = LibraryBuilder(ctx, taxonomy, params, expected_bases)
        try:
            for bases, off, taxa in batches:
                b.add(bases, off, taxa)
            return b.finish()
        finally:
            b.close()

    def __len__(self) -> int:
        return int(self.ctx._L.slk_index_size(self.h))

    def taxa(self) -> np.ndarray:
        """The raw taxon ids this index can answer with, ancestors included (slk_index_taxa); position i holds the taxon that
        the 4-byte hit format calls label i + 1."""
        n = C.c_uint32()
        check(self.ctx._L.slk_index_taxa(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        check(self.ctx._L.slk_index_taxa(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def records(self, sort: bool = True):
        """(id1 uint64[], taxon int32[]) -- the rows KeyValueIndex.writeRecords stores; sorted by id1 on request
        (the table itself has no order)."""
        n = len(self)
        id1 = np.zeros(n, dtype=np.int64)
        taxon = np.zeros(n, dtype=np.int32)
        got = C.c_uint64()
        check(self.ctx._L.slk_index_records(self.h, _ptr(id1), _ptr(taxon), n, C.byref(got)))
        if not sort:
            return id1.view(np.uint64), taxon
        o = np.argsort(id1.view(np.uint64), kind="stable")
        return id1.view(np.uint64)[o], taxon[o]

    def records_dev(self):
        """The same rows as torch tensors in device memory (id1 int64, taxon int32), unsorted."""
        import torch
        n = len(self)
        dev = torch.device("cuda", self.ctx.device)
        id1 = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        taxon = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
        got = C.c_uint64()
        check(self.ctx._L.slk_index_records(self.h, C.c_void_p(id1.data_ptr()), C.c_void_p(taxon.data_ptr()), n, C.byref(got)))
        return id1[:n], taxon[:n]

    @classmethod
    def from_records_dev(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, id1, taxon, world: int = 1):
        """from_records for torch tensors that already live on the context's device."""
        import torch
        id1, taxon = id1.contiguous(), taxon.contiguous()
        assert id1.dtype == torch.int64 and taxon.dtype == torch.int32 and id1.numel() == taxon.numel()
        torch.cuda.current_stream(id1.device).synchronize()   # torch's stream is not the library's
        p = params.c_params()
        h = C.c_void_p()
        check(ctx._L.slk_index_from_records_shard(ctx.h, taxonomy.h, C.byref(p), C.c_void_p(id1.data_ptr()),
                                                  C.c_void_p(taxon.data_ptr()), id1.numel(), int(world), C.byref(h)))
        return cls(ctx, taxonomy, params, h)

    @classmethod
    def from_cell_runs(cls, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, world: int, cells_dev: int,
                       run_cells: Sequence[int], dense_raw: np.ndarray, run_dense: Sequence[int]):
        """The owner's half of the distributed build: `cells_dev` holds len(run_cells) runs of reduced cells back to back,
        run r written with the dense taxa dense_raw[sum(run_dense[:r]) : +run_dense[r]] (LibraryBuilder.dense_taxa of
        its sender)."""
        n = len(run_cells)
        rc = (C.c_uint64 * max(n, 1))(*[int(c) for c in run_cells])
        rd = (C.c_uint32 * max(n, 1))(*[int(c) for c in run_dense])
        raw = np.ascontiguousarray(dense_raw, dtype=np.int32)
        assert len(raw) == sum(int(c) for c in run_dense)
        p = params.c_params()
        h = C.c_void_p()
        check(ctx._L.slk_index_from_cell_runs(ctx.h, taxonomy.h, C.byref(p), int(world), n, C.c_void_p(cells_dev), rc,
                                              _ptr(raw), rd, C.byref(h)))
        return cls(ctx, taxonomy, params, h)

    def close(self):
        if getattr(self, "h", None):
            self.ctx._L.slk_index_destroy(self.h)
            self.h = None


class LibraryBuilder:
    """Incremental form of KeyValueIndex.makeRecords: feed (taxon, genome fragment) batches, then finish()."""

    def __init__(self, ctx: GpuContext, taxonomy: Taxonomy, params: IndexParams, expected_bases: int = 0):
        self.ctx, self.taxonomy, self.params = ctx, taxonomy, params
        p = params.c_params()
        b = C.c_void_p()
        check(ctx._L.slk_build_begin(ctx.h, taxonomy.h, C.byref(p), int(expected_bases), C.byref(b)))
        self.h = b

    def add(self, bases: np.ndarray, frag_off: np.ndarray, frag_taxon: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        off = np.ascontiguousarray(frag_off, dtype=np.uint64)
        taxa = np.ascontiguousarray(frag_taxon, dtype=np.int32)
        check(self.ctx._L.slk_build_add(self.h, _ptr(bases), _ptr(off), _ptr(taxa), len(taxa)))

    def add_dev(self, bases_dev: int, frag_off_dev: int, frag_taxon_dev: int, n_frag: int, total_bases: int):
        """Fragments already resident in HBM (offsets index bases_dev directly)."""
        check(self.ctx._L.slk_build_add_dev(self.h, C.c_void_p(bases_dev), C.c_void_p(frag_off_dev),
                                            C.c_void_p(frag_taxon_dev), n_frag, total_bases))

    def finish(self) -> "KeyValueIndex":
        h = C.c_void_p()
        check(self.ctx._L.slk_build_finish(self.h, C.byref(h)))
        return KeyValueIndex(self.ctx, self.taxonomy, self.params, h)

    # the sending half of the distributed build (include/slacken_gpu.h, "Distributed build")
    def reduce(self, world: int) -> List[int]:
        """Sort + LCA reduce; returns the number of reduced cells bound for every owner."""
        cnt = (C.c_uint64 * world)()
        check(self.ctx._L.slk_build_reduce(self.h, int(world), cnt))
        return [int(c) for c in cnt]

    def cells_dev(self):
        """(device pointer, count) of the reduced cells, grouped by owner; valid until close()."""
        p, n = C.c_void_p(), C.c_uint64()
        check(self.ctx._L.slk_build_cells_dev(self.h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def cells_tensor(self):
        """The same memory as a torch int64 tensor (no copy; keep the builder open while it is in use)."""
        import torch
        ptr, n = self.cells_dev()
        dev = torch.device("cuda", self.ctx.device)
        if n == 0:
            return torch.empty(0, dtype=torch.int64, device=dev)

        class _View:   # the CUDA array interface torch.as_tensor understands
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2, "strides": None}
        with torch.cuda.device(dev):
            return torch.as_tensor(_View(), device=dev)

    def dense_taxa(self) -> np.ndarray:
        n = C.c_uint32()
        check(self.ctx._L.slk_build_dense_taxa(self.h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.int32)
        check(self.ctx._L.slk_build_dense_taxa(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def close(self):
        if getattr(self, "h", None):
            self.ctx._L.slk_build_destroy(self.h)
            self.h = None


class DeviceTimer:
    """CUDA events on a classifier's launch stream."""

    def __init__(self, classifier: "Classifier"):
        self.c, self.L = classifier, classifier.ctx._L
        self.a, self.b = C.c_void_p(), C.c_void_p()
        check(self.L.slk_event_create(classifier.ctx.h, C.byref(self.a)))
        check(self.L.slk_event_create(classifier.ctx.h, C.byref(self.b)))

    def start(self):
        check(self.L.slk_event_record(self.a, self.c.h))

    def stop(self):
        check(self.L.slk_event_record(self.b, self.c.h))

    def elapsed_ms(self) -> float:
        ms = C.c_float()
        check(self.L.slk_event_elapsed_ms(self.a, self.b, C.byref(ms)))
        return float(ms.value)

    def close(self):
        self.L.slk_event_destroy(self.a)
        self.L.slk_event_destroy(self.b)


@dataclass
class ClassifyParams:
    """slacken/Classifier.scala:47-63"""
    min_hit_groups: int = 2
    with_unclassified: bool = True
    thresholds: List[float] = field(default_factory=lambda: [0.0])
    sample_regex: Optional[str] = None
    per_read_output: bool = True


class ReportCounts:
    """Device-resident per-(sample, taxon) read counters: groupBy(sampleId, taxon).count, slacken/Classifier.scala:214-217."""

    def __init__(self, ctx: GpuContext, taxonomy: Taxonomy, n_samples: int = 1):
        self.ctx, self.taxonomy, self.n_samples = ctx, taxonomy, n_samples
        h = C.c_void_p()
        check(ctx._L.slk_counts_create(ctx.h, taxonomy.h, n_samples, C.byref(h)))
        self.h = h

    def add(self, taxon: np.ndarray, flags: np.ndarray, sample_id: Optional[np.ndarray] = None):
        taxon = np.ascontiguousarray(taxon, dtype=np.int32)
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        sid = np.ascontiguousarray(sample_id, dtype=np.int32) if sample_id is not None else None
        check(self.ctx._L.slk_counts_add(self.h, _ptr(taxon), _ptr(flags), _ptr(sid), len(taxon)))

    def fetch(self, sample: int = 0) -> np.ndarray:
        out = np.zeros(self.taxonomy.size, dtype=np.int64)
        check(self.ctx._L.slk_counts_fetch(self.h, sample, _ptr(out), len(out)))
        return out

    def pairs(self, sample: int = 0):
        """[(taxon, count)] with count > 0, the input of KrakenReport."""
        v = self.fetch(sample)
        nz = np.nonzero(v)[0]
        return [(int(t), int(v[t])) for t in nz]

    def device_ptr(self) -> int:
        return self.ctx._L.slk_counts_device_ptr(self.h)

    def reset(self):
        check(self.ctx._L.slk_counts_reset(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.ctx._L.slk_counts_destroy(self.h)
            self.h = None


@dataclass
class ClassifiedBatch:
    """Arrays behind a batch of ClassifiedRead (slacken/Classifier.scala:27-45)."""
    taxon: np.ndarray
    flags: np.ndarray
    detail: Optional[np.ndarray]
    hits: Optional[np.ndarray]
    hits_used: int = 0

    @property
    def classified(self) -> np.ndarray:
        return (self.flags & _lib.READ_CLASSIFIED) != 0

    @property
    def has_span(self) -> np.ndarray:
        return (self.flags & _lib.READ_HAS_SPAN) != 0

    def hits_of(self, i: int) -> np.ndarray:
        d = self.detail[i]
        return self.hits[int(d["hit_off"]):int(d["hit_off"]) + int(d["hit_cnt"])]


class Classifier:
    """slacken/Classifier.scala: classify batches of reads against a KeyValueIndex."""

    def __init__(self, index: KeyValueIndex):
        self.index, self.ctx = index, index.ctx
        h = C.c_void_p()
        check(self.ctx._L.slk_classifier_create(index.h, C.byref(h)))
        self.h = h

    def attach_counts(self, counts: Optional[ReportCounts], sample: int = 0):
        check(self.ctx._L.slk_classifier_attach_counts(self.h, counts.h if counts else None, sample))

    def hits_bound(self, n_reads: int, total_bases: int, paired: bool) -> int:
        p = self.index.params.c_params()
        return int(self.ctx._L.slk_classify_hits_bound(C.byref(p), n_reads, total_bases, 1 if paired else 0))

    def classify(self, bases1: np.ndarray, off1: np.ndarray, bases2: Optional[np.ndarray] = None,
                 off2: Optional[np.ndarray] = None, confidence: float = 0.0, min_hit_groups: int = 2,
                 per_read_output: bool = True, out: Optional[ClassifiedBatch] = None) -> ClassifiedBatch:
        """Classifier.classify (slacken/Classifier.scala:114-122) for one batch held in host arrays."""
        n = len(off1) - 1
        assert bases1.dtype == np.uint8 and off1.dtype == np.uint64
        paired = bases2 is not None
        if out is None:
            taxon = np.zeros(n, dtype=np.int32)
            flags = np.zeros(n, dtype=np.uint8)
            detail = np.zeros(n, dtype=DETAIL_DTYPE)
            hits = None
            if per_read_output:
                total = int(off1[-1] - off1[0]) + (int(off2[-1] - off2[0]) if paired else 0)
                hits = np.zeros(self.hits_bound(n, total, paired), dtype=HIT_DTYPE)
            out = ClassifiedBatch(taxon, flags, detail, hits)
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        used = C.c_uint64(0)
        hits = out.hits if per_read_output else None
        check(self.ctx._L.slk_classify_batch(self.h, C.byref(opts), _ptr(bases1), _ptr(off1), _ptr(bases2), _ptr(off2), n,
                                             _ptr(out.taxon), _ptr(out.flags), _ptr(out.detail), _ptr(hits),
                                             len(hits) if hits is not None else 0, C.byref(used)))
        out.hits_used = int(used.value)
        return out

    def classify_packed(self, r1: PackedReads, r2: Optional[PackedReads] = None, confidence: float = 0.0,
                        min_hit_groups: int = 2, per_read_output: bool = True,
                        out: Optional[ClassifiedBatch] = None) -> ClassifiedBatch:
        """The packed-input twin of classify(): 2-bit blocks + ambiguity masks in host memory."""
        n = len(r1.len)
        if out is None:
            hits = None
            if per_read_output:
                total = int(r1.len.sum()) + (int(r2.len.sum()) if r2 is not None else 0)
                hits = np.zeros(self.hits_bound(n, total, r2 is not None), dtype=HIT_DTYPE)
            out = ClassifiedBatch(np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.uint8), np.zeros(n, dtype=DETAIL_DTYPE), hits)
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        used = C.c_uint64(0)
        hits = out.hits if per_read_output else None
        m2 = (r2.codes, r2.mask, r2.boff, r2.len) if r2 is not None else (None, None, None, None)
        check(self.ctx._L.slk_classify_batch_packed(self.h, C.byref(opts), _ptr(r1.codes), _ptr(r1.mask), _ptr(r1.boff),
                                                    _ptr(r1.len), _ptr(m2[0]), _ptr(m2[1]), _ptr(m2[2]), _ptr(m2[3]), n,
                                                    _ptr(out.taxon), _ptr(out.flags), _ptr(out.detail), _ptr(hits),
                                                    len(hits) if hits is not None else 0, C.byref(used)))
        out.hits_used = int(used.value)
        return out

    @staticmethod
    def _multi_opts(thresholds, min_hit_groups: int) -> ClassifyMultiOpts:
        thr = [float(t) for t in thresholds]
        o = ClassifyMultiOpts()
        o.n_thresholds, o.min_hit_groups = len(thr), int(min_hit_groups)
        for i, t in enumerate(thr[:_lib.MAX_THRESHOLDS]):
            o.confidence[i] = t
        return o

    def classify_packed_thresholds(self, r1: PackedReads, r2: Optional[PackedReads], thresholds: Sequence[float],
                                   min_hit_groups: int = 2, per_read_output: bool = True):
        """Classifier.classify for several confidence thresholds (slacken/Classifier.scala:156-170): the reads are scanned and
        looked up once. Returns (taxon int32[n_thr, n], flags uint8[n_thr, n], detail, hits, hits_used)."""
        n, nt = len(r1.len), len(thresholds)
        taxon, flags = np.zeros((nt, n), dtype=np.int32), np.zeros((nt, n), dtype=np.uint8)
        detail = np.zeros(n, dtype=DETAIL_DTYPE)
        hits = None
        if per_read_output:
            total = int(r1.len.sum()) + (int(r2.len.sum()) if r2 is not None else 0)
            hits = np.zeros(self.hits_bound(n, total, r2 is not None), dtype=HIT_DTYPE)
        o = self._multi_opts(thresholds, min_hit_groups)
        used = C.c_uint64(0)
        m2 = (r2.codes, r2.mask, r2.boff, r2.len) if r2 is not None else (None, None, None, None)
        check(self.ctx._L.slk_classify_batch_packed_multi(self.h, C.byref(o), _ptr(r1.codes), _ptr(r1.mask), _ptr(r1.boff), _ptr(r1.len),
                                                          _ptr(m2[0]), _ptr(m2[1]), _ptr(m2[2]), _ptr(m2[3]), n, _ptr(taxon), _ptr(flags),
                                                          _ptr(detail), _ptr(hits), len(hits) if hits is not None else 0, C.byref(used)))
        return taxon, flags, detail, hits, int(used.value)

    def classify_compact(self, r1: "CompactReads", r2: Optional["CompactReads"] = None, thresholds: Sequence[float] = (0.0,),
                         min_hit_groups: int = 2, per_read_output: bool = True, out: Optional["CompactBatch"] = None,
                         short_hits: bool = False) -> "CompactBatch":
        """The compact boundary (include/slacken_gpu.h, slk_classify_batch_compact): codes + lengths + a sparse list of
        ambiguous positions in, 16-byte results + hits in read order out. short_hits: 4-byte hits and 8-byte results
        (slk_classify_batch_compact_short; out.hits is then a uint32 array and out.results RESULT_SHORT_DTYPE, see
        CompactBatch.decode_short_hits / lengths_from_hits)."""
        n, nt = len(r1.len), len(thresholds)
        if out is None:
            hits = None
            if per_read_output:
                total = int(r1.len.sum()) + (int(r2.len.sum()) if r2 is not None else 0)
                hits = np.zeros(self.hits_bound(n, total, r2 is not None), dtype=np.uint32 if short_hits else HIT_DTYPE)
            out = CompactBatch(np.zeros(n, dtype=RESULT_SHORT_DTYPE if short_hits else RESULT_DTYPE), np.zeros((max(nt - 1, 0), n), dtype=np.int32),
                               np.zeros((max(nt - 1, 0), n), dtype=np.uint8), hits)
        amb = merge_ambiguous(r1, r2)
        o = self._multi_opts(thresholds, min_hit_groups)
        used = C.c_uint64(0)
        hits = out.hits if per_read_output else None
        assert out.results.dtype == (RESULT_SHORT_DTYPE if short_hits else RESULT_DTYPE)
        if hits is not None:
            assert hits.dtype == (np.uint32 if short_hits else HIT_DTYPE)
        fn = self.ctx._L.slk_classify_batch_compact_short if short_hits else self.ctx._L.slk_classify_batch_compact
        check(fn(self.h, C.byref(o), _ptr(r1.codes), _ptr(r1.len), _ptr(r2.codes) if r2 is not None else None,
                 _ptr(r2.len) if r2 is not None else None, _ptr(amb) if len(amb) else None, len(amb), n,
                 _ptr(out.results), _ptr(out.taxon_more) if nt > 1 else None,
                 _ptr(out.flags_more) if nt > 1 else None, _ptr(hits),
                 len(hits) if hits is not None else 0, C.byref(used)))
        out.hits_used = int(used.value)
        return out

    def classify_packed_dev(self, codes1: int, mask1: int, boff1: int, len1: int, codes2: int, mask2: int, boff2: int,
                            len2: int, n_reads: int, taxon_out: int, flags_out: int, detail_out: int, hits_out: int,
                            hits_cap: int, hits_used_dev: int, confidence: float = 0.0, min_hit_groups: int = 2):
        """Device-resident packed input; every argument is a device address (0 = NULL); asynchronous."""
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        v = lambda p: C.c_void_p(p) if p else None
        check(self.ctx._L.slk_classify_packed_dev(self.h, C.byref(opts), v(codes1), v(mask1), v(boff1), v(len1), v(codes2),
                                                  v(mask2), v(boff2), v(len2), n_reads, v(taxon_out), v(flags_out),
                                                  v(detail_out), v(hits_out), hits_cap, v(hits_used_dev)))

    def classify_dev(self, bases1: int, off1: int, bases2: int, off2: int, n_reads: int, taxon_out: int, flags_out: int,
                     detail_out: int, hits_out: int, hits_cap: int, hits_used_dev: int, confidence: float = 0.0,
                     min_hit_groups: int = 2):
        """Device-resident variant: every argument is a device address (0 = NULL); asynchronous on self.stream."""
        opts = ClassifyOpts(float(confidence), int(min_hit_groups), 0)
        v = lambda p: C.c_void_p(p) if p else None
        check(self.ctx._L.slk_classify_batch_dev(self.h, C.byref(opts), v(bases1), v(off1), v(bases2), v(off2), n_reads,
                                                 v(taxon_out), v(flags_out), v(detail_out), v(hits_out), hits_cap,
                                                 v(hits_used_dev)))

    @property
    def stream(self) -> int:
        return self.ctx._L.slk_classifier_stream(self.h)

    def sync(self):
        check(self.ctx._L.slk_classifier_sync(self.h))

    @property
    def launches(self) -> int:
        return int(self.ctx._L.slk_classifier_launches(self.h))

    def stats(self):
        """(table probes, merged hits) accumulated over every launch of this classifier."""
        a, b = C.c_uint64(), C.c_uint64()
        check(self.ctx._L.slk_classifier_stats(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def close(self):
        if getattr(self, "h", None):
            self.ctx._L.slk_classifier_destroy(self.h)
            self.h = None


_CODE = np.full(256, 4, dtype=np.uint8)
for _ch, _c in ((b"Aa", 0), (b"Cc", 1), (b"Gg", 2), (b"TtUu", 3)):
    for _b in _ch:
        _CODE[_b] = _c


@dataclass
class PackedReads:
    """2-bit packed batch (include/slacken_gpu.h, "Packed input"): what the Scala host would fill while it copies the
    nucleotides of a partition into its pinned buffers."""
    codes: np.ndarray   # uint64[n_blocks]
    mask: np.ndarray    # uint32[n_blocks]
    boff: np.ndarray    # uint64[n+1]
    len: np.ndarray     # uint32[n]

    @property
    def nbytes(self) -> int:
        return self.codes.nbytes + self.mask.nbytes + self.boff.nbytes + self.len.nbytes


@dataclass
class CompactReads:
    """One mate of a batch at the compact boundary: 2-bit code blocks, lengths, and the ambiguous positions as
    (read index, position) pairs sorted by read."""
    codes: np.ndarray     # uint64[n_blocks]
    len: np.ndarray       # uint32[n]
    amb_read: np.ndarray  # uint32[n_amb]
    amb_pos: np.ndarray   # uint32[n_amb]

    @property
    def nbytes(self) -> int:
        return self.codes.nbytes + self.len.nbytes + 8 * len(self.amb_read)


@dataclass
class CompactBatch:
    results: np.ndarray      # RESULT_DTYPE[n]: threshold 0
    taxon_more: np.ndarray   # int32[n_thr - 1, n]
    flags_more: np.ndarray   # uint8[n_thr - 1, n]
    hits: Optional[np.ndarray]
    hits_used: int = 0

    @property
    def taxon(self):
        return self.results["taxon"]

    @property
    def flags(self):
        return (self.results["hits_flags"] & 3).astype(np.uint8)

    @property
    def hit_cnt(self):
        return self.results["hits_flags"] >> 2

    def decode_short_hits(self, index_taxa: np.ndarray, k: int) -> np.ndarray:
        """4-byte hits (slk_classify_batch_compact_short) -> HIT_DTYPE records: index_taxa = KeyValueIndex.taxa() (the list of
        slk_index_taxa), k = the k-mer width (the mate-pair border's count is -(k - 1))."""
        w = self.hits[:self.hits_used].astype(np.uint32)
        label, count = (w >> 16).astype(np.int64), (w & 0xFFFF).astype(np.int32)
        raw = np.concatenate([[0], np.asarray(index_taxa, dtype=np.int64), [0]])   # label 0 = no record; the pad is never read
        out = np.zeros(len(w), dtype=HIT_DTYPE)
        amb, border = label == 0xFFFF, w == 0xFFFFFFFF
        out["taxon"] = np.where(border, -2, np.where(amb, -1, raw[np.minimum(label, len(raw) - 1)]))
        out["count"] = np.where(border, -(k - 1), count)
        return out

    def lengths_from_hits(self, hits: np.ndarray, k: int, paired: bool):
        """(len1, len2) of every read from its hit list (HIT_DTYPE records in read order): the k-mers of the hits before the
        mate-pair border + k - 1, likewise after it; len2 = 0xFFFFFFFF for single-end reads (slk_read_result's fields, which
        the short results leave out)."""
        cnt = self.hit_cnt.astype(np.int64)
        read_of = np.repeat(np.arange(len(cnt), dtype=np.int64), cnt)
        border = hits["taxon"] == -2
        after = np.cumsum(border) - np.repeat((np.cumsum(np.bincount(read_of, weights=border, minlength=len(cnt))) -
                                               np.bincount(read_of, weights=border, minlength=len(cnt))).astype(np.int64), cnt)
        c = np.where(border, 0, hits["count"]).astype(np.int64)
        k1 = np.bincount(read_of, weights=np.where(after == 0, c, 0), minlength=len(cnt)).astype(np.int64)
        k2 = np.bincount(read_of, weights=np.where(after > 0, c, 0), minlength=len(cnt)).astype(np.int64)
        len1 = (k1 + k - 1).astype(np.uint32)
        len2 = (k2 + k - 1).astype(np.uint32) if paired else np.full(len(cnt), 0xFFFFFFFF, dtype=np.uint32)
        return len1, len2


def compact_reads(p: "PackedReads") -> CompactReads:
    """PackedReads -> CompactReads: drops the block offsets and turns the mask words into a list of positions."""
    nz = np.nonzero(p.mask)[0]
    if len(nz) == 0:
        e = np.zeros(0, dtype=np.uint32)
        return CompactReads(p.codes, p.len, e, e)
    blk_read = np.searchsorted(p.boff, nz, side="right") - 1
    reads, poss = [], []
    for b, r in zip(nz, blk_read):
        m = int(p.mask[b])
        base = 32 * (int(b) - int(p.boff[r]))
        for i in range(32):
            if (m >> i) & 1:
                reads.append(int(r)); poss.append(base + i)
    return CompactReads(p.codes, p.len, np.array(reads, dtype=np.uint32), np.array(poss, dtype=np.uint32))


def merge_ambiguous(r1: CompactReads, r2: Optional[CompactReads]) -> np.ndarray:
    """The ambiguity list of the ABI: read << 32 | mate << 31 | position, sorted by read."""
    a = (r1.amb_read.astype(np.uint64) << np.uint64(32)) | r1.amb_pos.astype(np.uint64)
    if r2 is not None and len(r2.amb_read):
        b = (r2.amb_read.astype(np.uint64) << np.uint64(32)) | np.uint64(1 << 31) | r2.amb_pos.astype(np.uint64)
        a = np.concatenate([a, b])
        a = a[np.argsort(a >> np.uint64(32), kind="stable")]
    return np.ascontiguousarray(a, dtype=np.uint64)


def block_offsets(off: np.ndarray) -> np.ndarray:
    """Block offsets of a packed batch: exclusive prefix of ceil(len / 32)."""
    ln = np.diff(off.astype(np.int64))
    boff = np.zeros(len(off), dtype=np.uint64)
    boff[1:] = np.cumsum((ln + 31) // 32)
    return boff


def pack_reads(bases: np.ndarray, off: np.ndarray) -> PackedReads:
    """Host-side packer (numpy): ASCII bases + offsets -> PackedReads. The JVM binding does the same per character."""
    off = off.astype(np.int64)
    n = len(off) - 1
    ln = np.diff(off)
    boff = block_offsets(off)
    nblk = int(boff[-1])
    code = _CODE[bases[off[0]:off[-1]]]
    # position of every base inside the padded block layout
    read_of = np.repeat(np.arange(n), ln)
    pos_in_read = np.arange(len(code)) - np.repeat(off[:-1] - off[0], ln)
    slot = boff[:-1].astype(np.int64)[read_of] * 32 + pos_in_read
    padded_code = np.zeros(nblk * 32, dtype=np.uint64)
    padded_mask = np.zeros(nblk * 32, dtype=np.uint32)
    padded_code[slot] = code & 3
    padded_mask[slot] = code >> 2
    sh = (2 * np.arange(32, dtype=np.uint64))[None, :]
    codes = np.bitwise_or.reduce(padded_code.reshape(nblk, 32) << sh, axis=1) if nblk else np.zeros(0, dtype=np.uint64)
    msh = np.arange(32, dtype=np.uint32)[None, :]
    mask = np.bitwise_or.reduce(padded_mask.reshape(nblk, 32) << msh, axis=1).astype(np.uint32) if nblk else np.zeros(0, dtype=np.uint32)
    return PackedReads(np.ascontiguousarray(codes), np.ascontiguousarray(mask), boff, ln.astype(np.uint32))


def pack_sequences(seqs):
    """list of bytes/str -> (uint8 bases, uint64 offsets[n+1]): what the JVM side assembles in a pinned buffer. Line breaks
    inside a sequence are dropped here, as the reference's scanner skips them (kmers/minimizer/ShiftScanner.scala:113-120):
    the device input is one sequence per offset range, without line breaks."""
    bs = [(s.encode("latin-1") if isinstance(s, str) else bytes(s)).translate(None, b"\n\r") for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    joined = b"".join(bs)
    bases = np.frombuffer(joined, dtype=np.uint8).copy() if joined else np.zeros(16, dtype=np.uint8)[:0]
    return bases, off
