"""TEST-ONLY host emulation of the CUDA kernel bodies (slacken_b200/csrc/slk_core.h compiled with g++).
Never imported by the product package; see emu.cpp."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_DEFS = []
_SO = os.path.join(_HERE, "libslk_emu.so")
_CORE = os.path.join(_HERE, "..", "..", "slacken_b200", "csrc", "slk_core.h")
BUILD_WPT = 96


class ScanParams(C.Structure):
    _fields_ = [("k", C.c_int32), ("m", C.c_int32), ("w", C.c_int32), ("canonical", C.c_int32), ("fshift", C.c_int32),
                ("key_bits", C.c_int32), ("fast_compress", C.c_int32), ("pad_", C.c_int32), ("xor_mask", C.c_uint64), ("sig_mask", C.c_uint64), ("mmask", C.c_uint64),
                ("cmv", C.c_uint64 * 6)]


RESULT_DTYPE = np.dtype([("taxon", "<i4"), ("flags", "<u4"), ("kmers1", "<u4"), ("kmers2", "<u4"),
                         ("num_distinct", "<u4"), ("n_hits", "<u4")])
HIT_DTYPE = np.dtype([("taxon", "<i4"), ("count", "<i4")])
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "emu.cpp")
        newest = max(os.path.getmtime(src), os.path.getmtime(_CORE))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden"] + _DEFS +
                                  ["-o", _SO, src])
        L = C.CDLL(_SO)
        assert L.emu_sizeof_scan_params() == C.sizeof(ScanParams), "ScanParams mirror out of date"
        L.emu_scan_params.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.POINTER(ScanParams)]
        L.emu_compress.restype = C.c_uint64
        L.emu_compress.argtypes = [C.POINTER(ScanParams), C.c_uint64]
        L.emu_expand.restype = C.c_uint64
        L.emu_expand.argtypes = [C.POINTER(ScanParams), C.c_uint64]
        L.emu_code.restype = C.c_uint32
        L.emu_code.argtypes = [C.c_uint32]
        L.emu_code4_mismatches.argtypes = [C.c_void_p, C.c_uint64]
        L.emu_code4_mismatches.restype = C.c_uint64
        L.emu_buckets_for.restype = C.c_uint64
        L.emu_buckets_for.argtypes = [C.c_uint64]
        L.emu_insert_cells.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_uint32]
        L.emu_classify.restype = C.c_int64
        L.emu_classify.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                   C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                   C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.emu_classify_split.restype = C.c_int64
        L.emu_classify_split.argtypes = L.emu_classify.argtypes[:-1]
        L.emu_scan_spans.restype = C.c_int64
        L.emu_scan_spans.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                     C.c_void_p, C.c_void_p, C.c_uint64]
        L.emu_probe_keys.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
        L.emu_resolve_spans.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_uint32, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        L.emu_bracken.restype = C.c_int64
        L.emu_bracken.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                  C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.emu_shard_of.restype = C.c_uint32
        L.emu_shard_of.argtypes = [C.c_uint64, C.c_uint32]
        L.emu_key_mix.argtypes = [C.c_uint64]
        L.emu_key_mix.restype = C.c_uint32
        L.emu_bucket_of.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
        L.emu_bucket_of.restype = C.c_uint64
        L.emu_emit_cells.restype = C.c_int64
        L.emu_emit_cells.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64]
        L.emu_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.emu_synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        L.emu_synth_mates.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def scan_params(k, m, spaces, mask, canonical) -> ScanParams:
    sp = ScanParams()
    rc = lib().emu_scan_params(k, m, spaces, mask, 1 if canonical else 0, C.byref(sp))
    if rc:
        raise ValueError(f"unsupported parameters ({rc})")
    return sp


class DenseTax:
    """Python twin of the host-side dense taxonomy (dense_add in slacken_gpu.cu), test use only."""

    def __init__(self, parents: np.ndarray, taxa):
        self.raw = [0]
        self.parent = [0]
        self.depth = [0]
        self.to_dense = {0: 0}
        self.parents = parents
        self.root = self.add(1)
        for t in sorted(set(int(x) for x in taxa)):
            if t > 0:
                self.add(t)

    def add(self, raw: int) -> int:
        path = []
        t = raw
        while t not in self.to_dense:
            path.append(t)
            t = int(self.parents[t])
        pd = self.to_dense[t]
        for x in reversed(path):
            self.raw.append(x)
            self.parent.append(pd)
            self.depth.append(self.depth[pd] + 1)
            pd = len(self.raw) - 1
            self.to_dense[x] = pd
        return pd

    def arrays(self):
        return (np.array(self.parent, dtype=np.uint16), np.array(self.depth, dtype=np.uint8),
                np.array(self.raw, dtype=np.int32))


class EmuIndex:
    def __init__(self, sp: ScanParams, parents: np.ndarray, id1: np.ndarray, taxon: np.ndarray, world: int = 1):
        """world > 1: the records are one shard of a library cut for `world` ranks (the table then spreads the shard's range
        of the line hash over all of its lines, as slk_index_from_records_shard does)."""
        self.sp, self.world = sp, world
        self.dt = DenseTax(parents, taxon)
        self.parent, self.depth, self.raw = self.dt.arrays()
        self.n_buckets = int(lib().emu_buckets_for(len(id1)))
        self.cells = np.zeros(self.n_buckets * 4, dtype=np.uint64)
        cells_in = np.array([(lib().emu_compress(C.byref(sp), int(k)) << 16) | self.dt.to_dense[int(t)]
                             for k, t in zip(id1.view(np.uint64), taxon)], dtype=np.uint64)
        self.insert(cells_in)

    def insert(self, cells_in: np.ndarray):
        """(compressed key << 16 | dense taxon of self.dt) cells; equal keys merge by LCA."""
        cells_in = np.ascontiguousarray(cells_in, dtype=np.uint64)
        lib().emu_insert_cells(_p(self.cells), self.n_buckets, _p(self.parent), _p(self.depth), self.dt.root,
                               _p(cells_in), len(cells_in), self.world)

    def records(self):
        """(id1 uint64[], raw taxon int32[]) of the table, sorted by id1."""
        c = self.cells[self.cells != 0]
        id1 = np.array([lib().emu_expand(C.byref(self.sp), int(x) >> 16) for x in c], dtype=np.uint64)
        tx = np.array([self.raw[int(x) & 0xffff] for x in c], dtype=np.int32)
        o = np.argsort(id1, kind="stable")
        return id1[o], tx[o]

    def classify(self, bases1, off1, bases2=None, off2=None, confidence=0.0, min_hit_groups=2, packed=False, split=False):
        """split=True: the scan | probe | merge+resolve bodies of the sharded-library path instead of the fused one."""
        n = len(off1) - 1
        off1 = np.ascontiguousarray(off1, dtype=np.uint64)
        if bases2 is not None:
            off2 = np.ascontiguousarray(off2, dtype=np.uint64)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        cap = int(off1[-1]) + (int(off2[-1]) if bases2 is not None else 0) + 5 * n + 8
        hits = np.zeros(cap, dtype=HIT_DTYPE)
        hit_off = np.zeros(n + 1, dtype=np.uint64)
        if split:
            used = lib().emu_classify_split(C.byref(self.sp), _p(self.cells), self.n_buckets, _p(self.parent), _p(self.depth),
                                            _p(self.raw), len(self.raw), self.dt.root, _p(bases1), _p(off1), _p(bases2),
                                            _p(off2), n, float(confidence), int(min_hit_groups), _p(res), _p(hit_off), _p(hits), cap)
        else:
            used = lib().emu_classify(C.byref(self.sp), _p(self.cells), self.n_buckets, _p(self.parent), _p(self.depth),
                                      _p(self.raw), len(self.raw), self.dt.root, _p(bases1), _p(off1), _p(bases2), _p(off2), n,
                                      float(confidence), int(min_hit_groups), _p(res), _p(hit_off), _p(hits), cap, 1 if packed else 0)
        assert used >= 0
        hit_off[n] = used
        return res, hit_off, hits[:used]


_SO2 = os.path.join(_HERE, "libslk_emu2.so")
_lib2 = None
DETAIL_DTYPE = np.dtype([("hit_off", "<u8"), ("hit_cnt", "<u4"), ("len1", "<u4"), ("len2", "<u4"), ("num_distinct", "<u4")])


def lib2():
    """The warp-cooperative classify kernel body (slk_group.h) under the fibre emulation of simt.h."""
    global _lib2
    if _lib2 is None:
        srcs = [os.path.join(_HERE, "emu2.cpp"), os.path.join(_HERE, "simt.h"), _CORE, os.path.join(os.path.dirname(_CORE), "slk_group.h")]
        if not os.path.exists(_SO2) or os.path.getmtime(_SO2) < max(os.path.getmtime(x) for x in srcs):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", _SO2, srcs[0]])
        _lib2 = C.CDLL(_SO2)
        _lib2.emu2_classify.restype = C.c_int
        _lib2.emu2_classify.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32,
                                        C.c_uint32] + [C.c_void_p] * 8 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.c_int] + [C.c_void_p] * 4 + \
                                       [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib2


def classify2(ix: "EmuIndex", r1, r2=None, confidence=0.0, min_hit_groups=2, want_hits=True):
    """r1 / r2: slacken_b200.host.PackedReads. Returns (taxon, flags, detail, hits, probes, merged_hits, counts); with a list
    of confidences taxon and flags have one row per threshold."""
    n = len(r1.len)
    multi = not np.isscalar(confidence)
    conf = np.atleast_1d(np.asarray(confidence, dtype=np.float64)).copy()
    taxon = np.zeros((len(conf), n), dtype=np.int32)
    flags = np.zeros((len(conf), n), dtype=np.uint8)
    detail = np.zeros(n, dtype=DETAIL_DTYPE)
    cap = int(r1.len.sum()) + (int(r2.len.sum()) if r2 is not None else 0) + 5 * n + 8
    hits = np.zeros(cap, dtype=HIT_DTYPE)
    used = C.c_uint64(0)
    stats = np.zeros(2, dtype=np.uint64)
    counts = np.zeros(int(ix.raw.max()) + 1, dtype=np.uint64)
    m2 = (r2.codes, r2.mask, r2.boff, r2.len) if r2 is not None else (None, None, None, None)
    pad = lambda a, dt: np.concatenate([a, np.zeros(4, dtype=dt)])   # the kernel never reads past a read's blocks; be strict anyway
    rc = lib2().emu2_classify(C.byref(ix.sp), _p(ix.cells), ix.n_buckets, _p(ix.parent), _p(ix.depth), _p(ix.raw), len(ix.raw),
                              ix.dt.root, _p(r1.codes), _p(r1.mask), _p(r1.boff), _p(r1.len), _p(m2[0]), _p(m2[1]), _p(m2[2]), _p(m2[3]),
                              n, _p(conf), len(conf), int(min_hit_groups), 1 if want_hits else 0, _p(taxon), _p(flags), _p(detail),
                              _p(hits), cap, C.byref(used), _p(stats), _p(counts))
    assert rc == 0, rc
    if not multi:
        taxon, flags = taxon[0], flags[0]
    return taxon, flags, detail, hits[:used.value], int(stats[0]), int(stats[1]), counts


def bracken_dests(ix: "EmuIndex", seq: bytes, read_len: int) -> np.ndarray:
    """Destination taxon (raw id) of every read of length read_len of one genome fragment."""
    b = np.frombuffer(seq, dtype=np.uint8).copy() if len(seq) else np.zeros(1, dtype=np.uint8)
    out = np.zeros(max(len(seq), 1), dtype=np.int32)
    n = lib().emu_bracken(C.byref(ix.sp), _p(ix.cells), ix.n_buckets, _p(ix.parent), _p(ix.depth), _p(ix.raw), len(ix.raw),
                          ix.dt.root, _p(b), len(seq), read_len, _p(out))
    assert n >= 0
    return out[:n]


def emit_cells(sp: ScanParams, seq: bytes, dense_taxon: int, wpt: int = BUILD_WPT) -> np.ndarray:
    b = np.frombuffer(seq, dtype=np.uint8)
    out = np.zeros(len(b) + 1, dtype=np.uint64)
    n = lib().emu_emit_cells(C.byref(sp), _p(b), len(b), dense_taxon, wpt, _p(out), len(out))
    assert n >= 0
    return out[:n]


def expand(sp: ScanParams, x: int) -> int:
    return lib().emu_expand(C.byref(sp), x)


def compress(sp: ScanParams, x: int) -> int:
    return lib().emu_compress(C.byref(sp), x)
