"""ctypes binding of libslacken_gpu.so (include/slacken_gpu.h). There is no CPU fallback: if the shared library
has not been built, or no CUDA device is present, the failure is raised to the caller."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, os.environ.get("SLK_SO", "libslacken_gpu.so"))

SLK_OK, SLK_E_INVALID, SLK_E_CUDA, SLK_E_NOMEM, SLK_E_NOSPACE, SLK_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
READ_CLASSIFIED, READ_HAS_SPAN = 1, 2


class SlackenGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libslacken_gpu error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("k", C.c_int32), ("m", C.c_int32), ("spaces", C.c_int32), ("canonical", C.c_int32),
                ("toggle_mask", C.c_uint64)]


MAX_THRESHOLDS = 8


class ClassifyMultiOpts(C.Structure):
    _fields_ = [("n_thresholds", C.c_uint32), ("min_hit_groups", C.c_int32), ("confidence", C.c_double * MAX_THRESHOLDS)]


class ClassifyOpts(C.Structure):
    _fields_ = [("confidence", C.c_double), ("min_hit_groups", C.c_int32), ("reserved", C.c_int32)]


# every exported symbol of include/slacken_gpu.h: name -> (restype, argtypes)
_VP, _U64, _U32, _I32, _INT = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_int
_PP = C.POINTER(C.c_void_p)
SIGNATURES = {
    "slk_last_error": (C.c_char_p, []),
    "slk_ctx_create": (_INT, [_INT, _PP]),
    "slk_ctx_destroy": (None, [_VP]),
    "slk_ctx_device": (_INT, [_VP]),
    "slk_ctx_sync": (_INT, [_VP]),
    "slk_host_alloc": (_INT, [C.c_size_t, _PP]),
    "slk_host_free": (None, [_VP]),
    "slk_host_register": (_INT, [_VP, C.c_size_t]),
    "slk_host_unregister": (_INT, [_VP]),
    "slk_params_init": (_INT, [_INT, _INT, _INT, _U64, _INT, C.POINTER(Params)]),
    "slk_taxonomy_create": (_INT, [_VP, _VP, _I32, _PP]),
    "slk_taxonomy_destroy": (None, [_VP]),
    "slk_index_from_records": (_INT, [_VP, _VP, C.POINTER(Params), _VP, _VP, _U64, _PP]),
    "slk_index_destroy": (None, [_VP]),
    "slk_index_size": (_U64, [_VP]),
    "slk_index_records": (_INT, [_VP, _VP, _VP, _U64, C.POINTER(_U64)]),
    "slk_build_begin": (_INT, [_VP, _VP, C.POINTER(Params), _U64, _PP]),
    "slk_build_add": (_INT, [_VP, _VP, _VP, _VP, _U32]),
    "slk_build_add_dev": (_INT, [_VP, _VP, _VP, _VP, _U32, _U64]),
    "slk_build_finish": (_INT, [_VP, _PP]),
    "slk_build_destroy": (None, [_VP]),
    "slk_classifier_create": (_INT, [_VP, _PP]),
    "slk_classifier_destroy": (None, [_VP]),
    "slk_classify_hits_bound": (_U64, [C.POINTER(Params), _U32, _U64, _INT]),
    "slk_classify_batch": (_INT, [_VP, C.POINTER(ClassifyOpts), _VP, _VP, _VP, _VP, _U32, _VP, _VP, _VP, _VP, _U64,
                                  C.POINTER(_U64)]),
    "slk_classify_batch_dev": (_INT, [_VP, C.POINTER(ClassifyOpts), _VP, _VP, _VP, _VP, _U32, _VP, _VP, _VP, _VP, _U64,
                                      _VP]),
    "slk_classify_batch_packed": (_INT, [_VP, C.POINTER(ClassifyOpts), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP, _VP,
                                         _VP, _VP, _U64, C.POINTER(_U64)]),
    "slk_classify_batch_packed_multi": (_INT, [_VP, C.POINTER(ClassifyMultiOpts), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP, _VP,
                                               _VP, _VP, _U64, C.POINTER(_U64)]),
    "slk_classify_batch_compact": (_INT, [_VP, C.POINTER(ClassifyMultiOpts), _VP, _VP, _VP, _VP, _VP, _U64, _U32, _VP, _VP, _VP, _VP, _U64,
                                          C.POINTER(_U64)]),
    "slk_classify_batch_compact_short": (_INT, [_VP, C.POINTER(ClassifyMultiOpts), _VP, _VP, _VP, _VP, _VP, _U64, _U32, _VP, _VP, _VP, _VP,
                                                _U64, C.POINTER(_U64)]),
    "slk_classify_packed_dev": (_INT, [_VP, C.POINTER(ClassifyOpts), _VP, _VP, _VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP, _VP,
                                       _VP, _VP, _U64, _VP]),
    "slk_pack_reads_dev": (_INT, [_VP, _VP, _VP, _U32, _VP, _VP, _VP, _VP]),
    "slk_classifier_sync": (_INT, [_VP]),
    "slk_classifier_stream": (_VP, [_VP]),
    "slk_classifier_launches": (_U64, [_VP]),
    "slk_classifier_stats": (_INT, [_VP, C.POINTER(_U64), C.POINTER(_U64)]),
    "slk_event_create": (_INT, [_VP, _PP]),
    "slk_event_destroy": (None, [_VP]),
    "slk_event_record": (_INT, [_VP, _VP]),
    "slk_event_record_ctx": (_INT, [_VP, _VP]),
    "slk_event_elapsed_ms": (_INT, [_VP, _VP, C.POINTER(C.c_float)]),
    "slk_counts_create": (_INT, [_VP, _VP, _I32, _PP]),
    "slk_counts_destroy": (None, [_VP]),
    "slk_classifier_attach_counts": (_INT, [_VP, _VP, _I32]),
    "slk_counts_add": (_INT, [_VP, _VP, _VP, _VP, _U32]),
    "slk_counts_fetch": (_INT, [_VP, _I32, _VP, _I32]),
    "slk_counts_device_ptr": (_VP, [_VP]),
    "slk_counts_reset": (_INT, [_VP]),
    "slk_synth_genome_dev": (_INT, [_VP, _U64, _U64, _U64, _VP]),
    "slk_synth_reads_dev": (_INT, [_VP, _U64, _U64, _U64, _U64, _U64, _U64, _U32, _VP]),
    "slk_synth_mates_dev": (_INT, [_VP, _U64, _U64, _U64, _U64, _U64, _U64, _U32, _U32, _VP]),
    "slk_dev_alloc": (_INT, [_VP, C.c_size_t, _PP]),
    "slk_dev_free": (None, [_VP, _VP]),
    "slk_memcpy_h2d": (_INT, [_VP, _VP, _VP, C.c_size_t]),
    "slk_memcpy_d2h": (_INT, [_VP, _VP, _VP, C.c_size_t]),
    "slk_debug_sort_u64": (_INT, [_VP, _VP, _U64, _INT, _INT]),
    "slk_debug_min62": (_INT, [_VP, _VP, _VP, _U64, _VP]),
    "slk_index_from_records_shard": (_INT, [_VP, _VP, _VP, _VP, _VP, _U64, _U32, _VP]),
    "slk_build_reduce": (_INT, [_VP, _U32, _VP]),
    "slk_build_cells_dev": (_INT, [_VP, _VP, _VP]),
    "slk_build_dense_taxa": (_INT, [_VP, _VP, _U32, _VP]),
    "slk_index_from_cell_runs": (_INT, [_VP, _VP, _VP, _U32, _U32, _VP, _VP, _VP, _VP, _VP]),
    "slk_shard_of_records": (_INT, [_VP, _VP, _U64, _U32, _VP]),
    "slk_shard_of_records_dev": (_INT, [_VP, _VP, _VP, _U64, _U32, _VP]),
    "slk_index_records_by_owner_dev": (_INT, [_VP, _U32, _VP, _VP, _U64, _VP]),
    "slk_index_taxa": (_INT, [_VP, _VP, _U32, _VP]),
    "slk_resolver_create": (_INT, [_VP, _VP, _VP, _VP, _U32, _PP]),
    "slk_resolver_destroy": (None, [_VP]),
    "slk_scan_spans_dev": (_INT, [_VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP, _VP, _U64, _VP]),
    "slk_bracken_weights": (_INT, [_VP, _VP, _VP, _VP, _U32, _U32, _VP, _U64, _VP]),
    "slk_emit_spans_dev": (_INT, [_VP, _VP, _VP, _VP, _VP, _VP, _U32, _VP, _VP]),
    "slk_route_spans_dev": (_INT, [_VP, _VP, _U64, _U32, _VP, _VP, _U64, _VP]),
    "slk_probe_keys_dev": (_INT, [_VP, _VP, _U64, _VP]),
    "slk_resolve_spans_dev": (_INT, [_VP, _VP, _VP, _VP, _U64, _U32, _INT, _VP, _VP, _U64, _VP, _VP, _VP, _VP]),
    "slk_memcpy_d2d": (_INT, [_VP, _VP, _VP, C.c_size_t]),
    "slk_mailbox_create": (_INT, [_VP, _U32, _U32, _U64, _PP, _VP]),
    "slk_mailbox_connect": (_INT, [_VP, _VP]),
    "slk_mailbox_connect_local": (_INT, [_PP, _U32]),
    "slk_mailbox_destroy": (None, [_VP]),
    "slk_mailbox_set_blocks_per_sm": (_INT, [_VP, _U32]),
    "slk_mailbox_route": (_INT, [_VP, _VP, _U64]),
    "slk_mailbox_probe": (_INT, [_VP, _VP]),
    "slk_mailbox_resolve": (_INT, [_VP, _VP, _VP, _VP, _VP, _U64, _U32, _INT, _VP, _VP, _VP, _VP]),
    "slk_mailbox_resolve_async": (_INT, [_VP, _VP, _VP, _VP, _VP, _U64, _U32, _INT, _VP, _VP, _VP, _VP]),
    "slk_mailbox_resolve_wait": (_INT, [_VP, _VP, _U32, _VP, _VP, _VP, _VP]),
}
IPC_HANDLE_BYTES = 64

_lib = None


def load() -> C.CDLL:
    """Load the shared library (no build attempt here: __graft_entry__.build() / `python -m slacken_b200.build`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise SlackenGpuError(SLK_E_CUDA, f"{SO_PATH} is missing: run `python -m slacken_b200.build` (nvcc, sm_100a). "
                                  "There is no CPU fallback.")
        L = C.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the ABI symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != SLK_OK:
        raise SlackenGpuError(rc, load().slk_last_error().decode("utf-8", "replace"))
