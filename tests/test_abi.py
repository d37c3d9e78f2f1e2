"""The C-ABI shared library loads and exports every symbol include/slacken_gpu.h declares (no compute calls here)."""
import ctypes
import os
import re

from slacken_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "slacken_gpu.h")).read()
    return sorted(set(re.findall(r"^SLK_API [^;(]*?\b(slk_[a-z0-9_]+)\(", text, flags=re.M)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 40
    lib = ctypes.CDLL(_lib.SO_PATH)
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms     # the Python binding covers the whole header


def test_errors_are_reported_not_thrown():
    L = _lib.load()
    p = _lib.Params()
    assert L.slk_params_init(35, 31, 7, 0xE37E28C4271B5A2D, 1, ctypes.byref(p)) == 0
    assert (p.k, p.m, p.spaces, p.canonical) == (35, 31, 7, 1)
    assert L.slk_params_init(35, 31, 0, 0, 1, ctypes.byref(p)) == _lib.SLK_E_UNSUPPORTED   # 62 key bits > 48
    assert b"48" in L.slk_last_error()
    assert L.slk_params_init(30, 31, 7, 0, 1, ctypes.byref(p)) < 0


def test_no_cpu_fallback_without_a_device():
    """On a machine without CUDA the product path must fail loudly instead of computing on the CPU."""
    L = _lib.load()
    h = ctypes.c_void_p()
    rc = L.slk_ctx_create(0, ctypes.byref(h))
    if rc == 0:
        L.slk_ctx_destroy(h)      # a GPU is present: nothing to check here
    else:
        assert rc == _lib.SLK_E_CUDA and b"no CPU fallback" in L.slk_last_error()
