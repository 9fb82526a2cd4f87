"""Deterministic synthetic inputs for BASELINE.json's configs.

Thin ctypes front end over csrc/tsq_workload.c (built by __graft_entry__.build()).
Every buffer is returned followed by PAD zero bytes, the parity contract of the
reference's memory path (tsq_threads.cpp:109; sample/main.cpp:55 allocates slack).
"""
import ctypes as C
import os

import numpy as np

PAD = 128
KINDS = {"text": 0, "random": 1, "rep8": 2}
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _load():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libtsq_workload.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(path)
        _lib.tsqw_fill.restype = None
        _lib.tsqw_fill.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _lib.tsqw_text_params.restype = None
        _lib.tsqw_text_params.argtypes = [C.c_double, C.c_uint32, C.c_uint32, C.c_uint32]
    return _lib


def fill(kind, n, seed=1, offset=0, out=None, threads=None):
    """Bytes [offset, offset+n) of stream (kind, seed), followed by PAD zeros."""
    lib = _load()
    if out is None:
        out = np.zeros(n + PAD, dtype=np.uint8)
    assert out.dtype == np.uint8 and out.size >= n + PAD
    threads = threads or min(32, os.cpu_count() or 1)
    lib.tsqw_fill(KINDS[kind], seed, offset, n, out.ctypes.data, threads)
    out[n:n + PAD] = 0
    return out


def text_params(zipf=1.355, rare=30, markup=80, number=40):
    _load().tsqw_text_params(zipf, rare, markup, number)
