"""CPU checks of the product's boundary: the C-ABI library loads, exports every symbol that
include/tsq_b200.h declares, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tsq_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(tsqb_\w+|tsq[A-Z]\w+)\s*\(", text))
    return sorted(names)


def test_header_declares_the_reference_entry_points():
    names = declared_symbols()
    for ref in ("tsqEncode", "tsqDecode", "tsqInit", "tsqAllocateContext", "tsqDeallocateContext", "tsqCompress", "tsqDecompress",
                "tsqCompress_MT", "tsqDecompress_MT", "tsqAllocateContextCompression_MT", "tsqDeallocateContextCompression_MT",
                "tsqAllocateContextDecompression_MT", "tsqDeallocateContextDecompression_MT"):
        assert ref in names


def test_library_exports_every_declared_symbol():
    import turbosqueeze_b200 as T
    lib = C.CDLL(T.library_path())
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/tsq_b200.h but not exported"


def test_no_cpu_fallback():
    import turbosqueeze_b200 as T
    L = T.library()
    if L.tsqb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(T.TsqError, match="no CUDA device"):
        T.Context(0)
    with pytest.raises(T.TsqError):
        T.tsqEncode(b"hello world hello world")


def test_slot_stride_matches_worst_case():
    import turbosqueeze_b200 as T
    from oraclelib import slot_stride
    for b in (1, 15, 16, 17, 4096, 65536, 262144, 1 << 20, 1 << 22):
        assert T.slot_stride(b) == slot_stride(b)
        assert T.slot_stride(b) >= 5 + b + (b + 15) // 16 * 2  # header + ctl + size bytes + literals


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under turbosqueeze_b200/ may reference it."""
    pkg = os.path.join(ROOT, "turbosqueeze_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oraclelib" not in src and "liboracle" not in src and "libtsq_ref" not in src, f
