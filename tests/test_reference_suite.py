"""The reference's OWN test program, unmodified, against the product library.

oracle/Makefile compiles /root/reference/test/test.cpp where it lies, against the reference's own headers
(turbosqueeze.h via ../tsq_context.h), and links it with libturbosqueeze_b200.so instead of the reference's library:
oracle/_ref/testturbosqueeze_b200 (git-ignored, travels to the GPU box like oracle/_ref/libtsq_ref.so).  The ten cases
are the ones the reference registers with CTest (test/CMakeLists.txt:9-18); each passes when the program exits 0.
"""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "testturbosqueeze_b200")

# test/CMakeLists.txt:9-18, in that order
CASES = [
    "test_tsq_context",
    "test_tsq_compress",
    "test_tsq_context_mt",
    "test_tsq_compress_mt",
    "test_tsq_queue_mt",
    "test_tsq_context_mt2",
    "test_tsq_decompress_mt",
    "test_tsq_compress_async_mt",
    "test_tsq_decompress_async_mt",
    "test_tsq_massive_async_mt",
]


def test_reference_test_program_links_to_the_product_only():
    """CPU check: the binary exists when the reference sources were present at build time, and its only
    turbosqueeze dependency is the product library (not the reference's)."""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/testturbosqueeze_b200 not built (reference sources absent at build time)")
    out = subprocess.run(["ldd", BIN], capture_output=True, text=True).stdout
    assert "libturbosqueeze_b200.so" in out and "libtsq_ref" not in out


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_reference_case(case):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/testturbosqueeze_b200 not built (reference sources absent at build time)")
    r = subprocess.run([BIN, case], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (case, r.returncode, r.stdout[-2000:], r.stderr[-2000:])
