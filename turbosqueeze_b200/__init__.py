"""turbosqueeze_b200 -- B200 (sm_100a) implementation of Turbosqueeze's per-block encode/decode path.

The product is the C-ABI shared library ``libturbosqueeze_b200.so`` (``include/tsq_b200.h``): CUDA
kernels plus the reference's own entry points (``tsqEncode`` / ``tsqDecode`` / ``tsqCompress`` ...,
reference ``turbosqueeze.h:458-670``).  This package is the host-side mirror used by the tests and
``bench.py``: thin ctypes bindings, with PyTorch only providing device memory and streams.

There is no CPU fallback: importing works anywhere, but every compute entry point raises when the
CUDA extension or a CUDA device is missing.
"""
from .api import (  # noqa: F401
    Context,
    TsqError,
    library,
    library_path,
    slot_stride,
    tsqDecode,
    tsqEncode,
    tsq_compress_mt,
    tsq_decompress_mt,
    INPUT_PAD,
    BLOCK_MAX,
)
