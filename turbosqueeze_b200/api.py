"""ctypes mirror of include/tsq_b200.h.

Layer 1 (`Context.encode_blocks` / `decode_blocks` / `pack_container` / `index_container`) works on
torch uint8 CUDA tensors (device-resident batch path).  Layer 2 mirrors the reference interface
(`tsqEncode`, `tsqDecode`, `tsqCompress_MT`, `tsqDecompress_MT`; reference turbosqueeze.h:508-670)
on host bytes, so the parity tests read like the reference's own tests (test/test.cpp:30-54,149-199).
"""
import ctypes as C
import os

import numpy as np

INPUT_PAD = 128          # TSQB_INPUT_PAD
BLOCK_MAX = 1 << 22      # TSQ_BLOCK_SZ (reference turbosqueeze.h:38)

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_libs = {}

_u8p = C.POINTER(C.c_uint8)
_vp = C.c_void_p


class TsqError(RuntimeError):
    pass


def library_path():
    """The product library.  TSQB_LIBRARY (development only: scripts/build_variants.sh builds kernel variants for A/B
    timing) overrides it; bench.py refuses to run with it set unless told so, and prints the path it loaded."""
    return os.environ.get("TSQB_LIBRARY") or os.path.join(_HERE, "libturbosqueeze_b200.so")


def xcheck_library_path():
    """Test-only library: the product's objects plus the superseded round-1 kernels (encode_impl 2, decode_lanes 1..33)."""
    return os.path.join(os.path.dirname(_HERE), "tests", "xcheck", "libturbosqueeze_b200_xcheck.so")


def library(path=None):
    """Load libturbosqueeze_b200.so (or the library at `path`); fails loudly when it has not been built."""
    global _lib
    if path is None and _lib is not None:
        return _lib
    product = path is None
    path = path or library_path()
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise TsqError(f"{path} is missing: the CUDA extension is required (run `make` or __graft_entry__.build()); "
                       "there is no CPU fallback")
    L = C.CDLL(path)
    L.tsqb_last_error.restype = C.c_char_p
    L.tsqb_device_count.restype = C.c_int
    L.tsqb_launch_count.restype = C.c_uint64
    L.tsqb_create.argtypes = [C.POINTER(_vp), C.c_int]
    L.tsqb_destroy.argtypes = [_vp]
    L.tsqb_destroy.restype = None
    L.tsqb_slot_stride.argtypes = [C.c_uint32]
    L.tsqb_slot_stride.restype = C.c_uint64
    L.tsqb_set_option.argtypes = [_vp, C.c_char_p, C.c_int64]
    L.tsqb_encode_blocks.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, _vp, C.c_uint64, _vp, C.c_uint32, _vp]
    L.tsqb_decode_blocks.argtypes = [_vp, _vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint32, _vp]
    L.tsqb_pack_container.argtypes = [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.c_uint64, C.c_uint32, _vp, _vp, _vp]
    L.tsqb_index_container.argtypes = [_vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, _vp]
    L.tsqb_encode_host.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, _vp, _vp, C.c_uint32]
    L.tsqb_decode_host.argtypes = [_vp, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint64, _vp, C.c_uint32]
    L.tsqb_compress_buffer.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(_vp), C.POINTER(C.c_uint64)]
    L.tsqb_decompress_buffer.argtypes = [_vp, _vp, C.c_uint64, C.POINTER(_vp), C.POINTER(C.c_uint64)]
    L.tsqb_compress_into.argtypes = [_vp, _vp, C.c_uint64, C.c_uint32, C.c_uint32, _vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.tsqb_decompress_into.argtypes = [_vp, _vp, C.c_uint64, _vp, C.c_uint64, C.POINTER(C.c_uint64)]
    L.tsqb_ipc_export.argtypes = [_vp, _vp, C.POINTER(C.c_uint64)]
    L.tsqb_ipc_open.argtypes = [_vp, C.POINTER(_vp)]
    L.tsqb_ipc_close.argtypes = [_vp]
    L.tsqb_copy_d2d.argtypes = [_vp, _vp, C.c_uint64, _vp]
    L.tsqAllocateContext.restype = _vp
    L.tsqDeallocateContext.argtypes = [_vp]
    L.tsqDeallocateContext.restype = None
    L.tsqInit.argtypes = [_vp]
    L.tsqInit.restype = None
    L.tsqEncode.argtypes = [_vp, _vp, _vp, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    L.tsqEncode.restype = None
    L.tsqDecode.argtypes = [_vp, _vp, C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    L.tsqDecode.restype = None
    L.tsqAllocateContextCompression_MT.argtypes = [C.c_bool]
    L.tsqAllocateContextCompression_MT.restype = _vp
    L.tsqDeallocateContextCompression_MT.argtypes = [_vp]
    L.tsqDeallocateContextCompression_MT.restype = None
    L.tsqAllocateContextDecompression_MT.argtypes = [C.c_bool]
    L.tsqAllocateContextDecompression_MT.restype = _vp
    L.tsqDeallocateContextDecompression_MT.argtypes = [_vp]
    L.tsqDeallocateContextDecompression_MT.restype = None
    L.tsqCompress_MT.argtypes = [_vp, _vp, C.c_size_t, C.c_bool, C.POINTER(_vp), C.POINTER(C.c_size_t), C.c_bool, C.c_bool, C.c_uint32]
    L.tsqCompress_MT.restype = C.c_bool
    L.tsqDecompress_MT.argtypes = [_vp, _vp, C.c_size_t, C.c_bool, C.POINTER(_vp), C.POINTER(C.c_size_t), C.c_bool]
    L.tsqDecompress_MT.restype = C.c_bool
    _libs[path] = L
    if product:
        _lib = L
    return L


_libc = C.CDLL(None)
_libc.free.argtypes = [_vp]
_libc.free.restype = None


def slot_stride(block):
    return int(library().tsqb_slot_stride(block))


def _check(status, what, lib=None):
    if status != 0:
        raise TsqError(f"{what}: {(lib or library()).tsqb_last_error().decode()}")


def _stream_handle(stream):
    import torch
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def _as_np(data):
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(bytes(data), dtype=np.uint8)


# ---- peer memory for the multi-GPU gather (sharding.py) --------------------------------------------------
def ipc_export(d_ptr):
    """(64-byte CUDA IPC handle of the allocation d_ptr lies in, d_ptr's offset inside it)."""
    h = (C.c_uint8 * 64)()
    off = C.c_uint64(0)
    _check(library().tsqb_ipc_export(d_ptr, h, C.byref(off)), "tsqb_ipc_export")
    return bytes(h), off.value


def ipc_open(handle):
    base = _vp()
    buf = (C.c_uint8 * 64).from_buffer_copy(handle)
    _check(library().tsqb_ipc_open(buf, C.byref(base)), "tsqb_ipc_open")
    return base.value


def ipc_close(base):
    _check(library().tsqb_ipc_close(base), "tsqb_ipc_close")


def copy_d2d(dst_ptr, src_ptr, n, stream=None):
    _check(library().tsqb_copy_d2d(dst_ptr, src_ptr, n, _stream_handle(stream)), "tsqb_copy_d2d")


class Context:
    """tsqb_context: one CUDA device + the hash-table scratch of the blocks in flight."""

    def __init__(self, device=0, lib=None):
        """lib: a library handle from library(path) (tests use the cross-check library); default = the product."""
        self._L = L = lib or library()
        h = _vp()
        _check(L.tsqb_create(C.byref(h), int(device)), "tsqb_create", L)
        self._h, self.device = h, int(device)

    def close(self):
        if getattr(self, "_h", None):
            self._L.tsqb_destroy(self._h)
            self._h = None

    __del__ = close

    def set_option(self, key, value):
        if self._L.tsqb_set_option(self._h, key.encode(), int(value)) != 0:
            raise TsqError(f"unknown option {key}")

    # ---- layer 1: device-resident batch path (torch uint8 CUDA tensors) -------------------------
    def encode_blocks(self, d_in, total, block, ext=0, slots=None, sizes=None, stream=None):
        """d_in: uint8 CUDA tensor holding `total` bytes followed by >= INPUT_PAD readable bytes.
        Returns (slots, sizes): block b's stream is slots[b*stride : b*stride + sizes[b]]."""
        import torch
        assert d_in.is_cuda and d_in.dtype == torch.uint8 and d_in.numel() >= total + INPUT_PAD
        nb = (total + block - 1) // block
        stride = slot_stride(block)
        if slots is None:
            slots = torch.zeros(max(nb, 1) * stride, dtype=torch.uint8, device=d_in.device)
        if sizes is None:
            sizes = torch.zeros(max(nb, 1), dtype=torch.int32, device=d_in.device)
        _check(self._L.tsqb_encode_blocks(self._h, d_in.data_ptr(), total, block, slots.data_ptr(), stride, sizes.data_ptr(),
                                            int(ext), _stream_handle(stream)), "tsqb_encode_blocks", self._L)
        return slots, sizes[:nb]

    def decode_blocks(self, d_comp, nb, block, ext=0, stride=None, offsets=None, comp_sizes=None, out=None, out_sizes=None,
                      stream=None):
        """Decode nb streams (at b*stride, or at offsets[b]) into out[b*block : ...]."""
        import torch
        if out is None:
            out = torch.empty(max(nb, 1) * block, dtype=torch.uint8, device=d_comp.device)
        if out_sizes is None:
            out_sizes = torch.zeros(max(nb, 1), dtype=torch.int32, device=d_comp.device)
        stride = slot_stride(block) if stride is None else stride
        _check(self._L.tsqb_decode_blocks(self._h, d_comp.data_ptr(), offsets.data_ptr() if offsets is not None else None,
                                            stride if offsets is None else 0,
                                            comp_sizes.data_ptr() if comp_sizes is not None else None, nb, out.data_ptr(), block,
                                            out_sizes.data_ptr(), int(ext), _stream_handle(stream)), "tsqb_decode_blocks", self._L)
        return out, out_sizes[:nb]

    def pack_container(self, slots, sizes, block, total, ext=0, stream=None):
        """TSQ1 container on the device (turbosqueeze.cpp:64-67,78-84). Returns (container, length tensor)."""
        import torch
        nb = sizes.numel() if total else 0
        stride = slot_stride(block)
        cont = torch.empty(16 + nb * (stride + 3) + 256, dtype=torch.uint8, device=slots.device)
        n = torch.zeros(1, dtype=torch.int64, device=slots.device)
        _check(self._L.tsqb_pack_container(self._h, slots.data_ptr(), stride, sizes.data_ptr(), nb, total, int(ext), cont.data_ptr(),
                                             n.data_ptr(), _stream_handle(stream)), "tsqb_pack_container", self._L)
        return cont, n

    def index_container(self, cont, csize, max_blocks, stream=None):
        import torch
        dev = cont.device
        offs = torch.zeros(max(max_blocks, 1), dtype=torch.int64, device=dev)
        sizes = torch.zeros(max(max_blocks, 1), dtype=torch.int32, device=dev)
        ext = torch.zeros(max(max_blocks, 1), dtype=torch.int32, device=dev)
        n = torch.zeros(1, dtype=torch.int64, device=dev)
        _check(self._L.tsqb_index_container(self._h, cont.data_ptr(), csize, max_blocks, offs.data_ptr(), sizes.data_ptr(),
                                              ext.data_ptr(), n.data_ptr(), _stream_handle(stream)), "tsqb_index_container", self._L)
        return offs, sizes, ext, n

    # ---- host buffers through the device ----------------------------------------------------------
    def encode_host(self, data, block, ext=0):
        a = _as_np(data)
        nb = (a.size + block - 1) // block
        stride = slot_stride(block)
        slots = np.zeros(max(nb, 1) * stride, dtype=np.uint8)
        sizes = np.zeros(max(nb, 1), dtype=np.uint32)
        _check(self._L.tsqb_encode_host(self._h, a.ctypes.data, a.size, block, slots.ctypes.data, sizes.ctypes.data, int(ext)),
               "tsqb_encode_host")
        return slots, sizes[:nb]

    def decode_host(self, slots, stride, comp_sizes, nb, block, ext=0):
        out = np.zeros(max(nb, 1) * block, dtype=np.uint8)
        osz = np.zeros(max(nb, 1), dtype=np.uint32)
        cs = np.ascontiguousarray(comp_sizes, dtype=np.uint32) if comp_sizes is not None else None
        _check(self._L.tsqb_decode_host(self._h, slots.ctypes.data, stride, cs.ctypes.data if cs is not None else None, nb,
                                          out.ctypes.data, block, osz.ctypes.data, int(ext)), "tsqb_decode_host", self._L)
        return out, osz[:nb]

    def compress_buffer(self, data, block=BLOCK_MAX, ext=0, ptr=None, size=None):
        """Host bytes -> TSQ1 container bytes (tsqCompress_MT memory->memory, tsq_threads.cpp:413-441)."""
        if ptr is None:
            a = _as_np(data)
            ptr, size = a.ctypes.data, a.size
        out, n = _vp(), C.c_uint64(0)
        _check(self._L.tsqb_compress_buffer(self._h, ptr, size, block, int(ext), C.byref(out), C.byref(n)), "tsqb_compress_buffer", self._L)
        try:
            return C.string_at(out, n.value)
        finally:
            _libc.free(out)

    def compress_into(self, in_ptr, total, block, ext, out_ptr, out_cap):
        """Host pointer -> TSQ1 container in a caller-owned host buffer; returns its length."""
        n = C.c_uint64(0)
        _check(self._L.tsqb_compress_into(self._h, in_ptr, total, block, int(ext), out_ptr, out_cap, C.byref(n)), "tsqb_compress_into", self._L)
        return n.value

    def decompress_into(self, in_ptr, in_size, out_ptr, out_cap):
        n = C.c_uint64(0)
        _check(self._L.tsqb_decompress_into(self._h, in_ptr, in_size, out_ptr, out_cap, C.byref(n)), "tsqb_decompress_into", self._L)
        return n.value

    def decompress_buffer(self, blob):
        a = _as_np(blob)
        out, n = _vp(), C.c_uint64(0)
        _check(self._L.tsqb_decompress_buffer(self._h, a.ctypes.data, a.size, C.byref(out), C.byref(n)), "tsqb_decompress_buffer", self._L)
        try:
            return C.string_at(out, n.value)
        finally:
            _libc.free(out)


# ---- layer 2: the reference's own entry points ------------------------------------------------------
def tsqEncode(data, ext=0, tail=b"", prefill=0):
    """reference tsqEncode (turbosqueeze.h:657) on one host block: tsqAllocateContext + tsqInit + tsqEncode.
    `tail` are the bytes following the block in memory (the reference reads <= 19 of them); zeros otherwise."""
    L = library()
    a = _as_np(data)
    src = np.zeros(a.size + INPUT_PAD, dtype=np.uint8)
    src[: a.size] = a
    t = _as_np(tail)[:INPUT_PAD]
    src[a.size: a.size + t.size] = t
    out = np.full(slot_stride(a.size) + 64, prefill, dtype=np.uint8)
    n = C.c_uint32(0)
    ctx = L.tsqAllocateContext()
    if not ctx:
        raise TsqError("tsqAllocateContext failed")
    try:
        L.tsqInit(ctx)
        L.tsqEncode(ctx, src.ctypes.data, out.ctypes.data, C.byref(n), a.size, int(ext))
    finally:
        L.tsqDeallocateContext(ctx)
    if a.size and n.value == 0:
        raise TsqError("tsqEncode produced nothing (no CUDA device? see stderr)")
    return out[: n.value].tobytes()


def tsqDecode(stream, ext=0):
    """reference tsqDecode (turbosqueeze.h:670) on one host block."""
    L = library()
    s = np.frombuffer(bytes(stream) + b"\0" * 64, dtype=np.uint8)
    size = int(s[0]) | int(s[1]) << 8 | int(s[2]) << 16
    out = np.zeros(min(size, BLOCK_MAX) + 256, dtype=np.uint8)
    n = C.c_uint32(0)
    L.tsqDecode(s.ctypes.data, out.ctypes.data, C.byref(n), len(stream), int(ext))
    return out[: n.value].tobytes()


def tsq_compress_mt(data, ext=0, level=0):
    """reference tsqAllocateContextCompression_MT + tsqCompress_MT memory->memory + deallocate (test/test.cpp:73-103)."""
    L = library()
    a = _as_np(data)
    ctx = L.tsqAllocateContextCompression_MT(False)
    if not ctx:
        raise TsqError("tsqAllocateContextCompression_MT failed (no CUDA device?)")
    try:
        out, n = _vp(), C.c_size_t(0)
        ok = L.tsqCompress_MT(ctx, a.ctypes.data, a.size, False, C.byref(out), C.byref(n), False, bool(ext), level)
        if not ok:
            return None
        try:
            return C.string_at(out, n.value)
        finally:
            _libc.free(out)
    finally:
        L.tsqDeallocateContextCompression_MT(ctx)


def tsq_decompress_mt(blob):
    L = library()
    a = _as_np(blob)
    ctx = L.tsqAllocateContextDecompression_MT(False)
    if not ctx:
        raise TsqError("tsqAllocateContextDecompression_MT failed (no CUDA device?)")
    try:
        out, n = _vp(), C.c_size_t(0)
        ok = L.tsqDecompress_MT(ctx, a.ctypes.data, a.size, False, C.byref(out), C.byref(n), False)
        if not ok:
            return None
        try:
            return C.string_at(out, n.value)
        finally:
            _libc.free(out)
    finally:
        L.tsqDeallocateContextDecompression_MT(ctx)
