"""ctypes bindings for the parity oracle (oracle/liboracle.so) and, when it was
built, the unmodified compiled reference (oracle/_ref/libtsq_ref.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never by turbosqueeze_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
PAD = 128  # zero bytes after every input buffer

_u8p = C.POINTER(C.c_uint8)
_u32p = C.POINTER(C.c_uint32)


def _ptr(a, t=_u8p):
    return a.ctypes.data_as(t)


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def slot_stride(block):
    """Per-block output slot: worst case + slack, multiple of 128."""
    n = 5 + block + (block >> 4) + ((block + 15) >> 4) + 32
    return (n + 127) // 128 * 128


def padded(data):
    """bytes/ndarray -> uint8 array followed by PAD zero bytes; returns (array, n)."""
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    buf = np.zeros(a.size + PAD, dtype=np.uint8)
    buf[: a.size] = a
    return buf, a.size


class _Codec:
    """Common front end: encode_blocks / decode_blocks on numpy buffers."""

    name = "?"

    def encode_blocks(self, buf, total, block, ext=0, threads=1):
        raise NotImplementedError

    def encode(self, data, block=None, ext=0):
        buf, n = padded(data)
        block = block or max(n, 1)
        slots, sizes, _ = self.encode_blocks(buf, n, block, ext)
        stride = slot_stride(block)
        return [bytes(slots[b * stride: b * stride + int(sizes[b])]) for b in range(len(sizes))]


class Oracle(_Codec):
    name = "oracle-port"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_encode.restype = C.c_uint32
        L.oracle_encode.argtypes = [C.c_void_p, _u8p, C.c_uint32, _u8p, C.c_uint32]
        L.oracle_decode.restype = C.c_uint32
        L.oracle_decode.argtypes = [_u8p, _u8p, C.c_uint32]
        L.oracle_encode_blocks.restype = None
        L.oracle_encode_blocks.argtypes = [_u8p, C.c_uint64, C.c_uint32, _u8p, C.c_uint64, _u32p, C.c_uint32]
        L.oracle_decode_blocks.restype = None
        L.oracle_decode_blocks.argtypes = [_u8p, C.c_uint64, C.c_uint64, _u8p, C.c_uint32, _u32p, C.c_uint32]

    def encode_blocks(self, buf, total, block, ext=0, threads=1):
        nb = (total + block - 1) // block
        stride = slot_stride(block)
        slots = np.zeros(max(nb, 1) * stride, dtype=np.uint8)
        sizes = np.zeros(max(nb, 1), dtype=np.uint32)
        self.lib.oracle_encode_blocks(_ptr(buf), total, block, _ptr(slots), stride, _ptr(sizes, _u32p), ext)
        return slots, sizes[:nb], 0.0

    def decode_blocks(self, slots, stride, nb, block, ext=0):
        out = np.zeros(nb * block + 256, dtype=np.uint8)
        sizes = np.zeros(max(nb, 1), dtype=np.uint32)
        self.lib.oracle_decode_blocks(_ptr(slots), stride, nb, _ptr(out), block, _ptr(sizes, _u32p), ext)
        return out, sizes[:nb]

    def decode_one(self, stream, ext=0):
        s = np.frombuffer(bytes(stream) + b"\0" * 64, dtype=np.uint8)
        size = int(s[0]) | int(s[1]) << 8 | int(s[2]) << 16
        out = np.zeros(min(size, 1 << 22) + 64, dtype=np.uint8)
        n = self.lib.oracle_decode(_ptr(s), _ptr(out), ext)
        return bytes(out[:n])


class Reference(_Codec):
    """The unmodified reference compiled from /root/reference (oracle/Makefile)."""

    name = "reference"

    def __init__(self):
        path = os.path.join(ORACLE_DIR, "_ref", "libtsq_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = L = C.CDLL(path)
        L.ref_encode_blocks.restype = C.c_double
        L.ref_encode_blocks.argtypes = [_u8p, C.c_uint64, C.c_uint32, _u8p, C.c_uint64, _u32p, C.c_uint32, C.c_int, C.c_int]
        L.ref_decode_blocks.restype = C.c_double
        L.ref_decode_blocks.argtypes = [_u8p, C.c_uint64, _u32p, C.c_uint64, _u8p, C.c_uint64, _u32p, C.c_uint32, C.c_int]
        L.ref_compress_mt.restype = C.c_int
        L.ref_compress_mt.argtypes = [_u8p, C.c_uint64, C.POINTER(_u8p), C.POINTER(C.c_uint64), C.c_int]
        L.ref_decompress_mt.restype = C.c_int
        L.ref_decompress_mt.argtypes = [_u8p, C.c_uint64, C.POINTER(_u8p), C.POINTER(C.c_uint64)]
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_hw_threads.restype = C.c_int

    @staticmethod
    def available():
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libtsq_ref.so"))

    def hw_threads(self):
        return int(self.lib.ref_hw_threads())

    def encode_blocks(self, buf, total, block, ext=0, threads=1, slots=None, zero=True):
        nb = (total + block - 1) // block
        stride = slot_stride(block)
        if slots is None:
            slots = np.zeros(max(nb, 1) * stride, dtype=np.uint8)
        sizes = np.zeros(max(nb, 1), dtype=np.uint32)
        secs = self.lib.ref_encode_blocks(_ptr(buf), total, block, _ptr(slots), stride, _ptr(sizes, _u32p), ext,
                                          threads, 1 if zero else 0)
        return slots, sizes[:nb], secs

    def decode_blocks(self, slots, stride, nb, block, ext=0, threads=1, comp_sizes=None, out=None):
        ostride = block + 256
        if out is None:
            out = np.zeros(nb * ostride, dtype=np.uint8)
        sizes = np.zeros(max(nb, 1), dtype=np.uint32)
        cs = _ptr(comp_sizes, _u32p) if comp_sizes is not None else None
        secs = self.lib.ref_decode_blocks(_ptr(slots), stride, cs, nb, _ptr(out), ostride, _ptr(sizes, _u32p), ext, threads)
        return out, sizes[:nb], secs

    def compress_mt(self, data, ext=0):
        a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        src = np.zeros(a.size + 256, dtype=np.uint8)
        src[: a.size] = a
        out, n = _u8p(), C.c_uint64(0)
        ok = self.lib.ref_compress_mt(_ptr(src), a.size, C.byref(out), C.byref(n), ext)
        res = C.string_at(out, n.value) if ok else None
        if out:
            self.lib.ref_free(out)
        return res

    def decompress_mt(self, blob):
        src = np.frombuffer(bytes(blob) + b"\0" * 256, dtype=np.uint8)
        out, n = _u8p(), C.c_uint64(0)
        ok = self.lib.ref_decompress_mt(_ptr(src), len(blob), C.byref(out), C.byref(n))
        res = C.string_at(out, n.value) if ok else None
        if out:
            self.lib.ref_free(out)
        return res


def best_cpu_codec():
    """The compiled reference when it exists, else the C port."""
    return Reference() if Reference.available() else Oracle()
