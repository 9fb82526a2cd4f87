"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C-ABI of
libturbosqueeze_b200.so; the oracle (tests/oraclelib.py) is only the checker.

Bar: bit-exact.  encode: every byte of every block's stream equals the reference's (oracle/_ref when
it travelled with the repo, else the pinned C restatement) under the contract of SURVEY.md 8(c);
decode: output equals the original input.
"""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oraclelib import PAD, Reference, slot_stride
from turbosqueeze_b200 import workloads as W

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_blocks.json")))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def ctx(torch):
    import turbosqueeze_b200 as T
    c = T.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def xctx(torch):
    """Context of the test-only cross-check library (product objects + the superseded round-1 kernels: encode_impl 2,
    decode_lanes 1..33).  The product library itself rejects those variants."""
    import turbosqueeze_b200 as T
    from turbosqueeze_b200 import api
    c = T.Context(0, lib=T.library(api.xcheck_library_path()))
    yield c
    c.close()


@pytest.fixture(scope="module")
def checker(oracle):
    """The compiled unmodified reference when present (multi-threaded), else the C restatement."""
    return Reference() if Reference.available() else oracle


def gpu_encode(torch, ctx, buf, n, block, ext, impl):
    import turbosqueeze_b200 as T
    assert T.slot_stride(block) == slot_stride(block)
    ctx.set_option("encode_impl", impl)
    d = torch.from_numpy(buf).cuda()
    slots, sizes = ctx.encode_blocks(d, n, block, ext)
    torch.cuda.synchronize()
    return slots.cpu().numpy(), sizes.cpu().numpy().astype(np.uint32)


def assert_streams_equal(a_slots, a_sizes, b_slots, b_sizes, block, what):
    assert np.array_equal(a_sizes, b_sizes), (what, "sizes differ", np.flatnonzero(a_sizes != b_sizes)[:5])
    stride = slot_stride(block)
    nb = len(a_sizes)
    if nb * stride <= (1 << 28):
        idx = np.arange(stride, dtype=np.int64)[None, :] < a_sizes.astype(np.int64)[:, None]
        A = a_slots[: nb * stride].reshape(nb, stride)
        B = b_slots[: nb * stride].reshape(nb, stride)
        bad = np.flatnonzero(((A != B) & idx).any(axis=1))
        assert bad.size == 0, (what, "blocks differ", bad[:5])
    else:
        for b in range(nb):
            s = int(a_sizes[b])
            assert np.array_equal(a_slots[b * stride: b * stride + s], b_slots[b * stride: b * stride + s]), (what, "block", b)


def golden_input(g):
    n = g["n"]
    if g["input_kind"] == "literal-bytes":
        buf = np.zeros(n + PAD, dtype=np.uint8)
        buf[:n] = np.frombuffer(bytes.fromhex(g["input_hex"]), dtype=np.uint8)
    elif g["input_kind"] == "zeros":
        buf = np.zeros(n + PAD, dtype=np.uint8)
    else:
        buf = W.fill(g["input_kind"], n, seed=g["seed"])
    return buf


@pytest.mark.parametrize("impl,ext", [(3, 0), (3, 1), (2, 0), (1, 0), (1, 1)], ids=["batch", "batch-ext", "warp", "scalar", "scalar-ext"])
def test_encode_matches_reference_golden_vectors(torch, ctx, xctx, impl, ext):
    """tests/golden/golden_blocks.json was produced by the compiled unmodified reference."""
    if impl == 2:
        ctx = xctx
    for g in GOLDEN:
        buf = golden_input(g)
        e = g["ext" if ext else "noext"]
        slots, sizes = gpu_encode(torch, ctx, buf, g["n"], g["block"], ext, impl)
        stride = slot_stride(g["block"])
        assert len(sizes) == e["n_blocks"], g["name"]
        assert [int(s) for s in sizes[:8]] == e["sizes_head"], g["name"]
        assert int(sizes.sum()) == e["sizes_sum"], g["name"]
        h = hashlib.sha256()
        for b, s in enumerate(sizes):
            h.update(slots[b * stride: b * stride + int(s)].tobytes())
        if "hex" in e:
            assert slots[: int(sizes[0])].tobytes().hex() == e["hex"], g["name"]
        assert h.hexdigest() == e["sha256"], g["name"]


CASES = [(1, 1), (2, 2), (5, 5), (31, 31), (32, 32), (33, 33), (34, 34), (63, 64), (100, 100), (4096, 4096), (65535, 65535),
         (65536, 65536), (65537, 65537), (70000, 70000), (200000, 65536), (200001, 66667), (1 << 20, 4096), (1 << 20, 1000), (3 << 20, 262144),
         ((3 << 20) + 12345, 262144), ((4 << 20) + 1, 1 << 22), (600000, 131072 + 7)]


@pytest.mark.parametrize("kind", ["text", "random", "rep8", "zeros", "runs"])
@pytest.mark.parametrize("impl,ext", [(3, 0), (3, 1), (2, 0), (1, 0), (1, 1)], ids=["batch", "batch-ext", "warp", "scalar", "scalar-ext"])
def test_encode_bit_exact_vs_oracle(torch, ctx, xctx, checker, kind, impl, ext):
    if impl == 2:
        ctx = xctx
    rng = np.random.default_rng(11)
    cases = CASES + [(int(rng.integers(1, 400000)), int(rng.integers(1, 300000))) for _ in range(6)]
    for n, block in cases:
        buf = make_input(kind, n, seed=n + 3)
        want_slots, want_sizes, _ = checker.encode_blocks(buf, n, block, ext)
        got_slots, got_sizes = gpu_encode(torch, ctx, buf, n, block, ext, impl)
        assert_streams_equal(got_slots, got_sizes, want_slots, want_sizes, block, (kind, n, block, impl, ext))


def make_input(kind, n, seed):
    if kind == "zeros":
        return np.zeros(n + PAD, dtype=np.uint8)
    if kind == "runs":   # short runs and near repeats: stresses same-hash positions inside one 32-wide window
        rng = np.random.default_rng(seed)
        sym = rng.integers(97, 101, size=n // 3 + 2, dtype=np.uint8)
        a = np.repeat(sym, rng.integers(1, 9, size=sym.size))[:n]
        buf = np.zeros(n + PAD, dtype=np.uint8)
        buf[: a.size] = a
        if a.size < n:
            buf[a.size:n] = 120
        return buf
    return W.fill(kind, n, seed=seed)


def test_decode_mixed_blocks_switch_copier_mode(torch, ctx, oracle):
    """The default decoder picks its copier per block (lane per pair, or lane per symbol for incompressible blocks).
    8192 blocks of 4 KiB (more than one round per block slot), text and random blocks interleaved irregularly, so
    that every slot switches mode back and forth; every byte must come back."""
    block, nb = 4096, 8192
    n = block * nb - 1234
    text = make_input("text", n, seed=77)
    rnd = make_input("random", n, seed=78)
    buf = text.copy()
    pick = (np.arange(nb, dtype=np.uint64) * np.uint64(2654435761) >> np.uint64(7)) & np.uint64(3)
    for b in np.nonzero(pick == 0)[0]:
        lo, hi = int(b) * block, min((int(b) + 1) * block, n)
        buf[lo:hi] = rnd[lo:hi]
    slots, sizes, _ = oracle.encode_blocks(buf, n, block, 0)
    assert (sizes > block).any() and (sizes < block * 0.9).any()
    d_slots = torch.from_numpy(slots).cuda()
    d_sizes = torch.from_numpy(sizes.astype(np.int32)).cuda()
    for lanes in (0, 35, 34):
        ctx.set_option("decode_lanes", lanes)
        try:
            out, osz = ctx.decode_blocks(d_slots, nb, block, 0, comp_sizes=d_sizes)
            torch.cuda.synchronize()
        finally:
            ctx.set_option("decode_lanes", 0)
        assert int(osz.cpu().numpy().sum()) == n
        assert np.array_equal(out.cpu().numpy()[:n], buf[:n]), lanes


@pytest.mark.parametrize("lanes", [35, 34, 33, 32, 16, 8, 4, 2, 1])
@pytest.mark.parametrize("ext", [0, 1])
def test_decode_restores_input(torch, ctx, xctx, checker, lanes, ext):
    """lanes 34 = walker + copier kernel (tsq_decode_split.cu, lane per symbol), 35 = the same with the lane-per-pair
    copier (64 symbols per step; the extension format runs as 34), 33 = warp-per-block step kernel
    (tsq_decode_warp.cu), 1..32 = sub-warp pair-step kernel."""
    if lanes == 33 and ext:
        pytest.skip("the v1 step kernel is no-extension only")
    if lanes <= 33:
        ctx = xctx
    oracle = checker                      # streams come from the compiled unmodified reference when it travelled with the repo
    ctx.set_option("decode_lanes", lanes)
    try:
        for kind in ("text", "random", "rep8", "zeros", "runs"):
            for n, block in [(1, 1), (5, 5), (100, 100), (4096, 4096), (70000, 70000), (200000, 65536), (200001, 66667), (1 << 20, 4096),
                             ((3 << 20) + 12345, 262144), ((4 << 20) + 1, 1 << 22)]:
                buf = make_input(kind, n, seed=n + 5)
                slots, sizes, _ = oracle.encode_blocks(buf, n, block, ext)
                nb = len(sizes)
                d_slots = torch.from_numpy(slots).cuda()
                d_sizes = torch.from_numpy(sizes.astype(np.int32)).cuda()
                out, osz = ctx.decode_blocks(d_slots, nb, block, ext, comp_sizes=d_sizes)
                torch.cuda.synchronize()
                osz = osz.cpu().numpy()
                assert int(osz.sum()) == n, (kind, n, block)
                assert np.array_equal(out.cpu().numpy()[:n], buf[:n]), (kind, n, block, lanes, ext)
    finally:
        ctx.set_option("decode_lanes", 0)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("ext", [0, 1])
def test_decode_of_garbage_streams_terminates(torch, ctx, ext):
    """The reference does not validate streams (tsq_decode.cpp:42-126 reads whatever the bytes say).  The device
    decoder must at least stay inside its buffers and terminate: random bytes behind a plausible header, truncated
    real streams, and offsets pointing before the block."""
    rng = np.random.default_rng(1234 + ext)
    nb, block, stride = 96, 65536, 8192
    comp = rng.integers(0, 256, size=nb * stride + 64, dtype=np.uint8)     # 16 readable bytes behind the last stream
    for b in range(nb):
        size = int(rng.integers(0, block + 1))
        comp[b * stride: b * stride + 3] = [size & 0xFF, (size >> 8) & 0xFF, size >> 16]
    sizes = rng.integers(0, stride, size=nb).astype(np.int32)
    d_comp = torch.from_numpy(comp).cuda()
    d_sizes = torch.from_numpy(sizes).cuda()
    out, osz = ctx.decode_blocks(d_comp, nb, block, ext, stride=stride, comp_sizes=d_sizes)
    torch.cuda.synchronize()
    assert out.numel() == nb * block and int(osz.max().item()) <= block
    # a real stream cut short
    buf = W.fill("text", 200000, seed=3)
    from oraclelib import Oracle
    slots, csz, _ = Oracle().encode_blocks(buf, 200000, 65536, ext)
    d = torch.from_numpy(slots).cuda()
    cut = torch.from_numpy((csz // 2).astype(np.int32)).cuda()
    out, osz = ctx.decode_blocks(d, len(csz), 65536, ext, comp_sizes=cut)
    torch.cuda.synchronize()
    assert int(osz.sum().item()) == 200000          # the header is reported as it is; the tail of each block is unspecified


def test_product_library_has_no_superseded_kernels(torch, ctx):
    """encode_impl 2 and decode_lanes 1..33 exist only in the cross-check library."""
    import turbosqueeze_b200 as T
    buf = W.fill("text", 70000, seed=1)
    d = torch.from_numpy(buf).cuda()
    slots, sizes = ctx.encode_blocks(d, 70000, 65536, 0)
    ctx.set_option("decode_lanes", 16)
    try:
        with pytest.raises(T.TsqError):
            ctx.decode_blocks(slots, sizes.numel(), 65536, 0, comp_sizes=sizes)
    finally:
        ctx.set_option("decode_lanes", 0)


def test_decode_rejects_oversize_header(torch, ctx):
    """tsq_decode.cpp:53 -- header size > 4 MiB => outputSize 0 (and nothing written)."""
    import turbosqueeze_b200 as T
    assert T.tsqDecode(bytes([0x01, 0x00, 0x40]) + bytes(32)) == b""
    assert T.tsqDecode(bytes([0x00, 0x00, 0x00]) + bytes(32)) == b""


def test_reference_block_api_round_trip(torch, checker):
    """Mirror of the reference's test_tsq_compress (test/test.cpp:30-54) through tsqEncode/tsqDecode,
    plus byte equality with the reference for both formats and for a non-zero pre-filled output."""
    import turbosqueeze_b200 as T
    g = next(x for x in GOLDEN if x["name"] == "testinput_699")
    data = bytes.fromhex(g["input_hex"])
    for ext in (0, 1):
        stream = T.tsqEncode(data, ext)
        assert stream.hex() == g["ext" if ext else "noext"]["hex"]
        assert T.tsqDecode(stream, ext) == data
    # trailing never-initialised bytes follow the caller's pre-fill
    if isinstance(checker, Reference):
        rng = np.random.default_rng(3)
        for n in (16, 64, 128, 1000, 4096, 65536):
            raw = rng.integers(0, 256, size=n, dtype=np.uint8)
            buf = np.zeros(n + PAD, dtype=np.uint8); buf[:n] = raw
            for fill in (0x00, 0xAB):
                pre = np.full(slot_stride(n), fill, dtype=np.uint8)
                slots, sizes, _ = checker.encode_blocks(buf, n, n, 0, slots=pre, zero=False)
                assert T.tsqEncode(raw, 0, prefill=fill) == slots[: int(sizes[0])].tobytes(), (n, fill)


def test_tail_bytes_are_read_like_the_reference(torch, ctx, checker):
    """the bytes after a block influence its stream."""
    import turbosqueeze_b200 as T
    n = 50000
    buf = W.fill("text", n + 64, seed=9)
    follow = buf.copy(); follow[n + 64:] = 0
    want_slots, want_sizes, _ = checker.encode_blocks(follow, n, n, 0)
    got = T.tsqEncode(buf[:n], 0, tail=buf[n:n + 64])
    assert got == want_slots[: int(want_sizes[0])].tobytes()


def test_container_pack_index_and_buffer_api(torch, ctx, oracle):
    """TSQ1 framing on the device + the synchronous buffer API (tsq_threads.cpp:413-441,862-890)."""
    import turbosqueeze_b200 as T
    for ext in (0, 1):
        for n, block in [(1, 4096), (5000, 4096), ((5 << 20) + 321, 1 << 22), (1 << 20, 65536)]:
            buf = W.fill("text", n, seed=77)
            blob = ctx.compress_buffer(buf[:n], block, ext)
            nb = (n + block - 1) // block
            assert blob[:4] == b"TSQ1" and int.from_bytes(blob[4:8], "little") == nb
            assert int.from_bytes(blob[8:16], "little") == n
            # walk the container on the host: every block equals the oracle's stream
            slots, sizes, _ = oracle.encode_blocks(buf, n, block, ext)
            stride = slot_stride(block)
            at = 16
            for b in range(nb):
                ln = int.from_bytes(blob[at:at + 3], "little"); at += 3
                assert bool(ln & 0x800000) == bool(ext)
                ln &= 0x7FFFFF
                assert ln == int(sizes[b])
                assert blob[at:at + ln] == slots[b * stride: b * stride + ln].tobytes()
                at += ln
            assert at == len(blob)
            assert ctx.decompress_buffer(blob) == buf[:n].tobytes()
    # reference-style MT buffer API, memory -> memory (test/test.cpp:149-199)
    n = (9 << 20) + 77
    buf = W.fill("text", n, seed=5)
    for ext in (0, 1):
        blob = T.tsq_compress_mt(buf[:n], ext)
        assert blob is not None and T.tsq_decompress_mt(blob) == buf[:n].tobytes()
    if Reference.available():
        ref = Reference()
        blob = T.tsq_compress_mt(buf[:n], 0)
        assert ref.decompress_mt(blob) == buf[:n].tobytes()           # the reference reads our container
        assert T.tsq_decompress_mt(ref.compress_mt(buf[:n], 1)) == buf[:n].tobytes()   # and we read its
        # smaller container blocks (more blocks in flight) stay readable by the reference's decoder
        L = T.library()
        assert L.tsqb_set_container_block_size(262144) == 0
        try:
            small = T.tsq_compress_mt(buf[:n], 0)
            assert int.from_bytes(small[4:8], "little") == (n + 262143) // 262144
            assert ref.decompress_mt(small) == buf[:n].tobytes() and T.tsq_decompress_mt(small) == buf[:n].tobytes()
        finally:
            assert L.tsqb_set_container_block_size(1 << 22) == 0


@pytest.mark.parametrize("fat", [0, 1], ids=["u16-tables", "sector-entries"])
def test_batch_encoder_table_formats_are_bit_exact(torch, ctx, checker, fat):
    """The batch encoder keeps the reference's 2^17 x u16 table when few blocks are in flight (L2-resident) and
    32-byte sector entries (exact word tag + 12 bytes + epoch, no zeroing) otherwise: same bytes either way."""
    ctx.set_option("encode_fat", fat)
    try:
        for kind, n, block in [("text", (3 << 20) + 777, 4096), ("text", (2 << 20) + 5, 262144), ("rep8", 1 << 20, 65536),
                               ("random", (1 << 20) + 3, 16384), ("runs", 300000, 300000), ("text", 70000, 70000),
                               ("text", (3 << 20) + 11, 1 << 20), ("text", (4 << 20) + 100, 1 << 22)]:   # entries aged by > 3 segments
            buf = make_input(kind, n, seed=n + 11)
            want_slots, want_sizes, _ = checker.encode_blocks(buf, n, block, 0)
            for rep in range(2):                                   # second pass: tables hold entries of older epochs
                got_slots, got_sizes = gpu_encode(torch, ctx, buf, n, block, 0, 3)
                assert_streams_equal(got_slots, got_sizes, want_slots, want_sizes, block, (kind, n, block, fat, rep))
    finally:
        ctx.set_option("encode_fat", -1)
        ctx.set_option("encode_impl", 0)


def test_cpp_caller_of_the_reference_api(torch, tmp_path):
    """A C++ program written against the reference's API (block calls, buffer API, async jobs chained from
    callbacks, drain on destroy, FILE* entry points -- modelled on the reference's test/test.cpp) links
    libturbosqueeze_b200.so and passes."""
    import subprocess
    import turbosqueeze_b200 as T
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "async_harness")
    libdir = os.path.dirname(T.library_path())
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(root, "include"), os.path.join(HERE, "async_harness.cpp"),
                    "-L" + libdir, "-lturbosqueeze_b200", "-Wl,-rpath," + libdir, "-lpthread", "-o", exe], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "async_harness ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-1500:])


def test_pipelined_host_path_equals_one_shot(torch, ctx):
    """The stream-overlapped host paths must produce the very same container as one-shot staging, including the bytes a
    chunk's last block reads from the next chunk: whole chunks on their own streams (H2D | kernels | D2H), and the
    piece-streamed path, where the encoder starts on a block's first piece and polls for the rest while it runs
    (tsq_capi.cu compress_streamed; 64 KiB ... 4 MiB blocks, ragged last block, both formats)."""
    ctx.set_option("pipeline_min", 1 << 20)
    try:
        for kind, n, block, ext in [("text", (5 << 20) + 4321, 65536, 0), ("text", (3 << 20), 4096, 0), ("rep8", (6 << 20) + 1, 1 << 20, 0),
                                    ("random", (2 << 20) + 99, 262144, 0), ("text", (40 << 20) + 4321, 262144, 0), ("text", (48 << 20) + 7, 262144, 1),
                                    ("text", 64 << 20, 1 << 20, 0), ("random", (24 << 20) + 100000, 262144, 0), ("rep8", (50 << 20) + 5, 1 << 20, 1),
                                    ("text", (200 << 20) + 1, 4 << 20, 0)]:
            buf = W.fill(kind, n, seed=31)
            ctx.set_option("pipeline", 0)
            one = ctx.compress_buffer(buf[:n], block, ext)
            ctx.set_option("pipeline", 1)
            for stream_in in (0, 1):
                ctx.set_option("stream_in", stream_in)
                piped = ctx.compress_buffer(buf[:n], block, ext)
                if piped != one:
                    a, b = np.frombuffer(one, np.uint8), np.frombuffer(piped, np.uint8)
                    m = min(a.size, b.size)
                    bad = np.flatnonzero(a[:m] != b[:m])
                    raise AssertionError((kind, n, block, ext, stream_in, a.size, b.size, bad[:8].tolist()))
            assert ctx.decompress_buffer(piped) == buf[:n].tobytes(), (kind, n, block)
            ctx.set_option("pipeline", 0)
            assert ctx.decompress_buffer(piped) == buf[:n].tobytes()
    finally:
        ctx.set_option("pipeline", 1)
        ctx.set_option("stream_in", 1)
        ctx.set_option("pipeline_min", 64 << 20)


@pytest.mark.parametrize("kind,block,ext", [("text", 262144, 0), ("random", 262144, 0), ("rep8", 1 << 20, 0), ("text", 4096, 0),
                                            ("text", 262144, 1), ("rep8", 65536, 1),
                                            ("text", 65536, 0), ("random", 65536, 0)])     # 4097 blocks: 28 per SM, one decode round above the L1 carve-out
def test_large_buffers_bit_exact_and_round_trip(torch, ctx, checker, kind, block, ext):
    """BASELINE.json shapes at a size the multi-threaded reference finishes in seconds (256 MiB):
    every stream byte equals the reference's, and decode(encode(x)) == x on the device."""
    n = (256 << 20) + 54321
    buf = W.fill(kind, n, seed=2024)
    ctx.set_option("encode_impl", 0)
    d = torch.from_numpy(buf).cuda()
    slots, sizes = ctx.encode_blocks(d, n, block, ext)
    out, osz = ctx.decode_blocks(slots, sizes.numel(), block, ext, comp_sizes=sizes)
    torch.cuda.synchronize()
    assert int(osz.sum().item()) == n
    assert torch.equal(out[:n], d[:n])
    threads = os.cpu_count() or 8
    want_slots, want_sizes, _ = checker.encode_blocks(buf, n, block, ext, threads=threads)
    got_sizes = sizes.cpu().numpy().astype(np.uint32)
    assert np.array_equal(got_sizes, want_sizes)
    assert_streams_equal(slots.cpu().numpy(), got_sizes, want_slots, want_sizes, block, (kind, block, ext))
