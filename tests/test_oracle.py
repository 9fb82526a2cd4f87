"""CPU tests: the C restatement (oracle/tsq_oracle.c) is pinned against
(a) golden vectors produced by the compiled, unmodified reference
    (tests/golden/make_golden.py) and
(b) the compiled reference itself when oracle/_ref exists.
"""
import hashlib
import json
import os

import numpy as np
import pytest

from oraclelib import PAD, slot_stride
from turbosqueeze_b200 import workloads as W

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden_blocks.json")))


def golden_input(g):
    n = g["n"]
    if g["input_kind"] == "literal-bytes":
        buf = np.zeros(n + PAD, dtype=np.uint8)
        buf[:n] = np.frombuffer(bytes.fromhex(g["input_hex"]), dtype=np.uint8)
    elif g["input_kind"] == "zeros":
        buf = np.zeros(n + PAD, dtype=np.uint8)
    else:
        buf = W.fill(g["input_kind"], n, seed=g["seed"])
    assert hashlib.sha256(buf[:n].tobytes()).hexdigest() == g["input_sha256"], "workload generator drifted"
    return buf


def check_against_golden(codec, g, ext):
    buf = golden_input(g)
    e = g["ext" if ext else "noext"]
    slots, sizes, _ = codec.encode_blocks(buf, g["n"], g["block"], ext)
    stride = slot_stride(g["block"])
    assert len(sizes) == e["n_blocks"]
    assert [int(s) for s in sizes[:8]] == e["sizes_head"]
    assert int(sizes.sum()) == e["sizes_sum"]
    h = hashlib.sha256()
    for b, s in enumerate(sizes):
        h.update(slots[b * stride: b * stride + int(s)].tobytes())
    if "hex" in e:
        assert slots[: int(sizes[0])].tobytes().hex() == e["hex"]
    assert h.hexdigest() == e["sha256"]
    return buf, slots, sizes


@pytest.mark.parametrize("g", GOLDEN, ids=[g["name"] for g in GOLDEN])
@pytest.mark.parametrize("ext", [0, 1])
def test_oracle_matches_golden_and_round_trips(oracle, g, ext):
    buf, slots, sizes = check_against_golden(oracle, g, ext)
    nb, block = len(sizes), g["block"]
    out, dsz = oracle.decode_blocks(slots, slot_stride(block), nb, block, ext)
    assert int(dsz.sum()) == g["n"]
    assert bytes(out[: g["n"]]) == bytes(buf[: g["n"]])


def test_survey_known_answers(oracle):
    """the two hand-checked vectors."""
    s = oracle.encode(b"abcdefghijklmnopqrstuvwxyz0123456789" * 2 + b"ABCDEFGH")[0]
    assert s.hex() == ("500000" "e3" "ff" + b"abcdefghijklmnop".hex() + b"qrstuvwxyz012345".hex() + "3f" + b"6789".hex()
                       + "2000" "f3" "2400" "1400" "70" + b"ABCDEFGH".hex())
    g = next(x for x in GOLDEN if x["name"] == "testinput_699")
    assert g["noext"]["sizes_sum"] == 570 and g["noext"]["hex"].startswith("bb0200eaff546865")


def test_decode_rejects_oversize_header(oracle):
    # tsq_decode.cpp:53 -- header size > 4 MiB => outputSize 0
    assert oracle.decode_one(bytes([0x01, 0x00, 0x40]) + bytes(32)) == b""
    assert oracle.decode_one(bytes([0x00, 0x00, 0x00]) + bytes(32)) == b""


@pytest.mark.parametrize("kind", ["text", "random", "rep8"])
@pytest.mark.parametrize("ext", [0, 1])
def test_oracle_equals_compiled_reference(oracle, reference, kind, ext):
    rng = np.random.default_rng(5)
    cases = [(1, 1), (5, 5), (100, 100), (4096, 4096), (70000, 70000), (200000, 65536), (1 << 20, 4096),
             (3 << 20, 262144), ((4 << 20) + 1, 1 << 22)]
    cases += [(int(rng.integers(1, 300000)), int(rng.integers(1, 300000))) for _ in range(6)]
    for n, block in cases:
        buf = W.fill(kind, n, seed=n + 3)
        a_slots, a_sizes, _ = oracle.encode_blocks(buf, n, block, ext)
        b_slots, b_sizes, _ = reference.encode_blocks(buf, n, block, ext, threads=4)
        assert np.array_equal(a_sizes, b_sizes), (kind, ext, n, block)
        stride = slot_stride(block)
        for b, s in enumerate(a_sizes):
            assert np.array_equal(a_slots[b * stride: b * stride + s], b_slots[b * stride: b * stride + s]), (n, block, b)
        nb = len(a_sizes)
        out, dsz, _ = reference.decode_blocks(a_slots, stride, nb, block, ext, threads=2)
        got = np.concatenate([out[b * (block + 256): b * (block + 256) + int(dsz[b])] for b in range(nb)])
        assert np.array_equal(got, buf[:n])


def test_tail_bytes_change_the_stream(oracle, reference):
    """bytes after the block are read and matter."""
    n = 50000
    buf = W.fill("text", n + 64, seed=9)
    alone = buf.copy(); alone[n:] = 0
    for codec in (oracle, reference):
        a = codec.encode_blocks(buf, n, n, 0)
        b = codec.encode_blocks(alone, n, n, 0)
        sa, sb = a[0][: a[1][0]].tobytes(), b[0][: b[1][0]].tobytes()
        # both decode to the same thing even when the streams differ
        assert oracle.decode_one(sa) == oracle.decode_one(sb) == buf[:n].tobytes()


def test_reference_container_golden(reference, oracle):
    """TSQ1 framing (turbosqueeze.cpp:64-67): header fields and block walk."""
    g = json.load(open(os.path.join(HERE, "golden", "golden_container.json")))
    for key, ext in (("noext", 0), ("ext", 1)):
        n = g[key]["n"]
        buf = W.fill("text", n, seed=g[key]["seed"])
        blob = reference.compress_mt(buf[:n], ext)
        assert blob[:19].hex()[:32] == g[key]["header_hex"][:32]
        assert blob[:4] == b"TSQ1" and int.from_bytes(blob[4:8], "little") == 2
        assert int.from_bytes(blob[8:16], "little") == n
        at, dec = 16, b""
        while at < len(blob):
            ln = int.from_bytes(blob[at:at + 3], "little"); at += 3
            assert bool(ln & 0x800000) == bool(ext)
            ln &= 0x7FFFFF
            dec += oracle.decode_one(blob[at:at + ln], ext); at += ln
        assert dec == buf[:n].tobytes()
        assert hashlib.sha256(dec).hexdigest() == g[key]["decoded_sha256"]
