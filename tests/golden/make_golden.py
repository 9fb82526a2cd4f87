"""Generate tests/golden/*.json from the compiled, UNMODIFIED reference
(oracle/_ref/libtsq_ref.so, built by oracle/Makefile from /root/reference).

Run here (the authoring container has /root/reference):  python tests/golden/make_golden.py
The fixtures pin (input recipe -> compressed length + sha256 [+ full hex for the
small ones]) for no-ext and ext streams, under the parity contract of SURVEY.md
8(c): zero-filled output slot, input followed by zero bytes, blocks encoded in
place inside one buffer.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oraclelib import Reference, slot_stride  # noqa: E402
from turbosqueeze_b200 import workloads as W  # noqa: E402

TESTINPUT = ("The names \"John Doe\" for males, \"Jane Doe\" or \"Jane Roe\" for females, or \"Jonnie Doe\" and \"Janie Doe\" "
             "for children, or just \"Doe\" non-gender-specifically are used as placeholder names for a party whose true "
             "identity is unknown or must be withheld in a legal action, case, or discussion. The names are also used to "
             "refer to acorpse or hospital patient whose identity is unknown. This practice is widely used in the United "
             "States and Canada, but is rarely used in other English-speaking countries including the United Kingdom "
             "itself, from where the use of \"John Doe\" in a legal context originates. The names Joe Bloggs or John Smith "
             "are used in the UK instead, as well as in Australia and New Zealand.")  # reference test/test.cpp:26


def recipes():
    """(name, kind-or-bytes, total bytes, block size)"""
    yield "testinput_699", TESTINPUT.encode(), 699, 699
    yield "alnum_80", b"abcdefghijklmnopqrstuvwxyz0123456789" * 2 + b"ABCDEFGH", 80, 80
    for n in (1, 2, 3, 4, 5, 6, 7, 15, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 511, 512, 513, 1000):
        yield f"text_{n}", "text", n, n
    yield "random_1000", "random", 1000, 1000
    yield "rep8_1000", "rep8", 1000, 1000
    yield "zeros_5000", bytes(5000), 5000, 5000
    yield "text_64k_block", "text", 65536, 65536                    # BASELINE.json configs[0]
    yield "text_1m_in_4k", "text", 1 << 20, 4096
    yield "text_1m_in_64k", "text", 1 << 20, 65536
    yield "text_3m_in_256k_ragged", "text", 3 * (1 << 20) + 12345, 262144
    yield "text_9m_in_4m", "text", 9 * (1 << 20) + 77, 1 << 22
    yield "random_1m_in_256k", "random", 1 << 20, 262144
    yield "rep8_2m_in_1m", "rep8", 2 << 20, 1 << 20
    yield "random_200k_one_block", "random", 200000, 200000
    yield "rep8_300k_one_block", "rep8", 300000, 300000


def make_input(src, n):
    if isinstance(src, bytes):
        buf = np.zeros(n + W.PAD, dtype=np.uint8)
        buf[:n] = np.frombuffer(src, dtype=np.uint8)
        return buf
    return W.fill(src, n, seed=1234)


def main():
    ref = Reference()
    out = []
    for name, src, n, block in recipes():
        buf = make_input(src, n)
        entry = {"name": name, "n": n, "block": block, "input_sha256": hashlib.sha256(buf[:n].tobytes()).hexdigest()}
        if isinstance(src, bytes):
            entry["input_hex"] = src.hex() if n <= 1024 else None
            entry["input_kind"] = "literal-bytes" if n <= 1024 else "zeros"
        else:
            entry["input_kind"] = src
            entry["seed"] = 1234
        for ext in (0, 1):
            slots, sizes, _ = ref.encode_blocks(buf, n, block, ext, threads=1)
            stride = slot_stride(block)
            h = hashlib.sha256()
            for b, s in enumerate(sizes):
                h.update(slots[b * stride: b * stride + int(s)].tobytes())
            e = {"sizes_sum": int(sizes.sum()), "n_blocks": len(sizes), "sizes_head": [int(s) for s in sizes[:8]],
                 "sha256": h.hexdigest()}
            if n <= 1024:
                e["hex"] = slots[: int(sizes[0])].tobytes().hex()
            entry["ext" if ext else "noext"] = e
        out.append(entry)
        print(name, entry["noext"]["sizes_sum"], entry["ext"]["sizes_sum"])
    with open(os.path.join(HERE, "golden_blocks.json"), "w") as f:
        json.dump(out, f, indent=1)
    # one TSQ1 container produced by the reference's own MT pipeline (4 MiB blocks)
    n = 5 * (1 << 20) + 321
    buf = W.fill("text", n, seed=77)
    cont = {}
    for ext in (0, 1):
        blob = ref.compress_mt(buf[:n], ext)
        cont["ext" if ext else "noext"] = {"n": n, "seed": 77, "len": len(blob), "header_hex": blob[:19].hex(),
                                           "sha256_masked_note": "MT output recycles buffers; only decode result is pinned",
                                           "decoded_sha256": hashlib.sha256(ref.decompress_mt(blob)).hexdigest()}
    with open(os.path.join(HERE, "golden_container.json"), "w") as f:
        json.dump(cont, f, indent=1)


if __name__ == "__main__":
    main()
