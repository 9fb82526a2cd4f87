"""Development helper: small encodes + decodes of both formats and both table formats, the chunked and piece-streamed host
paths and the decoder's 27..30-slot launch, for compute-sanitizer (memcheck: minutes; sizes are small on purpose)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W
ctx = T.Context(0)


def gen(kind, n, seed):
    if kind == "zeros":
        return np.zeros(n + W.PAD, dtype=np.uint8)
    if kind == "runs":                                   # runs of one byte, 1 .. 300 long: overlapping matches of every distance
        rng = np.random.default_rng(seed)
        lens = rng.integers(1, 300, size=n // 100 + 2)
        vals = rng.integers(0, 256, size=lens.size, dtype=np.uint8)
        buf = np.repeat(vals, lens)[:n]
        return np.concatenate([buf, np.zeros(n + W.PAD - buf.size, dtype=np.uint8)])
    return W.fill(kind, n, seed=seed)


for kind, n, block in [("text", (3 << 20) + 77, 4096), ("text", 700001, 262144), ("rep8", 300000, 65536), ("random", 200000, 200000),
                       ("random", 300007, 65536), ("runs", 400000, 65536), ("zeros", 100000, 32768)]:
    buf = gen(kind, n, 5)
    d = torch.from_numpy(buf).cuda()
    for ext in (0, 1):
        for fat in (0, 1):
            ctx.set_option("encode_fat", fat)
            slots, sizes = ctx.encode_blocks(d, n, block, ext)
            out, osz = ctx.decode_blocks(slots, sizes.numel(), block, ext, comp_sizes=sizes)
            torch.cuda.synchronize()
            assert torch.equal(out[:n], d[:n]), (kind, n, block, ext, fat)
ctx.set_option("encode_fat", 1)
# 28 blocks per SM at 64 KiB blocks: the one-round launch above the shared-memory carve-out (text: lane-per-pair copier, random: dense)
sm = torch.cuda.get_device_properties(0).multi_processor_count
for kind in ("text", "random"):
    n = (27 * sm + 5) * 65536 + 123
    buf = W.fill(kind, n, seed=6)
    d = torch.from_numpy(buf).cuda()
    slots, sizes = ctx.encode_blocks(d, n, 65536, 0)
    out, osz = ctx.decode_blocks(slots, sizes.numel(), 65536, 0, comp_sizes=sizes)
    torch.cuda.synchronize()
    assert torch.equal(out[:n], d[:n]), kind
# host paths: one-shot, chunked, piece-streamed
ctx.set_option("pipeline_min", 1 << 20)
for kind, n, block, ext in [("text", (5 << 20) + 4321, 65536, 0), ("text", (24 << 20) + 7, 262144, 1), ("random", (3 << 20) + 99, 262144, 0)]:
    buf = W.fill(kind, n, seed=7)
    ctx.set_option("pipeline", 0)
    one = ctx.compress_buffer(buf[:n], block, ext)
    ctx.set_option("pipeline", 1)
    for stream_in in (0, 1):
        ctx.set_option("stream_in", stream_in)
        assert ctx.compress_buffer(buf[:n], block, ext) == one, (kind, stream_in)
    assert ctx.decompress_buffer(one) == buf[:n].tobytes()
print("sanitize_case ok")
