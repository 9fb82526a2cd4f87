"""Development helper: a small encode + decode of both formats and both table formats, for compute-sanitizer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W
ctx = T.Context(0)
for kind, n, block in [("text", (3 << 20) + 77, 4096), ("text", 700001, 262144), ("rep8", 300000, 65536), ("random", 200000, 200000)]:
    buf = W.fill(kind, n, seed=5)
    d = torch.from_numpy(buf).cuda()
    for ext in (0, 1):
        for fat in (0, 1):
            ctx.set_option("encode_fat", fat)
            slots, sizes = ctx.encode_blocks(d, n, block, ext)
            out, osz = ctx.decode_blocks(slots, sizes.numel(), block, ext, comp_sizes=sizes)
            torch.cuda.synchronize()
            assert torch.equal(out[:n], d[:n]), (kind, n, block, ext, fat)
blob = ctx.compress_buffer(buf[:n], 65536, 0)
assert ctx.decompress_buffer(blob) == buf[:n].tobytes()
print("sanitize_case ok")
