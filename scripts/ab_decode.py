"""Development helper: A/B timing of decoder builds in ONE process (same box, same clocks, same buffers): every library under
build/variants/ decodes the same streams, round-robin, several rounds; best and median per library.
usage: python scripts/ab_decode.py <kind> <bytes> <block> [rounds]"""
import glob, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W

kind, n, block = sys.argv[1], int(float(sys.argv[2])), int(sys.argv[3])
rounds = int(sys.argv[4]) if len(sys.argv) > 4 else 5
buf = W.fill(kind, n, seed=20240917)
d = torch.from_numpy(buf).cuda()
ctx0 = T.Context(0)
nb = (n + block - 1) // block
slots, sizes = ctx0.encode_blocks(d, n, block, 0)
out = torch.empty(nb * block, dtype=torch.uint8, device="cuda")
osz = torch.zeros(nb, dtype=torch.int32, device="cuda")
libs = sorted(glob.glob(os.path.join("build", "variants", "lib_*.so")))
ctxs = {os.path.basename(p)[4:-3]: T.Context(0, lib=T.library(os.path.abspath(p))) for p in libs}
times = {k: [] for k in ctxs}
for k, c in ctxs.items():                      # warm-up + correctness
    out.zero_()
    c.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz)
    torch.cuda.synchronize()
    assert torch.equal(out[:n], d[:n]), k
for r in range(rounds):
    for k, c in ctxs.items():
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); c.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz); b.record()
        torch.cuda.synchronize()
        times[k].append(a.elapsed_time(b))
for k in ctxs:
    print(f"{kind} {n} {block}: {k:16s} best {min(times[k]):8.3f} ms  median {statistics.median(times[k]):8.3f} ms  ({n / min(times[k]) / 1e6:.1f} GB/s)", flush=True)
