"""Development helper: time the host-buffer calls (pinned memory) separately."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W
n = 10**9; block = 262144
buf = W.fill("text", n, seed=20240917)
ctx = T.Context(0)
if len(sys.argv) > 1: ctx.set_option("pipe_chunks", int(sys.argv[1]))
if len(sys.argv) > 2: ctx.set_option("pipe_taper", int(sys.argv[2]))
if len(sys.argv) > 3: ctx.set_option("stream_in", int(sys.argv[3]))
if len(sys.argv) > 4: ctx.set_option("stream_stagger", int(sys.argv[4]))
pin = torch.empty(n + 128, dtype=torch.uint8, pin_memory=True); pin.numpy()[:] = buf
nb = (n + block - 1)//block; stride = T.slot_stride(block); cap = 16 + nb*(stride+3)
pc = torch.empty(cap, dtype=torch.uint8, pin_memory=True); po = torch.empty(n+128, dtype=torch.uint8, pin_memory=True)
for rep in range(4):
    t0 = time.perf_counter(); c = ctx.compress_into(pin.data_ptr(), n, block, 0, pc.data_ptr(), cap); t1 = time.perf_counter()
    m = ctx.decompress_into(pc.data_ptr(), c, po.data_ptr(), n+128); t2 = time.perf_counter()
    print(f"compress {1e3*(t1-t0):.1f} ms  decompress {1e3*(t2-t1):.1f} ms  ok={m==n and bool((po.numpy()[:n]==buf[:n]).all())}", flush=True)
