"""Development helper: time encode/decode kernels on one GPU for a few shapes (CUDA events)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best

def main():
    mb = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    ctx = T.Context(0)
    for kind, block in [("text", 262144), ("random", 262144), ("rep8", 1 << 20), ("text", 4096), ("text", 65536), ("text", 1 << 22)]:
        n = mb << 20
        buf = W.fill(kind, n, seed=1)
        d = torch.from_numpy(buf).cuda()
        nb = (n + block - 1) // block
        stride = T.slot_stride(block)
        slots = torch.zeros(nb * stride, dtype=torch.uint8, device="cuda")
        sizes = torch.zeros(nb, dtype=torch.int32, device="cuda")
        out = torch.empty(nb * block, dtype=torch.uint8, device="cuda")
        osz = torch.zeros(nb, dtype=torch.int32, device="cuda")
        res = {}
        for impl in (2,):
            if impl == 1 and nb > 20000: continue
            ctx.set_option("encode_impl", impl)
            for slots_opt in (0,):
                ctx.set_option("encode_slots", slots_opt)
                t = timeit(lambda: ctx.encode_blocks(d, n, block, 0, slots=slots, sizes=sizes), reps=2)
                res[f"enc{impl}/s{slots_opt}"] = n / t / 1e6
        ctx.set_option("encode_impl", 0); ctx.set_option("encode_slots", 0)
        c = int(sizes.sum().item())
        for lanes in (33, 32, 8, 4):
            ctx.set_option("decode_lanes", lanes)
            t = timeit(lambda: ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz))
            res[f"dec/w{lanes}"] = n / t / 1e6
        ctx.set_option("decode_lanes", 0)
        ok = torch.equal(out[:n], d[:n])
        print(kind, block, f"ratio={c/n:.4f} ok={ok}", " ".join(f"{k}={v:.1f}GB/s" for k, v in res.items()), flush=True)

main()
