"""Development helper: decode time against blocks per SM and the "decode_slots" option (one process, CUDA events)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]

def main():
    block = int(sys.argv[5]) if len(sys.argv) > 5 else 262144
    nmax = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1600000000
    counts = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [3815, 3848, 4096, 4440, 5000, 6000]
    slot_opts = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 26, 28, 30]
    ctx = T.Context(0)
    kind = sys.argv[4] if len(sys.argv) > 4 else "text"
    buf = W.fill(kind, nmax, seed=20240917)
    d = torch.from_numpy(buf).cuda()
    nb_all = nmax // block
    stride = T.slot_stride(block)
    slots = torch.zeros(nb_all * stride, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(nb_all, dtype=torch.int32, device="cuda")
    out = torch.empty(nb_all * block, dtype=torch.uint8, device="cuda")
    osz = torch.zeros(nb_all, dtype=torch.int32, device="cuda")
    ctx.encode_blocks(d, nb_all * block, block, 0, slots=slots, sizes=sizes)
    torch.cuda.synchronize()
    for nb in counts:
        if nb > nb_all:
            continue
        n = nb * block
        for s in slot_opts:
            ctx.set_option("decode_slots", s)
            out.zero_()
            t = timeit(lambda: ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz))
            ok = torch.equal(out[:n], d[:n])
            print(f"{kind} nb={nb} per_sm={-(-nb // 148)} decode_slots={s}: {t:.3f} ms {n / t / 1e6:.1f} GB/s ok={ok}", flush=True)

main()
