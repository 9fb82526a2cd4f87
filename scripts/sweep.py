"""Development helper: sweep kernel knobs on the bench workload (CUDA events, one GPU)."""
import sys, os
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
    kind = sys.argv[1] if len(sys.argv) > 1 else "text"
    n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10**9
    block = int(sys.argv[3]) if len(sys.argv) > 3 else 262144
    enc_slots = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0]
    dec_lanes = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [0]
    ctx = T.Context(0)
    if os.environ.get("ENC_HINTS"):
        ctx.set_option("encode_hints", int(os.environ["ENC_HINTS"]))
    if os.environ.get("DEC_SLOTS"):
        ctx.set_option("decode_slots", int(os.environ["DEC_SLOTS"]))
    if os.environ.get("ENC_FAT"):
        ctx.set_option("encode_fat", int(os.environ["ENC_FAT"]))
    if os.environ.get("L2_FETCH"):
        ctx.set_option("l2_fetch", int(os.environ["L2_FETCH"]))
    buf = W.fill(kind, n, seed=20240917)
    d = torch.from_numpy(buf).cuda()
    assert d.numel() >= n + T.INPUT_PAD
    nb = (n + block - 1) // block
    stride = T.slot_stride(block)
    slots = torch.zeros(nb * stride, dtype=torch.uint8, device="cuda")
    sizes = torch.zeros(nb, dtype=torch.int32, device="cuda")
    out = torch.empty(nb * block, dtype=torch.uint8, device="cuda")
    osz = torch.zeros(nb, dtype=torch.int32, device="cuda")
    for s in enc_slots:
        ctx.set_option("encode_slots", s)
        t = timeit(lambda: ctx.encode_blocks(d, n, block, 0, slots=slots, sizes=sizes), reps=2)
        # checksum of every stream byte (slots are zero outside the streams): equal checksums <=> equal output for A/B runs
        chk = int(slots.view(torch.int64).sum().item()) & 0xFFFFFFFFFFFF
        print(f"{kind} {block} nb={nb} encode slots={s}: {t:.3f} ms {n/t/1e6:.1f} GB/s ratio={int(sizes.sum().item())/n:.4f} chk={chk:012x}", flush=True)
    ctx.set_option("encode_slots", 0)
    for lanes in dec_lanes:
        ctx.set_option("decode_lanes", lanes)
        out.zero_()
        t = timeit(lambda: ctx.decode_blocks(slots, nb, block, 0, comp_sizes=sizes, out=out, out_sizes=osz))
        ok = torch.equal(out[:n], d[:n])
        print(f"{kind} {block} nb={nb} decode lanes={lanes}: {t:.3f} ms {n/t/1e6:.1f} GB/s ok={ok}", flush=True)

main()
