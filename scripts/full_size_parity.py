"""Bit-exact parity at BASELINE.json's full single-GPU sizes: GPU streams vs the unmodified reference (oracle/_ref),
block by block, plus decode(encode(x)) == x on the device.  python scripts/full_size_parity.py  (GPU box)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W
from oraclelib import Reference, slot_stride

ref = Reference()
ctx = T.Context(0)
for name, kind, n, block in [("cfg2 enwik9-shape 1 GB / 256 KiB", "text", 10**9, 262144),
                             ("cfg3 uniform random 4 GiB / 256 KiB", "random", 4 << 30, 262144),
                             ("cfg4 8-byte period 2 GiB (one GPU's share of 16 GiB) / 1 MiB", "rep8", 2 << 30, 1 << 20)]:
    buf = W.fill(kind, n, seed=20240917)
    d = torch.from_numpy(buf).cuda()
    slots, sizes = ctx.encode_blocks(d, n, block, 0)
    out, osz = ctx.decode_blocks(slots, sizes.numel(), block, 0, comp_sizes=sizes)
    torch.cuda.synchronize()
    rt = bool(torch.equal(out[:n], d[:n])) and int(osz.sum().item()) == n
    t0 = time.time()
    want_slots, want_sizes, _ = ref.encode_blocks(buf, n, block, 0, threads=os.cpu_count())
    got_sizes = sizes.cpu().numpy().astype(np.uint32)
    same_sizes = bool(np.array_equal(got_sizes, want_sizes))
    stride = slot_stride(block)
    got = slots.cpu().numpy()
    bad = 0
    nb = len(want_sizes)
    for b0 in range(0, nb, 512):                       # compare in chunks of 512 blocks (memory)
        b1 = min(nb, b0 + 512)
        A = got[b0 * stride: b1 * stride].reshape(b1 - b0, stride)
        B = want_slots[b0 * stride: b1 * stride].reshape(b1 - b0, stride)
        idx = np.arange(stride, dtype=np.int64)[None, :] < want_sizes[b0:b1].astype(np.int64)[:, None]
        bad += int(((A != B) & idx).any(axis=1).sum())
    print(f"{name}: blocks {nb}, C/U {want_sizes.sum() / n:.4f}, sizes equal {same_sizes}, differing blocks {bad}, "
          f"device round trip {rt}  (reference took {time.time() - t0:.1f} s)", flush=True)
    del d, slots, sizes, out, osz, got, want_slots
    torch.cuda.empty_cache()
