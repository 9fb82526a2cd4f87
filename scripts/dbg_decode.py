"""Development helper: locate decode mismatches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import turbosqueeze_b200 as T
from turbosqueeze_b200 import workloads as W
from oraclelib import Oracle, slot_stride

kind = sys.argv[1]; n = int(sys.argv[2]); block = int(sys.argv[3]); lanes = int(sys.argv[4]) if len(sys.argv) > 4 else 34
ctx = T.Context(0)
ctx.set_option("decode_lanes", lanes)
buf = W.fill(kind, n, seed=n + 5)
slots, sizes, _ = Oracle().encode_blocks(buf, n, block, 0)
nb = len(sizes)
d_slots = torch.from_numpy(slots).cuda(); d_sizes = torch.from_numpy(sizes.astype(np.int32)).cuda()
out, osz = ctx.decode_blocks(d_slots, nb, block, 0, comp_sizes=d_sizes)
torch.cuda.synchronize()
o = out.cpu().numpy()[:n]
bad = np.flatnonzero(o != buf[:n])
print(kind, n, block, "mismatches", bad.size, "osz ok", int(osz.sum().item()) == n)
if bad.size:
    print("first", bad[:20], "block", bad[0] // block, "offset", bad[0] % block)
    i = int(bad[0]); print("got ", o[max(0, i - 8):i + 24].tolist()); print("want", buf[max(0, i - 8):i + 24].tolist())
    runs = np.split(bad, np.flatnonzero(np.diff(bad) > 1) + 1)
    print("runs", len(runs), [(int(r[0]) % block, len(r)) for r in runs[:12]])
