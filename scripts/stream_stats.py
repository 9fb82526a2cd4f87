"""Development helper (CPU): statistics of reference token streams that drive the decoder design."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from oraclelib import Oracle, slot_stride
from turbosqueeze_b200 import workloads as W

def parse(s):
    size = s[0] | s[1] << 8 | s[2] << 16
    i, j = 3, 0
    syms = []  # (is_lit, len, dst, src, pair_start)
    while j < size:
        ctl = s[i]; i += 1
        for p in range(4):
            if j >= size: break
            nib = s[i]; i += 1
            org = j
            for half in range(2):
                if j >= size: break
                ln = ((nib >> 4) if half == 0 else (nib & 15)) + 1
                lit = (ctl >> (7 - (2 * p + half))) & 1
                if lit:
                    syms.append((1, ln, j, -1, org)); i += ln
                else:
                    off = s[i] | s[i + 1] << 8; i += 2
                    syms.append((0, ln, j, org - off, org))
                j += ln
    return syms, i

kind = sys.argv[1] if len(sys.argv) > 1 else "text"
block = 262144
n = block * 4
buf = W.fill(kind, n, seed=20240917)
slots, sizes, _ = Oracle().encode_blocks(buf, n, block, 0)
stride = slot_stride(block)
s = [int(x) for x in slots[stride:stride + sizes[1]]]
syms, used = parse(s)
ns = len(syms)
lits = [x for x in syms if x[0]]
mats = [x for x in syms if not x[0]]
print(f"{kind}: symbols {ns}, literals {len(lits)} ({100*len(lits)/ns:.1f}%), mean lit {np.mean([x[1] for x in lits]):.2f}, mean match {np.mean([x[1] for x in mats]) if mats else 0:.2f}, comp {sizes[1]}")
offs = np.array([x[2] - x[3] for x in mats])
for t in (16, 64, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768):
    print(f"  match distance(dst-src) < {t}: {100*np.mean(offs < t):.1f}%")
for step in (8, 16, 32, 64):
    rounds_tot = 0; nsteps = 0; pend_tot = 0
    for a in range(0, ns, step):
        grp = syms[a:a + step]
        J0 = grp[0][2]
        depth = {}
        maxd = 0
        # depth of symbol = 1 + max depth of symbols in this step overlapping its source
        for k, (lit, ln, dst, src, org) in enumerate(grp):
            d = 0
            if not lit and src + ln > J0:
                pend_tot += 1
                for k2 in range(k):
                    l2, ln2, dst2, _, _ = grp[k2]
                    if dst2 < src + ln and dst2 + ln2 > src:
                        d = max(d, depth[k2] + 1)
                d = max(d, 1)
            depth[k] = d
            maxd = max(maxd, d)
        rounds_tot += maxd + 1; nsteps += 1
    print(f"  step {step}: mean rounds {rounds_tot/nsteps:.2f}, pending symbols/step {pend_tot/nsteps:.2f}")
