"""Development helper: instructions / stall samples of one kernel aggregated over line ranges ("phases") of one source file.
usage: python scripts/ncu_phases.py rep.ncu-rep kernel-regex file-suffix  lo-hi:name [lo-hi:name ...]"""
import csv, io, subprocess, sys
rep, kre, suffix = sys.argv[1:4]
phases = []
for a in sys.argv[4:]:
    rng, name = a.split(":")
    lo, hi = rng.split("-")
    phases.append((int(lo), int(hi), name))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + kre], capture_output=True, text=True).stdout
cur = None; hdr = None
acc = {}
tot_i = tot_s = 0
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or not r[0].isdigit() or len(r) < len(hdr): continue
    g = lambda n: int(r[hdr.index(n)]) if r[hdr.index(n)].isdigit() else 0
    i, s, l = g("Instructions Executed"), g("# Samples"), g("stall_long_sb")
    tot_i += i; tot_s += s
    key = "other files: " + cur.split("/")[-1]
    if cur.endswith(suffix):
        ln = int(r[0]); key = "unassigned"
        for lo, hi, name in phases:
            if lo <= ln <= hi: key = name; break
    a = acc.setdefault(key, [0, 0, 0]); a[0] += i; a[1] += s; a[2] += l
print(f"total inst {tot_i/1e9:.2f} G, samples {tot_s}")
for k, (i, s, l) in sorted(acc.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*i/tot_i:5.1f}% inst  {100*s/tot_s:5.1f}% samples  {100*l/tot_s:5.1f}% long_sb   {k}")
