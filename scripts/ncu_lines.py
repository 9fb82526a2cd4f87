"""Development helper: per-source-line instruction counts and stall samples from an .ncu-rep.
usage: python scripts/ncu_lines.py rep.ncu-rep [kernel-regex] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["-k", "regex:" + sys.argv[2]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
out = []
for r in rows:
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    gi = lambda name: int(r[hdr.index(name)]) if r[hdr.index(name)].isdigit() else 0
    out.append((gi("Instructions Executed"), gi("# Samples"), int(r[0]), r[1][:110]))
tot_i = sum(x[0] for x in out) or 1
tot_s = sum(x[1] for x in out) or 1
print(f"total inst {tot_i}  samples {tot_s}")
for i, s, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*i/tot_i:5.1f}% inst {100*s/tot_s:5.1f}% stall  L{ln:4d}  {src}")
