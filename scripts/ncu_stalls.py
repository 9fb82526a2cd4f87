"""Development helper: stall samples per source line (long scoreboard / wait / total) from an .ncu-rep.
usage: python scripts/ncu_stalls.py rep.ncu-rep [kernel-regex] [top]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["-k", "regex:" + sys.argv[2]]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None; out = []
for r in rows:
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit(): continue
    g = lambda n: int(r[hdr.index(n)]) if r[hdr.index(n)].isdigit() else 0
    out.append((g('# Samples'), g('stall_long_sb'), g('stall_wait'), g('stall_short_sb'), g('Instructions Executed'), int(r[0]), r[1][:100]))
tot = sum(o[0] for o in out) or 1
print("total samples", tot, " total inst %.2f G" % (sum(o[4] for o in out) / 1e9))
for o in sorted(out, reverse=True)[:top]:
    print(f"{100*o[0]/tot:5.1f}% samp  long_sb {100*o[1]/tot:5.1f}%  wait {100*o[2]/tot:4.1f}%  short {100*o[3]/tot:4.1f}%  inst {o[4]/1e6:8.1f}M  L{o[5]} {o[6]}")
