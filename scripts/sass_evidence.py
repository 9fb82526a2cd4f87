"""Per-kernel SASS evidence for profiles/: counts of the mnemonics that show what the kernels are built from
(bulk async copies + mbarrier transaction barriers = TMA without a tensor map, 128/256-bit global accesses, warp votes /
shuffles / match) and that nothing of another architecture is in the library.
usage: python scripts/sass_evidence.py [library.so] > profiles/rNN_sass_evidence.txt"""
import collections, re, subprocess, sys

lib = sys.argv[1] if len(sys.argv) > 1 else "turbosqueeze_b200/libturbosqueeze_b200.so"
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
PAT = collections.OrderedDict([
    ("UBLKCP (cp.async.bulk global->shared)", r"\bUBLKCP"), ("SYNCS (mbarrier arrive / try_wait, incl. TRANS64)", r"\bSYNCS"),
    ("LDG.*.128", r"\bLDG\.[A-Z0-9.]*128"), ("LDG.*.256 (ENL2.256)", r"\bLDG\.[A-Z0-9.]*256"), ("STG.*.128", r"\bSTG\.[A-Z0-9.]*128"),
    ("STG.*.256", r"\bSTG\.[A-Z0-9.]*256"), ("LDG (all)", r"\bLDG\b|\bLDG\."), ("STG (all)", r"\bSTG\b|\bSTG\."), ("LDS", r"\bLDS\b|\bLDS\."),
    ("STS", r"\bSTS\b|\bSTS\."), ("VOTE / VOTEU", r"\bVOTEU?\b|\bVOTEU?\."), ("SHFL", r"\bSHFL"), ("MATCH", r"\bMATCH"), ("REDUX", r"\bREDUX"),
    ("R2P", r"\bR2P"), ("SHF (funnel shifts)", r"\bSHF\."), ("PRMT", r"\bPRMT"), ("NANOSLEEP", r"\bNANOSLEEP"), ("HMMA/IMMA/UTCMMA (tensor cores)", r"\b(HMMA|IMMA|UTC\w*MMA|QGMMA)"),
])
print(f"# SASS evidence for {lib}")
print("# cubins in the library (cuobjdump -lelf):")
for l in elf.splitlines():
    if l.strip():
        print("#   " + l.strip())
cur, counts, sizes = None, collections.OrderedDict(), collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter(); sizes[cur] = 0
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        sizes[cur] += 1
        for name, pat in PAT.items():
            if re.search(pat, line):
                counts[cur][name] += 1
def demangle(n):
    try:
        return subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n
for fn in counts:
    print(f"\n== {demangle(fn)}\n   {sizes[fn]} SASS instructions ({sizes[fn] * 16 / 1024:.1f} KiB)")
    for name in PAT:
        if counts[fn][name]:
            print(f"   {name:55s} {counts[fn][name]}")
