"""Per-region stall attribution with instruction-mix labels. usage: ncu_regions.py rep [region]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; region = int(sys.argv[2]) if len(sys.argv) > 2 else 100
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src))); hdr = rows[1]; data = rows[2:]
ia, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
st = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(x[ia]) for x in data)
def op(x):
    t = x[1].strip().split()
    o = t[1] if t[0].startswith("@") else t[0]
    return o.split(".")[0]
for k in range(0, len(data), region):
    seg = data[k:k + region]; s = sum(int(x[ia]) for x in seg)
    if 100 * s / tot < 0.3: continue
    mix = Counter(op(x) for x in seg)
    stalls = Counter()
    for x in seg:
        for c in st: stalls[hdr[c][6:]] += int(x[c])
    top = ", ".join(f"{n}:{100*v/tot:.1f}" for n, v in stalls.most_common(3))
    ex = max(int(x[ie]) for x in seg)
    lab = " ".join(f"{n}{mix[n]}" for n in ("LDS", "FFMA", "FADD", "SHFL", "MUFU", "IMAD", "BAR", "STG", "STS", "SYNCS", "BRA") if mix[n])
    print(f"{k:5d} {100*s/tot:5.1f}%  exec {ex:8d}  [{top}]  {lab}")
