"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python scripts/launch_summary.py launches.csv "command line that was profiled" """
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    us = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu].strip(), 1.0) * v
    a = agg[r[ik]]; a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else ''}")
print("# (per-launch times under ncu are cold-cache and serialised: compare shares)")
print("# kernel, launches, total us, share")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:80]:80s} {a[0]:6d} {a[1]:12.1f} {100 * a[1] / tot:6.2f}%")
