"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total, share."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
t, c = defaultdict(float), defaultdict(int)
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0][-60:]
    t[name] += float(r[vi].replace(",", ""))
    c[name] += 1
tot = sum(t.values())
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"{k:62s} n={c[k]:4d} total_us={v / 1e3:10.1f} avg_us={v / 1e3 / c[k]:8.1f} share={v / tot:.3f}")
