#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, data = r, rows[i + 1 :]
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
tot = 0.0
for r in data:
    name = r[ki].split("(")[0].replace("void <unnamed>::", "").replace("<unnamed>::", "")[:48]
    v = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r[ui], 1)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{'kernel':48s} {'n':>5s} {'total_us':>11s} {'avg_us':>9s} {'share':>6s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} {n:5d} {v / 1e3:11.1f} {v / n / 1e3:9.1f} {v / tot * 100:5.1f}%")
print(f"{'TOTAL':48s} {sum(n for n, _ in agg.values()):5d} {tot / 1e3:11.1f}")
