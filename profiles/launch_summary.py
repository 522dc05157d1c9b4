#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file launches.csv python bench.py ...
    python profiles/launch_summary.py launches.csv > profiles/launches_X.txt
"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
cols = {h: i for i, h in enumerate(rows[hdr])}
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rows[hdr + 1:]:
    if len(r) <= cols["Metric Value"] or r[cols["Metric Name"]] != "gpu__time_duration.sum":
        continue
    unit = r[cols["Metric Unit"]]
    val = float(r[cols["Metric Value"]].replace(",", ""))
    ms = val * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
    name = r[cols["Kernel Name"]]
    tot[name] += ms
    cnt[name] += 1
total = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)")
for name, ms in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{ms:10.3f} ms {100 * ms / total:5.1f}%  x{cnt[name]:3d}  {name[:110]}")
