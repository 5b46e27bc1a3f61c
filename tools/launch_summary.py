#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    k = row["Kernel Name"][:72]
    agg[k][0] += 1
    agg[k][1] += v
    tot += v
print(f"{'us':>10s} {'n':>5s} {'share':>6s}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} {n:5d} {100 * t / tot:5.1f}%  {k}")
print(f"{tot:10.1f} {sum(n for n, _ in agg.values()):5d} total (cold-cache, serialised: compare shares, not absolutes)")
