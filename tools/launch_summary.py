#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel name (time share, and DRAM traffic per launch when the byte counters were collected).
usage: launch_summary.py launches.csv"""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for row in csv.DictReader(lines):
    k = row["Kernel Name"][:72]
    m = row.get("Metric Name")
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        agg[k][0] += 1
        agg[k][1] += v
        tot += v
    elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        agg[k][2] += v * SC.get(u, 1.0)
print(f"{'us':>10s} {'n':>5s} {'share':>6s} {'dram MB/launch':>15s}  kernel")
for k, (n, t, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t:10.1f} {n:5d} {100 * t / tot:5.1f}% {b / max(n, 1) / 1e6:15.2f}  {k}")
print(f"{tot:10.1f} {sum(n for n, _, _ in agg.values()):5d} total (cold-cache, serialised: compare shares, not absolutes)")
