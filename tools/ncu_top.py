#!/usr/bin/env python
"""Summarise the SASS page of an .ncu-rep: top instructions by warp-stall samples, with their dominant stall reason.
usage: ncu_top.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(lines[start:]))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
tot = 0
for n, r in enumerate(rows[1:]):
    if len(r) < len(hdr):
        continue
    s = int(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    tot += s
    recs.append((s, n, r))
print("total samples", tot)
for s, n, r in sorted(recs, key=lambda x: -x[0])[:top]:
    st = sorted(((int(r[idx[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{s:7d} {100 * s / tot:5.1f}%  #{n:4d} {r[idx['Source']].strip()[:70]:70s} {st[0][1]}={st[0][0]} {st[1][1]}={st[1][0]}")
