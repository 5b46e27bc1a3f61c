#!/usr/bin/env python
"""Stall samples of an .ncu-rep aggregated by CUDA source line (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
recs = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 7 and r[0] not in ("", "Line No"):
        try:
            recs.append((int(r[4]), fname, int(r[0]), r[1][:100], r[7]))
        except ValueError:
            pass
tot = sum(x[0] for x in recs)
print("total samples", tot)
for s, f, l, src, ie in sorted(recs, reverse=True)[:top]:
    print(f"{s:6d} {100 * s / tot:5.1f}% {f}:{l:<4d} instr={ie:>9s} {src}")
