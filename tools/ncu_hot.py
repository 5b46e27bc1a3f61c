#!/usr/bin/env python
"""Program-order listing of the hot SASS instructions of kernel #k in an .ncu-rep source page (csv pre-exported).
usage: ncu_hot.py cases_src.csv k [min_frac]"""
import csv, sys
lines = open(sys.argv[1]).read().splitlines()
k = int(sys.argv[2]); mf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0025
starts = [i for i, l in enumerate(lines) if l.startswith('"Address"')] + [len(lines)]
rows = [r for r in csv.reader(lines[starts[k]:starts[k + 1]])]
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
rows = [r for r in rows[1:] if len(r) >= len(hdr)]
g = lambda r, c: int(r[idx[c]] or 0)
tot = sum(g(r, 'Instructions Executed') for r in rows)
tsm = sum(g(r, 'Warp Stall Sampling (All Samples)') for r in rows)
print('total warp inst', tot, 'samples', tsm)
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for n, r in enumerate(rows):
    ie = g(r, 'Instructions Executed'); sm = g(r, 'Warp Stall Sampling (All Samples)')
    if ie > mf * tot or sm > 0.006 * tsm:
        st = sorted(((g(r, c), c) for c in stall), reverse=True)[0]
        print(f"#{n:4d} ie={ie:8d} smp={sm:5d} {st[1][6:]:14s} {r[idx['Source']].strip()[:100]}")
