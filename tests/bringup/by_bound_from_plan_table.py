"""TEST INFRASTRUCTURE (run by hand, CPU only, ~25 GB of RAM): the roofline.by_bound split bench.py reports, derived offline from
a committed per-plan timing table (profiles/r01d_plans.txt: CUDA-event time of every implicit-GEMM launch of one eager
teacher+student step at B=4, 800x1344) and the algorithmic HBM bytes of the same plans (engine.seg_bytes; the full-size plan
list is built on tests/emu_lib.py, nothing is executed).  python tests/bringup/by_bound_from_plan_table.py"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from dsl_b200 import engine as E
from dsl_b200.trainer import DSLEngine
from tests import emu_lib
with emu_lib.installed():
    net = E.FCOSNet(4, 800, 1344, 50, 80, train=True, device="cpu", seed=0, parts="all")
    plans = [getattr(op, "__self__", None) for op in net.fwd_ops + net.bwd_ops]
    by_name = {}
    for p in plans:
        if isinstance(p, E.ConvPlan):
            by_name.setdefault(p.what, p.bytes)
    del net, plans
rows, missing = [], []
for ln in open(os.path.join(ROOT, 'profiles', 'r01d_plans.txt')):
    m = re.match(r"\s*\d+\s+(\S+)\s+(conv_igemm|conv_wgrad)\s+([\d.]+) us\s+([\d.]+) GF", ln)
    if not m or m.group(2) != "conv_igemm": continue
    name, us, gf = m.group(1), float(m.group(3)), float(m.group(4))
    if name not in by_name: missing.append(name); continue
    rows.append((us*1e-3, gf*1e9, float(by_name[name])))
print("joined", len(rows), "missing", sorted(set(missing))[:10])
ridge = 1378.7e12/6550.7e9
out = DSLEngine.split_by_bound(rows, ridge)
t, h = out["tensor"], out["hbm"]
print(json.dumps(dict(ridge=round(ridge,1),
  tensor=dict(n=t["n"], ms=round(t["ms"],3), tflops=round(t["flops"]/t["ms"]/1e9,1), frac=round(t["flops"]/t["ms"]/1e9/1378.7,3)),
  hbm=dict(n=h["n"], ms=round(h["ms"],3), gbps=round(h["bytes"]/h["ms"]/1e6,1), frac=round(h["bytes"]/h["ms"]/1e6/6550.7,3), tflops=round(h["flops"]/h["ms"]/1e9,1)))))
names = {"tensor": {}, "hbm": {}}
for ln in open(os.path.join(ROOT, 'profiles', 'r01d_plans.txt')):
    m = re.match(r"\s*\d+\s+(\S+)\s+(conv_igemm|conv_wgrad)\s+([\d.]+) us\s+([\d.]+) GF", ln)
    if not m or m.group(2) != "conv_igemm":
        continue
    name, us, gf = m.group(1), float(m.group(3)), float(m.group(4))
    c = "hbm" if gf * 1e9 / by_name[name] < ridge else "tensor"
    key = re.sub(r"\.\d+\.", ".*.", name)
    a = names[c].setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += us
for c in names:
    print(c + ":", ", ".join(f"{k} x{n} {us:.0f}us" for k, (n, us) in sorted(names[c].items(), key=lambda kv: -kv[1][1])))
