#!/usr/bin/env bash
set -u
O=gpurun_out/r2f
mkdir -p $O
python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -s -k "stem" > $O/t_stem.log 2>&1; echo "stem parity rc=$?"
python - > $O/stem_micro.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, ".")
from tools.ncu_cases import build_stem
for env in ("new",):
    plans = build_stem()
    for p in plans: p.run()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plans[i % 2].run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print("stem (pre-pass + conv) us:", sorted(ts))
PY
cat $O/stem_micro.txt
DSLB_OLD_STEM=1 python - <<'PY'
import sys, torch
sys.path.insert(0, ".")
from tools.ncu_cases import build_stem
plans = build_stem()
for p in plans: p.run()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plans[i % 2].run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print("old stem (pre-pass + conv) us:", sorted(ts))
PY
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('now', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
tail -n 3 $O/t_stem.log $O/t_all.log
