#!/usr/bin/env bash
set -u
O=gpurun_out/r2l
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -s -k "wgrad" > $O/t_wgrad.log 2>&1; echo "wgrad tests rc=$?"
grep -E "rel |passed|failed|Error|timeout|error" $O/t_wgrad.log | head -20
if grep -q "passed" $O/t_wgrad.log && ! grep -q "failed" $O/t_wgrad.log; then
timeout 400 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/t_all.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d["roofline"]; print(sys.argv[1], d["value"], d["ms_per_step"], r["head_tower"]["tflops"], r["wgrad"]["achieved"], r["frac_of_sustained_peak"], d["clocks"]["sm_mhz"])'
for i in 1 2; do
  $B 2>/dev/null | python -c "$P" cta2
  DSLB_CTA2=0 $B 2>/dev/null | python -c "$P" single
done
fi
