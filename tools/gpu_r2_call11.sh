#!/usr/bin/env bash
set -u
O=gpurun_out/r2k
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_cta2.py -m gpu -q -x -s > $O/t_cta2.log 2>&1; echo "cta2 tests rc=$?"
grep -E "rel |passed|failed|Error|timeout|error" $O/t_cta2.log | head -20
timeout 400 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/t_all.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["head_tower"]["tflops"], d["roofline"]["frac_of_sustained_peak"], d["clocks"]["sm_mhz"])'
for i in 1 2; do
  $B 2>/dev/null | python -c "$P" cta2-kmin8
  DSLB_CTA2_KMIN=18 $B 2>/dev/null | python -c "$P" cta2-kmin18
  DSLB_CTA2_KMIN=4 $B 2>/dev/null | python -c "$P" cta2-kmin4
  DSLB_CTA2=0 $B 2>/dev/null | python -c "$P" single
done
