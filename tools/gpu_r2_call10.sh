#!/usr/bin/env bash
set -u
O=gpurun_out/r2j
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_cta2.py -m gpu -q -x -s > $O/t_cta2.log 2>&1; echo "cta2 tests rc=$?"
grep -E "rel |passed|failed|Error|timeout|error" $O/t_cta2.log | head -20
if grep -q "passed" $O/t_cta2.log && ! grep -q "failed" $O/t_cta2.log; then
  python tools/conv_micro.py tower_3x3_256 > $O/micro_single.txt 2>&1
  DSLB_CTA2=1 python tools/conv_micro.py tower_3x3_256 > $O/micro_cta2.txt 2>&1
  cat $O/micro_single.txt $O/micro_cta2.txt
  B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic"
  P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["head_tower"]["tflops"], d["clocks"]["sm_mhz"])'
  for i in 1 2; do
    DSLB_CTA2=1 $B 2>/dev/null | python -c "$P" cta2
    $B 2>/dev/null | python -c "$P" single
  done
fi
tail -n 15 $O/t_cta2.log
