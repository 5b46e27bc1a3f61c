#!/usr/bin/env bash
# Round 2 (1 GPU): GroupNorm backward sums in the dgrad epilogue + finalize folded into the apply kernels.
set -u
O=gpurun_out/r2o
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -k "gn_bwd or gn_stats" -s > $O/t_gnb.log 2>&1; echo "gnb tests rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one fused
  DSLB_GN_BWD_FUSED=0 one unfused
done
timeout 600 python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1; echo "timeline rc=$?"
grep -E "gnb sums|passed|failed|Error|error" $O/t_gnb.log | head -20
tail -n 3 $O/t_all.log
grep -E "gn_|span" $O/timeline.txt
