#!/usr/bin/env bash
set -u
O=gpurun_out/r2h
mkdir -p $O
python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
tail -n 3 $O/t_all.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["gpu_launches_per_step"])'
for i in 1 2; do
$B 2>/dev/null | python -c "$P" joint
DSLB_JOINT_FWD=0 $B 2>/dev/null | python -c "$P" separate
done
$B --workload configs3 2>/dev/null | python -c "$P" c3-joint
DSLB_JOINT_FWD=0 $B --workload configs3 2>/dev/null | python -c "$P" c3-separate
python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1
grep -E "^rank|us/step" $O/timeline.txt | head -8
