#!/usr/bin/env bash
# Round 2, fourth GPU call (1 GPU): kernel timeline of the graph step (halo on / off), full-size halo parity, loop test.
set -u
O=gpurun_out/r2d
mkdir -p $O
python -m pytest tests/test_train_detector.py tests/test_gpu_rla.py -m gpu -q -x > $O/t_loop.log 2>&1; echo "loop+rla tests rc=$?"
python tools/step_timeline.py --steps 2 --out $O/timeline_halo.jsonl > $O/timeline_halo.txt 2>&1; echo "timeline rc=$?"
DSLB_NO_HALO=1 python tools/step_timeline.py --steps 2 --out $O/timeline_nohalo.jsonl > $O/timeline_nohalo.txt 2>&1; echo "timeline nohalo rc=$?"
for i in 1 2 3; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('halo', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
DSLB_NO_HALO=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nohalo', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
tail -n 5 $O/t_loop.log
cat $O/timeline_halo.txt $O/timeline_nohalo.txt
