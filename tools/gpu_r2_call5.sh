#!/usr/bin/env bash
# Round 2, fifth GPU call (1 GPU): direct-patch stem kernel (parity, A/B), ncu --set full of the final-state tower launch,
# the halo kernel and the stem.
set -u
O=gpurun_out/r2e
mkdir -p $O
python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -s -k "stem" > $O/t_stem.log 2>&1; echo "stem parity rc=$?"
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
for i in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stem2', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
DSLB_OLD_STEM=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('oldstem', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
done
ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $O/cases \
    python tools/ncu_cases.py tower_full stem_full l1_conv2_3x3_64 l2_conv2_3x3_128 > $O/ncu_cases.log 2>&1; echo "ncu rc=$?"
ncu -i $O/cases.ncu-rep --page raw --csv > $O/cases_raw.csv 2>/dev/null
python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1
tail -n 5 $O/t_stem.log $O/t_all.log
grep -E "us/step" $O/timeline.txt | head -12
