#!/usr/bin/env bash
set -u
O=gpurun_out/r2m
mkdir -p $O
DSLB_PLAN_TABLE=$O/plans_cta2.jsonl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_cta2.json 2>/dev/null
DSLB_CTA2=0 DSLB_PLAN_TABLE=$O/plans_single.jsonl python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_single.json 2>/dev/null
python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1
grep -E "^rank|us/step" $O/timeline.txt | head -14
