#!/usr/bin/env bash
# Round 2, closing 1-GPU call: full GPU suite, smoke, the driver's default bench line, the sustained 1000-step line,
# compute-sanitizer over the kernels added since the second sanitizer pass, timeline of the final build.
set -u
O=gpurun_out/r2s
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py > $O/bench_c1.json 2> $O/bench_c1.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 1000 --warmup 10 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_c1_1000.json 2> $O/bench_c1_1000.err; echo "bench 1000 rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x \
  -k "pack_table or gn_bwd or maxpool or sgd_with_ema" > $O/memcheck_new.log 2>&1; echo "memcheck rc=$?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x \
  -k "pack_table or gn_bwd_one_pass" > $O/racecheck_new.log 2>&1; echo "racecheck rc=$?"
timeout 600 python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 3 $O/t_all.log; tail -n 2 $O/smoke.log
grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $O/memcheck_new.log $O/racecheck_new.log
for f in bench_c1 bench_c1_1000; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"].get("frac"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
head -3 $O/timeline.txt
