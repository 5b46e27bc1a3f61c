#!/usr/bin/env bash
# Round 2, second GPU call (1 GPU): new GPU tests (runner LR / closed loop / kernels), full suite, smoke, bench lines for
# configs[1] / [3] / [4] and a 1000-step sustained configs[1] line.   gpurun --timeout 1800 -- 'bash tools/gpu_r2_call2.sh'
set -u
O=gpurun_out/r2b
mkdir -p $O
python -m pytest tests/test_train_detector.py tests/test_gpu_kernels_r2.py -m gpu -q -x -s > $O/t_new.log 2>&1; echo "new tests rc=$?"
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench_c1.json 2> $O/bench_c1.err; echo "bench c1 rc=$?"
python bench.py --workload configs3 --steps 20 --warmup 5 --no-view-bench > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
python bench.py --workload configs4 --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench c4 rc=$?"
python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_c1_1000.json 2> $O/bench_c1_1000.err; echo "bench c1 x1000 rc=$?"
tail -n 4 $O/t_new.log $O/t_all.log $O/smoke.log
for f in c1 c3 c4 c1_1000; do python - "$O/bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "steady", "cand_counts")}, d["e2e"], r.get("frac"), r.get("traffic"), r.get("traffic_source", "")[:40], d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
