#!/usr/bin/env bash
# Round 2, closing build: the other single-GPU workloads (configs[3] R101 bs 2, configs[4] multi-scale, RLA_R50 backbone).
set -u
O=gpurun_out/r2v
mkdir -p $O
X="--no-cpu-baseline --no-view-bench --no-ncu-traffic"
timeout 200 python bench.py --workload configs3 --steps 30 --warmup 5 $X > $O/bench_c3.json 2> $O/bench_c3.err; echo "c3 rc=$?"
timeout 240 python bench.py --workload configs4 --steps 40 --warmup 8 $X > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 rc=$?"
timeout 200 python bench.py --backbone rla --steps 30 --warmup 5 $X > $O/bench_rla.json 2> $O/bench_rla.err; echo "rla rc=$?"
for f in bench_c3 bench_c4 bench_rla; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("steady"), d["clocks"]["sm_mhz"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
done
