#!/usr/bin/env bash
set -u
O=gpurun_out/r2o
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520"
run() { S=$(date +%s); "${@:2}" > $O/$1.json 2> $O/$1.err; echo "$1 rc=$? $(( $(date +%s) - S ))s"; }
run bench_c1_n1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic
run bench_c1_n8 timeout 150 $TR bench.py --gpus 8 --steps 30 --warmup 5
run bench_c3_n8 timeout 150 $TR bench.py --gpus 8 --workload configs3 --steps 30 --warmup 5
run bench_c4_n8 timeout 200 $TR bench.py --gpus 8 --workload configs4 --steps 40 --warmup 8
for f in bench_c1_n1 bench_c1_n8 bench_c3_n8 bench_c4_n8; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("steady"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
done
