#!/usr/bin/env bash
set -u
O=gpurun_out/r2m2c
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
S=$(date +%s)
timeout 150 $TR tools/check_buckets.py > $O/check_buckets.txt 2>&1; echo "check_buckets rc=$? $(( $(date +%s) - S ))s"
S=$(date +%s)
timeout 120 $TR bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_c1_n2.json 2> $O/bench_c1_n2.err; echo "bench c1 n2 rc=$? $(( $(date +%s) - S ))s"
S=$(date +%s)
timeout 150 $TR bench.py --gpus 2 --workload configs4 --steps 40 --warmup 8 > $O/bench_c4_n2.json 2> $O/bench_c4_n2.err; echo "bench c4 n2 rc=$? $(( $(date +%s) - S ))s"
grep -E "OK|diverged|Error" $O/check_buckets.txt | head
for f in bench_c1_n2 bench_c4_n2; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("steady"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
done
