#!/usr/bin/env bash
# Round 2, 8-GPU call: BASELINE configs[2] (= configs[1] at 8 x B200), configs[3] (R101 bs 2), configs[4] (multi-scale) lines,
# an N=1 line on the same box for the efficiency, the kernel timeline of an N=8 step.
set -u
O=gpurun_out/r2m8
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518"
run() { S=$(date +%s); "${@:2}" > $O/$1.json 2> $O/$1.err; echo "$1 rc=$? $(( $(date +%s) - S ))s"; }
run bench_c1_n1 timeout 120 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic
run bench_c1_n8 timeout 150 $TR bench.py --gpus 8 --steps 30 --warmup 5
run bench_c1_n8_sixgraphs timeout 150 env DSLB_GRAPH_NCCL=0 $TR bench.py --gpus 8 --steps 30 --warmup 5
run bench_c1_n8_maxctas8 timeout 150 env NCCL_MAX_CTAS=8 $TR bench.py --gpus 8 --steps 30 --warmup 5
run bench_c3_n8 timeout 150 $TR bench.py --gpus 8 --workload configs3 --steps 30 --warmup 5
run bench_c4_n8 timeout 200 $TR bench.py --gpus 8 --workload configs4 --steps 40 --warmup 8
S=$(date +%s); timeout 150 $TR tools/step_timeline.py --steps 2 --out $O/timeline_n8.jsonl > $O/timeline_n8.txt 2>&1; echo "timeline rc=$? $(( $(date +%s) - S ))s"
rm -f $O/timeline_n8.rank[1-7].jsonl
for f in bench_c1_n1 bench_c1_n8 bench_c1_n8_sixgraphs bench_c1_n8_maxctas8 bench_c3_n8 bench_c4_n8; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("steady"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-800:])
PY
done
grep -E "^rank 0|nccl" $O/timeline_n8.txt | head -8
