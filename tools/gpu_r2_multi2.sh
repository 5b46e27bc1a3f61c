#!/usr/bin/env bash
# Round 2, 2-GPU call: bucketed / one-graph all-reduce equivalence + rank-local capture (tools/check_buckets.py), bench at
# N=2 with the collectives inside ONE graph vs six graphs, kernel timeline of an N=2 step.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_multi2.sh'
set -u
O=gpurun_out/r2m2
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/check_buckets.py > $O/check_buckets.txt 2>&1; echo "check_buckets rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench n1 rc=$?"
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2_onegraph.json 2> $O/bench_n2_onegraph.err; echo "bench n2 one-graph rc=$?"
DSLB_GRAPH_NCCL=0 timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > $O/bench_n2_sixgraphs.json 2> $O/bench_n2_sixgraphs.err; echo "bench n2 six-graphs rc=$?"
timeout 300 $TR tools/step_timeline.py --steps 2 --out $O/timeline_n2.jsonl > $O/timeline_n2.txt 2>&1; echo "timeline n2 rc=$?"
tail -n 6 $O/check_buckets.txt
for f in bench_n1 bench_n2_onegraph bench_n2_sixgraphs; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
cat $O/timeline_n2.txt | tail -40
