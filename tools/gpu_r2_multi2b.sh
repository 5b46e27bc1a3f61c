#!/usr/bin/env bash
set -u
O=gpurun_out/r2m2b
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
DSLB_HANG_DUMP=100 timeout 150 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_n2_onegraph.json 2> $O/bench_n2_onegraph.err; echo "bench n2 one-graph rc=$?"
grep -v "^W1017\|^\*\*\*\|OMP_NUM" $O/bench_n2_onegraph.err | head -80
cat $O/bench_n2_onegraph.json | cut -c1-300
