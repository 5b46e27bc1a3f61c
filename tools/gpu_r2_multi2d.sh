#!/usr/bin/env bash
# Round 2, 2 GPUs: bucket boundaries without a dgrad-stream join + per-bucket |g|^2 on the communication stream:
# correctness (bucketed == single all-reduce, ranks identical) and an interleaved A/B of the step time.
set -u
O=gpurun_out/r2m2d
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
S=$(date +%s)
timeout 150 $TR tools/check_buckets.py > $O/check_buckets.txt 2>&1; echo "check_buckets rc=$? $(( $(date +%s) - S ))s"
grep -E "OK|diverged|Error|graphs=" $O/check_buckets.txt | cut -c1-110 | head
one() {
  timeout 120 $TR bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>$O/err_$1.txt | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2; do
  one lazy
  DSLB_BUCKET_JOIN=1 DSLB_BUCKET_SQNORM=0 one joined
done
echo "total $(( $(date +%s) - S ))s"
