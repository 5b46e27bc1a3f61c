#!/usr/bin/env bash
# Round 2 (1 GPU): max-pool with two outputs per thread, |g|^2 per gradient bucket on the side stream.
set -u
O=gpurun_out/r2p
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one bucket_sqnorm
  DSLB_BUCKET_SQNORM=0 one whole_sqnorm
done
timeout 600 python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 3 $O/t_all.log
grep -E "maxpool|sqnorm|span" $O/timeline.txt
