#!/usr/bin/env bash
# Round 2 (1 GPU): tower wgrad issued behind its layer's dgrad (DSLB_WGRAD_AFTER_DGRAD=1) vs beside it.
set -u
O=gpurun_out/r2t
mkdir -p $O
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one beside
  DSLB_WGRAD_AFTER_DGRAD=1 one behind
done
DSLB_WGRAD_AFTER_DGRAD=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or backward" > $O/t_behind.log 2>&1; echo "parity (behind) rc=$?"
tail -n 2 $O/t_behind.log
