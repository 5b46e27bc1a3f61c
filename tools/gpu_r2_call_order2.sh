#!/usr/bin/env bash
# Round 2 (1 GPU): weight-gradient launch order, modes 0 / 1 / 2 of DSLB_WGRAD_AFTER_DGRAD.
set -u
O=gpurun_out/r2u
mkdir -p $O
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  DSLB_WGRAD_AFTER_DGRAD=1 one towers
  DSLB_WGRAD_AFTER_DGRAD=2 one towers_and_predictors
done
DSLB_WGRAD_AFTER_DGRAD=2 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size or backward" > $O/t_mode2.log 2>&1; echo "parity (mode 2) rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite (default = mode 1) rc=$?"
tail -n 2 $O/t_mode2.log $O/t_all.log
