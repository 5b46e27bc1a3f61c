#!/usr/bin/env bash
# Round 2 (1 GPU): gradient-buffer memsets + target assignment under the forward pass.
set -u
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one prep_under_fwd
  DSLB_PREP_UNDER_FWD=0 one prep_in_bwd
done
tail -n 3 $O/t_all.log
