#!/usr/bin/env bash
# Round 2 (1 GPU): gradient-buffer memsets beside the loss kernel (DSLB_PREP_UNDER_FWD=loss) vs in front of the backward.
set -u
O=gpurun_out/r2x
mkdir -p $O
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one in_backward
  DSLB_PREP_UNDER_FWD=loss one beside_loss
done
DSLB_PREP_UNDER_FWD=loss timeout 600 python -m pytest tests -m gpu -q -x > $O/t_all_loss_mode.log 2>&1; echo "gpu suite (beside-loss mode) rc=$?"
tail -n 2 $O/t_all_loss_mode.log
