#!/usr/bin/env bash
# Round 2 (1 GPU): EMA of the trainable regions inside the SGD kernel; backward preparation forked at the head towers.
set -u
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one fused_ema
  DSLB_FUSE_EMA=0 one separate_ema
  DSLB_PREP_UNDER_FWD=towers one fused_ema_prep_towers
done
tail -n 3 $O/t_all.log
