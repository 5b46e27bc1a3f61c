#!/usr/bin/env bash
# Round 2 (1 GPU): the fused GroupNorm backward re-measured under the new wgrad-behind-dgrad order.
set -u
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one unfused
  DSLB_GN_BWD_FUSED=1 one fused
done
