#!/usr/bin/env bash
# Round 2 (1 GPU): merged / wide-tile operand refresh. Parity of the table launch, then an interleaved A/B of the step
# against a library built with the previous pack kernel (DSLB_LIB), and the timeline of the new build.
set -u
O=gpurun_out/r2n
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -k "pack_table" > $O/t_pack.log 2>&1; echo "pack test rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
one() {
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['clocks']['sm_mhz'])"
}
for i in 1 2 3; do
  one new
  DSLB_LIB=$PWD/dsl_b200/libdslb_old.so one old
done
timeout 600 python tools/step_timeline.py --steps 2 --out $O/timeline.jsonl > $O/timeline.txt 2>&1; echo "timeline rc=$?"
tail -n 3 $O/t_pack.log $O/t_all.log
grep -E "pack_|ema_kernel|sgd_kernel|span" $O/timeline.txt
