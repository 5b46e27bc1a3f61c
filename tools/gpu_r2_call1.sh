#!/usr/bin/env bash
# Round 2, first GPU call: full GPU suite incl. the round-2 identical-input parity net, compute-sanitizer memcheck /
# racecheck on the per-kernel tests, smoke, baseline bench line.   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'
set -u
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -s > $O/t_r2.log 2>&1; echo "r2 kernel tests rc=$?"
python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x \
    -k "fprop or dgrad or wgrad or stem or topk or predictor" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x \
    -k "topk or stem or (fprop and 3x3)" > $O/racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "loss_kernels or identical or decode_nms or pseudo_label" > $O/memcheck_parity.log 2>&1; echo "memcheck parity rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
tail -n 5 $O/t_r2.log $O/t_all.log $O/smoke.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/memcheck.log $O/racecheck.log $O/memcheck_parity.log | tail -12
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2a/bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], d["roofline"]["frac"], d["clocks"])
except Exception as e:
    print("bench line unreadable:", e)
PY
