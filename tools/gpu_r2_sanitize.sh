#!/usr/bin/env bash
set -u
O=gpurun_out/r2p
mkdir -p $O
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_cta2.py -m gpu -q -x > $O/memcheck_cta2.log 2>&1; echo "memcheck cta2 rc=$?"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -k "halo or stem or (wgrad and pairs) or topk or split or predictor or gn_stats" > $O/memcheck_r2.log 2>&1; echo "memcheck r2 rc=$?"
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -k "halo or stem" > $O/racecheck_r2.log 2>&1; echo "racecheck r2 rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bf16x3" > $O/memcheck_bf16x3.log 2>&1; echo "memcheck bf16x3 rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $O/*.log | tail -12
