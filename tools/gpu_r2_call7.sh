#!/usr/bin/env bash
set -u
O=gpurun_out/r2g
mkdir -p $O
python -m pytest tests/test_gpu_parity.py tests/test_plugin.py -m gpu -q -x -s -k "bf16x3 or accurate or head_forward" > $O/t_head.log 2>&1; echo "head tests rc=$?"
grep -E "rel |passed|failed|Error" $O/t_head.log | head -30
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
tail -n 4 $O/t_all.log
