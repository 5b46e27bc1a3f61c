#!/usr/bin/env bash
# Round 2, third GPU call (1 GPU): halo-tile 3x3 kernel — parity, A/B micro-benchmarks, step time.
set -u
O=gpurun_out/r2c
mkdir -p $O
python -m pytest tests/test_gpu_kernels_r2.py -m gpu -q -x -s -k "fprop or dgrad" > $O/t_halo.log 2>&1; echo "halo parity rc=$?"
python -m pytest tests/test_train_detector.py -m gpu -q -x -s > $O/t_loop.log 2>&1; echo "loop test rc=$?"
CASES="l1_conv2_3x3_64 l2_conv2_3x3_128 l2_conv2dgrad_3x3_128_mask rla_c3_recurrent_3x3_64 pred_cls_3x3_256_80_fp32"
python tools/conv_micro.py $CASES > $O/micro_halo.txt 2>&1; echo "micro halo rc=$?"
DSLB_NO_HALO=1 python tools/conv_micro.py $CASES > $O/micro_im2col.txt 2>&1; echo "micro im2col rc=$?"
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_c1.json 2> $O/bench_c1.err; echo "bench rc=$?"
DSLB_NO_HALO=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_c1_nohalo.json 2> $O/bench_c1_nohalo.err; echo "bench nohalo rc=$?"
tail -n 6 $O/t_halo.log $O/t_loop.log $O/t_all.log
cat $O/micro_halo.txt $O/micro_im2col.txt
for f in bench_c1 bench_c1_nohalo; do python - "$O/$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d.get("cand_counts"), d.get("det_counts"), d["clocks"])
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
