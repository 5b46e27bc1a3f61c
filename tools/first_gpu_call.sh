#!/usr/bin/env bash
# First GPU call of a round: everything that was written without a GPU gets run and measured once.
#   gpurun --timeout 900 -- 'bash tools/first_gpu_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_view_image_gpu.py -m gpu -x -q > gpurun_out/t_new.log 2>&1
echo "new gpu tests rc=$?"
python tools/view_image_micro.py > gpurun_out/view_image_micro.json 2> gpurun_out/view_image_micro.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:view_images --csv --log-file gpurun_out/view_images_launches.csv \
    python tools/view_image_micro.py --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:view_images -c 1 -f -o gpurun_out/view_images \
    python tools/view_image_micro.py --iters 1 > /dev/null 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/t_all.log 2>&1
echo "gpu suite rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
echo "bench rc=$?"
python bench.py --impl eager-gpu --steps 5 --warmup 2 > gpurun_out/eager_gpu_comparator.jsonl 2> gpurun_out/eager_gpu_comparator.err
echo "eager comparator rc=$?"
cat gpurun_out/eager_gpu_comparator.jsonl
tail -n 3 gpurun_out/t_new.log gpurun_out/t_all.log
cat gpurun_out/view_image_micro.json
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_first.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"], d["roofline"].get("by_bound"), d.get("view_images"))
except Exception as e:
    print("bench line unreadable:", e)
PY
