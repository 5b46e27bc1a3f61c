#!/usr/bin/env bash
# Round 2 final verification (1 GPU): suite, smoke, bench lines, sustained run, launch list + one ncu --set full capture.
set -u
O=gpurun_out/r2n
mkdir -p $O
python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "gpu suite rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 20 --warmup 5 > $O/bench_c1.json 2> $O/bench_c1.err; echo "bench c1 rc=$?"
python bench.py --workload configs3 --steps 20 --warmup 5 --no-view-bench > $O/bench_c3.json 2> $O/bench_c3.err; echo "bench c3 rc=$?"
python bench.py --workload configs4 --steps 40 --warmup 8 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; echo "bench c4 rc=$?"
python bench.py --steps 1000 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_c1_1000.json 2> $O/bench_c1_1000.err; echo "bench c1 x1000 rc=$?"
python bench.py --backbone rla --steps 20 --warmup 5 --no-cpu-baseline --no-view-bench --no-ncu-traffic > $O/bench_rla.json 2> $O/bench_rla.err; echo "bench rla rc=$?"
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "ncu launch list rc=$?"
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
ncu --profile-from-start off --set full --import-source on --clock-control none -f -o $O/tower_pair \
    python tools/ncu_cases.py tower_full > $O/ncu_tower.log 2>&1; echo "ncu tower rc=$?"
ncu -i $O/tower_pair.ncu-rep --page raw --csv > $O/tower_pair_raw.csv 2>/dev/null
tail -n 3 $O/t_all.log; tail -n 1 $O/smoke.log
for f in c1 c3 c4 c1_1000 rla; do python - "$O/bench_$f.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[1], {k: d.get(k) for k in ("value", "ms_per_step", "steady")}, (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("frac_of_sustained_peak"), (r.get("head_tower") or {}).get("tflops"), d.get("clocks"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1200:])
PY
done
head -8 $O/launches_summary.txt
