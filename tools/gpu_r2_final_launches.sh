#!/usr/bin/env bash
# Round 2, closing build: ncu launch list (duration + DRAM bytes per launch) of one eager teacher+student step.
set -u
O=gpurun_out/r2w
mkdir -p $O
timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $O/launches.csv python tools/profile_step.py > $O/profile_step.log 2>&1; echo "ncu launch list rc=$?"
python tools/launch_summary.py $O/launches.csv > $O/launches_summary.txt 2>&1
head -40 $O/launches_summary.txt
