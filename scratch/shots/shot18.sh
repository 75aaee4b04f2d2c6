#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 230 -c 56 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 3 --warmup 3 --kf-interval 2 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s18.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/s18.log | cut -c1-200
wc -l gpurun_out/r2_launches_step.csv
