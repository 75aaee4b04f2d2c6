#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_track -s 2 -c 1 -o gpurun_out/r2_k_track_v8 -f python scratch/track_bench.py --configs 0 --quick --reps 2 > gpurun_out/ncu_track.log 2>&1
tail -3 gpurun_out/ncu_track.log
ls -la gpurun_out/*.ncu-rep
