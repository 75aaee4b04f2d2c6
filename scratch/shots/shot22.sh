#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_track -s 5 -c 1 -o gpurun_out/r2_k_track_bench -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s22_ncu_track.log 2>&1
echo "ncu track rc=$?"; tail -1 gpurun_out/s22_ncu_track.log | cut -c1-200
timeout 900 ncu --set full --clock-control none -k regex:'k_gray|k_canny|k_group|k_edt|k_opt_struct|k_pyrdown|k_depth|k_hist|k_fill' -s 130 -c 48 -o gpurun_out/r2_pyr_kernels_b -f python bench.py --steps 3 --warmup 3 --kf-interval 2 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s22_ncu_pyr.log 2>&1
echo "ncu pyr rc=$?"; tail -1 gpurun_out/s22_ncu_pyr.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -3
