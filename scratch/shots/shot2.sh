#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== population" ; date
timeout 900 python -m pytest tests/test_gpu_parity_population.py -m gpu -q -s > gpurun_out/s2_population.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/s2_population.log
echo "== track_bench" ; date
timeout 600 python scratch/track_bench.py --configs 0,6,8,9,10,11,12,13,14 --quick > gpurun_out/s2_track_bench.log 2>&1
cat gpurun_out/s2_track_bench.log | tail -12
date
