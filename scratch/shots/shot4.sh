#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest track" ; date
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_parity_population.py -m gpu -q -x > gpurun_out/s7_pytest.log 2>&1
echo "rc=$?"; tail -4 gpurun_out/s7_pytest.log
echo "== track_bench" ; date
timeout 600 python scratch/track_bench.py --quick > gpurun_out/s7_track_bench.log 2>&1
cat gpurun_out/s7_track_bench.log | tail -12
REVO_TRACK_PROF=1 timeout 300 python scratch/track_bench.py --configs 0,1 --quick --reps 2 > gpurun_out/s7_track_prof.log 2>&1
grep prof gpurun_out/s7_track_prof.log | tail -8
date
