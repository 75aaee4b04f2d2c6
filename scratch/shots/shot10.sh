#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_parity_population.py tests/test_gpu_pyramid.py -m gpu -q -x > gpurun_out/s10_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/s10_pytest.log
timeout 600 python scratch/track_bench.py --quick --configs 0,2,3,5,6,9,10,11 > gpurun_out/s10_track_bench.log 2>&1
cat gpurun_out/s10_track_bench.log | tail -9
REVO_TRACK_PROF=1 timeout 300 python scratch/track_bench.py --configs 0 --quick --reps 2 2>&1 | grep prof | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s10_bench.json 2> gpurun_out/s10_bench.err
tail -1 gpurun_out/s10_bench.err
