#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_track.py tests/test_gpu_parity_population.py -m gpu -q -x > gpurun_out/s9_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/s9_pytest.log
timeout 600 python scratch/track_bench.py --quick --configs 0,1,2,3,5,8 > gpurun_out/s9_track_bench.log 2>&1
cat gpurun_out/s9_track_bench.log | tail -8
REVO_TRACK_PROF=1 timeout 300 python scratch/track_bench.py --configs 0 --quick --reps 2 2>&1 | grep prof | tail -2
REVO_TRACK_HINT=2 timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_hint2.json 2> gpurun_out/s9_bench_hint2.err
tail -1 gpurun_out/s9_bench_hint2.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
tail -1 gpurun_out/s9_bench.err
