#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scratch/track_bench.py --quick --configs 0,4,8,9,10,11,12 > gpurun_out/s8_track_bench.log 2>&1
cat gpurun_out/s8_track_bench.log | tail -8
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
tail -2 gpurun_out/s8_bench.err
