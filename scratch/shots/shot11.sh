#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scratch/track_bench.py --quick > gpurun_out/s11_track_bench.log 2>&1
cat gpurun_out/s11_track_bench.log | tail -16
