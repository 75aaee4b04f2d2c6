#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 20 --warmup 3 > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
echo "rc=$?"; tail -5 gpurun_out/s13_bench.err | cut -c1-400
