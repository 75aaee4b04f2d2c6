#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/s31_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/s31_tests.log
timeout 120 python bench.py --steps 12 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/s31_bench.json 2> gpurun_out/s31_bench.err
echo "bench rc=$?"; tail -1 gpurun_out/s31_bench.err | cut -c1-300
python - <<PY
import json
for l in open('gpurun_out/s31_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('kfprom %.3f'%d['roofline_pyramid']['keyframe']['ms_per_promotion'], 'value %.0f'%d['value'])
PY
