#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pyramid.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/s25_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/s25_tests.log
for c in 0 57 58 59 60 61 62 60 0; do
  timeout 300 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline --track-max-clusters $c > gpurun_out/s25_c$c.json 2> gpurun_out/s25_c$c.err
  python - <<PY
import json
for l in open('gpurun_out/s25_c$c.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d['phase_ms_per_step']
        print('cap $c', 'value %.0f'%d['value'], 'step %.3f'%d['ms_per_step'], 'e2e %.0f'%d['e2e']['value'], 'pyr %.3f kf %.3f track %.3f'%(p['pyramid'],p['keyframe'],p['track_kernel']), 'roof %.3f'%d['roofline']['frac'], 'kfprom %.3f'%d['roofline_pyramid']['keyframe']['ms_per_promotion'])
PY
done
