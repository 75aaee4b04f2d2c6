#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pyramid.py -m gpu -x -q > gpurun_out/s27_tests.log 2>&1
echo "tests rc=$?"; tail -2 gpurun_out/s27_tests.log
B="python bench.py --steps 12 --warmup 3 --no-extras --no-cpu-baseline"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/s27_$name.json 2> gpurun_out/s27_$name.err
  python - <<PY
import json
for l in open('gpurun_out/s27_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d['phase_ms_per_step']
        print('$name', 'value %.0f'%d['value'], 'step %.3f'%d['ms_per_step'], 'pyr %.3f kf %.3f track %.3f'%(p['pyramid'],p['keyframe'],p['track_kernel']), 'roof %.3f'%d['roofline']['frac'])
PY
}
run base REVO_DUMMY=1
run gray4 REVO_GRAY_NO16=1
run band15 REVO_HYST_BAND=15
run band20 REVO_HYST_BAND=20
run band24 REVO_HYST_BAND=24
