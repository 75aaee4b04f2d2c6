#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s24_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/s24_tests.log
B="python bench.py --steps 12 --warmup 3 --no-extras --no-cpu-baseline"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/s24_$name.json 2> gpurun_out/s24_$name.err
  python - <<PY
import json
for l in open('gpurun_out/s24_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d['phase_ms_per_step']
        print('$name', 'value %.0f'%d['value'], 'step %.3f'%d['ms_per_step'], 'pyr %.3f kf %.3f track %.3f'%(p['pyramid'],p['keyframe'],p['track_kernel']), 'roof %.3f'%d['roofline']['frac'], 'kfprom %.3f'%d['roofline_pyramid']['keyframe']['ms_per_promotion'])
PY
}
run base REVO_DUMMY=1
run edtnofuse REVO_EDT_NO_FUSE=1
run maxc52 REVO_TRACK_MAX_CLUSTERS=52
run maxc56 REVO_TRACK_MAX_CLUSTERS=56
run maxc60 REVO_TRACK_MAX_CLUSTERS=60
run maxc64 REVO_TRACK_MAX_CLUSTERS=64
