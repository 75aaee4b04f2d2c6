#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 12 --warmup 3 --no-extras --no-cpu-baseline"
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 $B > gpurun_out/s23_$name.json 2> gpurun_out/s23_$name.err
  python - <<PY
import json
for l in open('gpurun_out/s23_$name.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d['phase_ms_per_step']
        print('$name', 'value %.0f'%d['value'], 'step %.3f'%d['ms_per_step'], 'pyr %.3f kf %.3f track %.3f'%(p['pyramid'],p['keyframe'],p['track_kernel']), 'roof %.3f'%d['roofline']['frac'])
PY
}
run base REVO_DUMMY=1
run nofuse REVO_CANNY_NO_FUSE=1
run nms5 REVO_NMS_MINBLOCKS=5
run nms6 REVO_NMS_MINBLOCKS=6
run maxc60 REVO_TRACK_MAX_CLUSTERS=60
run maxc66 REVO_TRACK_MAX_CLUSTERS=66
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_track -s 5 -c 1 -o gpurun_out/r2_k_track_bench -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s23_ncu_track.log 2>&1
echo "ncu track rc=$?"
ncu -i gpurun_out/r2_k_track_bench.ncu-rep --page raw --csv > gpurun_out/r2_k_track_bench_raw.csv 2>/dev/null
rm -f gpurun_out/r2_k_track_bench.ncu-rep
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section Occupancy --section LaunchStats --section SchedulerStats --clock-control none -k regex:'k_gray|k_canny|k_group|k_edt|k_opt_struct|k_pyrdown|k_depth|k_hist|k_fill' -s 130 -c 40 -o gpurun_out/r2_pyr_kernels_b -f python bench.py --steps 3 --warmup 3 --kf-interval 2 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s23_ncu_pyr.log 2>&1
echo "ncu pyr rc=$?"
ncu -i gpurun_out/r2_pyr_kernels_b.ncu-rep --page raw --csv > gpurun_out/r2_pyr_kernels_b_raw.csv 2>/dev/null
rm -f gpurun_out/r2_pyr_kernels_b.ncu-rep
du -sh gpurun_out
