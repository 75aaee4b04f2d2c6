#!/bin/bash
set -u
mkdir -p gpurun_out
# level-0 launches of the Canny kernels in steady state (grid (5,4,256) / 512-thread CTAs)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_canny_nms|k_canny_hyst_smem|k_canny_expand|k_group_mask|k_group_scatter' -s 60 -c 20 -o gpurun_out/r2_pyr_kernels -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/ncu_pyr.log 2>&1
tail -2 gpurun_out/ncu_pyr.log | cut -c1-150
ls -la gpurun_out/r2_pyr_kernels.ncu-rep
