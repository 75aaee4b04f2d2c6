#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s28_tests.log 2>&1
echo "tests rc=$?"; tail -3 gpurun_out/s28_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s28_smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/s28_smoke.log | cut -c1-250
SAN=/usr/local/cuda/bin/compute-sanitizer
timeout 600 $SAN --tool memcheck --print-limit 20 python __graft_entry__.py smoke > gpurun_out/s28_san_memcheck_smoke.log 2>&1
echo "memcheck smoke rc=$?"; grep -E "ERROR SUMMARY" gpurun_out/s28_san_memcheck_smoke.log | tail -2
timeout 900 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_pyramid.py -m gpu -q -k "fill_in or noise_and_empty or uint16 or colored" > gpurun_out/s28_san_memcheck_pyr.log 2>&1
echo "memcheck pyr rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/s28_san_memcheck_pyr.log | tail -3
timeout 900 $SAN --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_track.py -m gpu -q -k "fixed_iterations and 22 or batch_matches_single or quality_vote or batched_vote or vote_policy" > gpurun_out/s28_san_memcheck_track.log 2>&1
echo "memcheck track rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/s28_san_memcheck_track.log | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -s 230 -c 60 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 3 --warmup 3 --kf-interval 2 --no-cpu-baseline --no-extras --no-pipeline --streams 256 > gpurun_out/s28_launches.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2_launches_step.csv
timeout 900 python bench.py > gpurun_out/s28_bench.json 2> gpurun_out/s28_bench.err
echo "bench rc=$?"; tail -1 gpurun_out/s28_bench.err | cut -c1-400
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/s28_ref.json 2> gpurun_out/s28_ref.err
echo "ref rc=$?"; cut -c1-200 gpurun_out/s28_ref.json | tail -1
