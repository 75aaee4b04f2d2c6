#!/bin/bash
# Round-2 first GPU shot: full GPU test suite (incl. the never-validated tests), engine A/B, compute-sanitizer, baseline bench.
set -u
mkdir -p gpurun_out
export REVO_RUN_UNVALIDATED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/s1_smi.txt 2>&1
echo "== pytest" ; date
timeout 1500 python -m pytest tests -m gpu -q -rs --durations=15 > gpurun_out/s1_pytest.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/s1_pytest.log
tail -5 gpurun_out/s1_pytest.log
echo "== track_bench" ; date
timeout 600 python scratch/track_bench.py --configs 0,1,6,7,8 --quick > gpurun_out/s1_track_bench.log 2>&1
cat gpurun_out/s1_track_bench.log | tail -8
REVO_TRACK_PROF=1 timeout 300 python scratch/track_bench.py --configs 0,6,7 --quick --reps 2 > gpurun_out/s1_track_prof.log 2>&1
grep prof gpurun_out/s1_track_prof.log | tail -6
echo "== smoke" ; date
timeout 600 python __graft_entry__.py smoke > gpurun_out/s1_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/s1_smoke.log
tail -3 gpurun_out/s1_smoke.log
echo "== sanitizer" ; date
SAN=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  timeout 900 $SAN --tool $tool --print-limit 20 python __graft_entry__.py smoke > gpurun_out/s1_san_${tool}_smoke.log 2>&1
  echo "$tool smoke rc=$?" | tee -a gpurun_out/s1_san_${tool}_smoke.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke:" gpurun_out/s1_san_${tool}_smoke.log | tail -4
  REVO_ASSUME_GPU=1 timeout 900 $SAN --tool $tool --print-limit 20 python -m pytest tests/test_gpu_pyramid.py -m gpu -q -k "fill_in or noise_and_empty or uint16" > gpurun_out/s1_san_${tool}_pyr.log 2>&1
  echo "$tool pyr rc=$?" | tee -a gpurun_out/s1_san_${tool}_pyr.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s1_san_${tool}_pyr.log | tail -4
  REVO_ASSUME_GPU=1 timeout 900 $SAN --tool $tool --print-limit 20 python -m pytest tests/test_gpu_track.py -m gpu -q -k "fixed_iterations and 22 or batch_matches_single or quality_vote" > gpurun_out/s1_san_${tool}_track.log 2>&1
  echo "$tool track rc=$?" | tee -a gpurun_out/s1_san_${tool}_track.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/s1_san_${tool}_track.log | tail -4
done
echo "== bench" ; date
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err
tail -2 gpurun_out/s1_bench.err
date
