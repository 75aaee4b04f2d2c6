#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/s21_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/s21_tests.log
timeout 600 python bench.py --kf-policy vote --no-extras --no-cpu-baseline --steps 20 > gpurun_out/s21_bench_vote.json 2> gpurun_out/s21_bench_vote.err
echo "bench vote rc=$?"; tail -1 gpurun_out/s21_bench_vote.err | cut -c1-400
