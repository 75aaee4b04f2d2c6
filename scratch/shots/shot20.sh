#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s20_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/s20_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s20_smoke.log 2>&1
echo "smoke rc=$?"; tail -3 gpurun_out/s20_smoke.log
timeout 900 python bench.py > gpurun_out/s20_bench.json 2> gpurun_out/s20_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/s20_bench.err | cut -c1-400
