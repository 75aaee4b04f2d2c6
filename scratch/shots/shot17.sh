#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/s17_pytest.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/s17_pytest.log
for v in "REVO_PYR_SERIAL=1" ""; do
  echo "== $v"
  env $v timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | tail -1 | cut -c1-260
done
