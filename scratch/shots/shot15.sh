#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_vs_reference.py -m gpu -q -s > gpurun_out/s15_ref.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/s15_ref.log | cut -c1-200
