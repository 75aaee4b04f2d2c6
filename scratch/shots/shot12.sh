#!/bin/bash
set -u
mkdir -p gpurun_out
for B in 148 222 256 296; do
  echo "== B=$B"
  timeout 600 python scratch/track_bench.py --quick --configs 0,4 --streams $B 2>&1 | tail -2
done
