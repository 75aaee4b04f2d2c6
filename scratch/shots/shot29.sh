#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29730 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s29_bench8.json 2> gpurun_out/s29_bench8.err
echo "bench8 rc=$?"; tail -1 gpurun_out/s29_bench8.err | cut -c1-300
