#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29740 bench.py --gpus 2 --steps 20 --warmup 3 --no-extras > gpurun_out/s30_bench2.json 2> gpurun_out/s30_bench2.err
echo "bench2 rc=$?"; tail -1 gpurun_out/s30_bench2.err | cut -c1-300
