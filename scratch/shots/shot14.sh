#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1200 python -m pytest tests/test_gpu_split.py -m gpu -q -s > gpurun_out/s14_split.log 2>&1
echo "rc=$?"; tail -6 gpurun_out/s14_split.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29700 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/s14_bench2.json 2> gpurun_out/s14_bench2.err
echo "rc=$?"; tail -3 gpurun_out/s14_bench2.err | cut -c1-300
