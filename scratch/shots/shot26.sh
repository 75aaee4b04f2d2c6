#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29720 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s26_bench2.json 2> gpurun_out/s26_bench2.err
echo "bench2 rc=$?"; tail -2 gpurun_out/s26_bench2.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_split.py -m gpu -q -k "2-" > gpurun_out/s26_split.log 2>&1
echo "split rc=$?"; tail -3 gpurun_out/s26_split.log
