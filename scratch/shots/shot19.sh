#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29710 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/s19_bench8.json 2> gpurun_out/s19_bench8.err
echo "bench8 rc=$?"; tail -2 gpurun_out/s19_bench8.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/s19_bench4.json 2> gpurun_out/s19_bench4.err
echo "bench4 rc=$?"; tail -1 gpurun_out/s19_bench4.err | cut -c1-300
timeout 900 python -m pytest tests/test_gpu_split.py -m gpu -q -s -k "1920" > gpurun_out/s19_split.log 2>&1
echo "split rc=$?"; tail -5 gpurun_out/s19_split.log
