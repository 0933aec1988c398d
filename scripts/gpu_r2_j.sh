#!/bin/bash
set -u
mkdir -p gpurun_out
JIMM_BENCH_DEBUG=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>&1 | grep -E "bench debug|^\{" | cut -c1-400
JIMM_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 --no-collective 2>&1 | grep -E "bench debug|^\{" | cut -c1-600
