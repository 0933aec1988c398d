#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; cat gpurun_out/attn_perf.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?"; cut -c1-3000 gpurun_out/bench_n2.log; tail -3 gpurun_out/bench_n2.err
