#!/bin/bash
# round 2, GPU call B: new decoupled attention kernel + relaxed/early accumulator release in the GEMM epilogue + uint8 host path
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?"; tail -n 15 gpurun_out/kernels.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; echo "attnperf rc=$?"; cat gpurun_out/attn_perf.log
JIMM_ATTN_IMPL=flash ONLY_S=50 timeout 100 python scripts/gpu_attn_perf.py; JIMM_ATTN_IMPL=flash ONLY_S=77 timeout 100 python scripts/gpu_attn_perf.py
timeout 300 python scripts/gpu_gemm_perf.py > gpurun_out/gemm_perf.log 2>&1; echo "gemm_perf rc=$?"; cat gpurun_out/gemm_perf.log
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_parity_gpu.py -q -m gpu -x --timeout 600 > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 15 gpurun_out/parity.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
