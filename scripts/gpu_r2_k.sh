#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "attention" > gpurun_out/kernels.log 2>&1; echo "attention tests rc=$?"; tail -n 5 gpurun_out/kernels.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; cat gpurun_out/attn_perf.log
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/parity.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log | cut -c1-1800; tail -3 gpurun_out/bench.err
