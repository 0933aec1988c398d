#!/bin/bash
# round 2, GPU call C: parity + front-end tests on the new kernels, default bench (uint8 e2e, extra workloads)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_preprocess_gpu.py tests/test_parity_gpu.py tests/test_loader_and_abi.py -q -m gpu --timeout 600 > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 12 gpurun_out/parity.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
