#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 4 gpurun_out/kernels.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; cat gpurun_out/attn_perf.log
timeout 100 python scripts/gpu_kernel_driver.py patchify 20
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 5 gpurun_out/parity.log
for wl in vit_l16_map; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$wl', 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'gemm frac', round(d['roofline']['frac'],3))
"; done
