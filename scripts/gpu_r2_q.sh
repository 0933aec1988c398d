#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/kernels.log 2>&1; echo "kernel tests rc=$?"; tail -n 12 gpurun_out/kernels.log
timeout 1200 python -m pytest tests/test_parity_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -n 8 gpurun_out/parity.log
for f in 1 0; do JIMM_FUSE_LN=$f timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('fuse_ln=$f value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']), round(d['roofline']['frac'],3), 'share', round(d['roofline']['gemm_share_of_step'],3), 'launches', d['gpu_launches'])
"; done
timeout 300 python scripts/gpu_e2e_probe.py 2>&1 | grep " ms" | head -4
