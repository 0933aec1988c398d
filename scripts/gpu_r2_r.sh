#!/bin/bash
# column-split attention: correctness, per-shape timing, step time
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "attention" > gpurun_out/kernels_attn.log 2>&1; echo "attention tests rc=$?"; tail -n 6 gpurun_out/kernels_attn.log
timeout 600 python scripts/gpu_attn_perf.py 2>&1 | tail -n 14
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']), round(d['roofline']['frac'],3), 'share', round(d['roofline']['gemm_share_of_step'],3))
"
