#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 300 python scripts/gpu_gemm_perf.py > gpurun_out/gemm_perf.log 2>&1; cat gpurun_out/gemm_perf.log
JIMM_GEMM_TAIL_SPLIT=0 timeout 300 python scripts/gpu_gemm_perf.py 2>&1 | grep clip
for wl in clip_b32; do for ts in 1 0; do JIMM_GEMM_TAIL_SPLIT=$ts timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('tail_split=$ts', '$wl', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
"; done; done
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
