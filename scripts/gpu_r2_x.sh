#!/bin/bash
# uint8 frames through the dual towers: tests, N=2 bench (collective leg e2e from uint8), N=1 bench with extras
set -u
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_preprocess_gpu.py -q -m gpu --timeout 300 2>&1 | tail -n 3
SKIP_TESTS=1 bash scripts/gpu_r2_scale.sh
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu > gpurun_out/bench_x.log 2> gpurun_out/bench_x.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_x.log').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']))
for k,v in d['extra_workloads'].items(): print(k, {kk:(round(vv) if isinstance(vv,float) else vv) for kk,vv in v.items() if kk in ('value','error')}, {kk:(round(vv) if isinstance(vv,float) else vv) for kk,vv in v.get('e2e',{}).items() if kk in ('value','h2d_bytes_per_step','input')}, 'fp32', round(v.get('e2e',{}).get('fp32_input',{}).get('value',0)))
PY
tail -3 gpurun_out/bench_x.err
