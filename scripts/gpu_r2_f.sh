#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_preprocess_gpu.py tests/test_loader_and_abi.py -q -m gpu --timeout 600 -x > gpurun_out/sub.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/sub.log
timeout 900 python scripts/gpu_load_time.py > gpurun_out/load_time_after.log 2>&1; echo "load(after) rc=$?"; grep -v Warning gpurun_out/load_time_after.log | tail -16
if [ -d _old_tree ]; then (cd _old_tree && timeout 900 python scripts/gpu_load_time.py > ../gpurun_out/load_time_before.log 2>&1; echo "load(before) rc=$?"); grep -v Warning gpurun_out/load_time_before.log | tail -16; fi
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value', round(d['value']), 'e2e', json.dumps(d['e2e']))
"
