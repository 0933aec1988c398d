#!/bin/bash
# step-time A/B: default launches vs whole-tower CUDA-graph replay at B = 256
set -u
for g in 32 256 32 256; do JIMM_GRAPH_MAX_BATCH=$g timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('graph_max_batch=$g value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
"; done
