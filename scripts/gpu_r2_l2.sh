#!/bin/bash
# residual stream resident in L2 (JIMM_L2_PERSIST=<MB>): step-time A/B on ViT-B/16 B=256
set -u
for mb in 0 80 0 80 48; do JIMM_L2_PERSIST=$mb timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('l2_persist_mb=$mb value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'gemm', round(d['roofline']['achieved']), round(d['roofline']['frac'],3), 'share', round(d['roofline']['gemm_share_of_step'],3))
"; done
