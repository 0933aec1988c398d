#!/bin/bash
# attention A/B: one-thread-per-row kernel (default) vs the two-threads-per-row variant (JIMM_ATTN_SPLIT=1), plus phase probes of the variant
# (JIMM_ATC_DEBUG: 1 = no ex2, 2 = only first score chunk loaded, 4 = no O read-out / store, 8 = cycle trace of CTA 0)
set -u
echo "== default kernel"; timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
echo "== split kernel"; JIMM_ATTN_SPLIT=1 timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
for s in 197 256; do for d in 1 2 4 3 7; do echo -n "split dbg=$d "; JIMM_ATTN_SPLIT=1 JIMM_ATC_DEBUG=$d ONLY_S=$s timeout 120 python scripts/gpu_attn_perf.py 2>&1 | tail -n 1; done; done
for s in 197 256; do JIMM_ATTN_SPLIT=1 JIMM_ATC_DEBUG=8 ONLY_S=$s timeout 120 python scripts/gpu_attn_perf.py 2>&1 | grep -i "atc \|attention" | sort | uniq -c | sort -rn | head -12; done
JIMM_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "attention" 2>&1 | tail -n 2
