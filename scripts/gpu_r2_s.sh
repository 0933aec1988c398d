#!/bin/bash
# (1) attention A/B: one-thread-per-row kernel (default) vs the two-threads-per-row variant (JIMM_ATTN_SPLIT=1), phase probes of the variant
#     (JIMM_ATC_DEBUG: 1 = no ex2, 2 = only first score chunk loaded, 4 = no O read-out / store, 8 = cycle trace of CTA 0)
# (2) end-of-round validation of the tree: full -m gpu suite, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
{
echo "== default kernel"; timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
echo "== split kernel"; JIMM_ATTN_SPLIT=1 timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
for s in 197 256; do for d in 1 2 4 3 7; do echo -n "split dbg=$d "; JIMM_ATTN_SPLIT=1 JIMM_ATC_DEBUG=$d ONLY_S=$s timeout 120 python scripts/gpu_attn_perf.py 2>&1 | tail -n 1; done; done
for s in 197 256; do JIMM_ATTN_SPLIT=1 JIMM_ATC_DEBUG=8 ONLY_S=$s timeout 120 python scripts/gpu_attn_perf.py 2>&1 | grep -i "atc \|attention" | sort | uniq -c | sort -rn | head -12; done
JIMM_ATTN_SPLIT=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "attention" 2>&1 | tail -n 2
} 2>&1 | tee gpurun_out/attn_ab.log
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"; cat gpurun_out/bench_ref.log
