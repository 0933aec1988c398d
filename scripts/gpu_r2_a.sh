#!/bin/bash
# round 2, GPU call A: full -m gpu suite (records achieved parity errors), isolated kernel perf tables, FC1 GEMM ncu capture
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 8 gpurun_out/alltests.log
timeout 300 python scripts/gpu_gemm_perf.py > gpurun_out/gemm_perf.log 2>&1; echo "gemm_perf rc=$?"; cat gpurun_out/gemm_perf.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; echo "attnperf rc=$?"; cat gpurun_out/attn_perf.log
timeout 200 python scripts/gpu_cublas_ref.py > gpurun_out/cublas.log 2>&1; tail -n 14 gpurun_out/cublas.log
ONLY=fc1 REPS=2 timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -o gpurun_out/prof_fc1 -f python scripts/gpu_gemm_perf.py > gpurun_out/ncu_fc1.log 2>&1; echo "ncu_fc1 rc=$?"
ONLY=qkv REPS=2 timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 1 -o gpurun_out/prof_qkv -f python scripts/gpu_gemm_perf.py > gpurun_out/ncu_qkv.log 2>&1; echo "ncu_qkv rc=$?"
