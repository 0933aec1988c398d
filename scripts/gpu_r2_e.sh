#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 8 gpurun_out/alltests.log
KERNELS="gemm_qkv layernorm comm_logits map_attention patchify" bash scripts/gpu_ncu_evidence.sh
head -3 gpurun_out/ncu_r2/gemm_qkv.csv | cut -c1-600
timeout 300 python scripts/gpu_e2e_probe.py 2>&1 | grep slices | head -2
