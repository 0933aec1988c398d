#!/bin/bash
# ncu evidence for the rewritten fp32 logits tile (single-GPU logits kernel at the c5 row-block shape, and the fused head kernel)
set -u
KERNELS="logits_c5 comm_logits" bash scripts/gpu_ncu_evidence.sh
# fp32-pipe metrics (names vary by ncu version: this capture is allowed to fail)
timeout 300 ncu --metrics sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:logits_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_r2/logits_c5_fma.csv python scripts/gpu_kernel_driver.py logits_c5 1 > gpurun_out/ncu_r2/logits_c5_fma.log 2>&1; echo "fma capture rc=$?"; tail -n 3 gpurun_out/ncu_r2/logits_c5_fma.csv | cut -c1-600
