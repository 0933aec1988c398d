#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:attention_tc_kernel -s 2 -c 1 -o gpurun_out/prof_attn197 -f python scripts/gpu_kernel_driver.py attn_197 1 > gpurun_out/ncu_attn197.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:attention_tc_long -s 2 -c 1 -o gpurun_out/prof_attn1024 -f python scripts/gpu_kernel_driver.py attn_1024 1 > gpurun_out/ncu_attn1024.log 2>&1; echo "ncu rc=$?"
