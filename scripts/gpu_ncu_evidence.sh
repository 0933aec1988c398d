#!/bin/bash
# Per-kernel ncu evidence: one CSV per kernel under gpurun_out/ncu_r2/ with the metric NAMES in the header, plus the
# CUDA-event timing of the same driver without the profiler.  scripts/make_profiles_r2.py condenses them into profiles/r2_kernels.md.
set -u
OUT=gpurun_out/ncu_r2
mkdir -p $OUT
M="gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,l1tex__m_xbar2l1tex_read_bytes.sum,l1tex__m_l1tex2xbar_write_bytes.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,launch__grid_size,launch__block_size"
: > $OUT/timings.txt
for k in ${KERNELS:-gemm_qkv gemm_fc1 gemm_fc2 gemm_out gemm_patch attn_197 attn_256 attn_50 attn_77c attn_576 attn_1024 layernorm patchify map_attention comm_logits}; do
  timeout 120 python scripts/gpu_kernel_driver.py $k 20 >> $OUT/timings.txt 2>&1
  case $k in
    gemm_*) pat="regex:gemm_tcgen05" ;;
    attn_576|attn_1024) pat="regex:attention_tc_long" ;;
    attn_*) pat="regex:attention_tc_kernel" ;;
    layernorm) pat="regex:layernorm_kernel" ;;
    patchify) pat="regex:patchify" ;;
    map_attention) pat="regex:map_attention" ;;
    comm_logits) pat="regex:comm_logits" ;;
    logits_c5) pat="regex:logits_kernel" ;;
  esac
  timeout 300 ncu --metrics $M --clock-control none -k $pat -s 2 -c 1 --csv --log-file $OUT/$k.csv python scripts/gpu_kernel_driver.py $k 1 > $OUT/$k.log 2>&1
  echo "$k rc=$?"
done
cat $OUT/timings.txt
