#!/bin/bash
# One GPU round trip: bring-up diagnostics, tests, bench.  Usage (from the repo root): gpurun -- bash scripts/gpu_round.sh [stage...]
set -u
mkdir -p gpurun_out
STAGES="${@:-debug kernels parity smoke bench}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for s in $STAGES; do
  case $s in
    debug)   timeout 300 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "debug rc=$?" ;;
    attn)    timeout 240 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k attention --timeout 60 > gpurun_out/attn.log 2>&1; echo "attn rc=$?"; tail -n 12 gpurun_out/attn.log ;;
    kernels) timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?" ;;
    kernels_all) timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu --timeout 300 > gpurun_out/kernels.log 2>&1; echo "kernels rc=$?" ;;
    parity)  timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu --timeout 600 > gpurun_out/parity.log 2>&1; echo "parity rc=$?" ;;
    alltests) timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log ;;
    smoke)   timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" ;;
    bench)   timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" ;;
    benchref) timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?" ;;
    probe)   for d in 0 1 2 3; do echo "== JIMM_GEMM_DEBUG=$d"; JIMM_GEMM_DEBUG=$d PERF_ONLY=1 MODES=2 timeout 120 python scripts/gpu_debug_gemm.py 2>&1 | grep perf; done > gpurun_out/probe.log 2>&1; echo "probe rc=$?" ;;
    ncu_src) timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 2 -o gpurun_out/prof_gemm_src -f env PERF_ONLY=1 MODES=2 python scripts/gpu_debug_gemm.py > gpurun_out/ncu_src.log 2>&1; echo "ncu_src rc=$?" ;;
    attnperf) timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; JIMM_ATTN_IMPL=flash timeout 200 python scripts/gpu_attn_perf.py >> gpurun_out/attn_perf.log 2>&1; echo "attnperf rc=$?"; cat gpurun_out/attn_perf.log ;;
    ncu_attn) NCU=1 timeout 600 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:attention_tc -s 3 -c 1 -o gpurun_out/prof_attn -f python scripts/gpu_attn_perf.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu_attn rc=$?" ;;
    pairprobe) for pm in 0 1; do echo "== JIMM_GEMM_PAIR=$pm"; JIMM_GEMM_PAIR=$pm MODES=2 timeout 200 python scripts/gpu_debug_gemm.py 2>&1 | grep -E "perf|WORST|BAD|best matches"; done > gpurun_out/pairprobe.log 2>&1; echo "pairprobe rc=$?"; cat gpurun_out/pairprobe.log ;;
    ncu_list) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_list.log 2>&1; echo "ncu_list rc=$?" ;;
    ncu_full) timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 20 -c 4 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" ;;
  esac
done
tail -n 30 gpurun_out/debug_gemm.log 2>/dev/null
tail -n 15 gpurun_out/kernels.log 2>/dev/null
tail -n 15 gpurun_out/parity.log 2>/dev/null
tail -n 5 gpurun_out/smoke.log 2>/dev/null
tail -n 3 gpurun_out/bench.log gpurun_out/bench.err 2>/dev/null
cat gpurun_out/probe.log 2>/dev/null
