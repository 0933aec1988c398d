#!/bin/bash
# final 1-GPU evidence: full -m gpu suite (parity records), per-kernel ncu CSVs + event timings, launch list of a bench run, default bench
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
bash scripts/gpu_ncu_evidence.sh > gpurun_out/ncu_evidence.log 2>&1; tail -n 20 gpurun_out/ncu_evidence.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_list.log 2>&1; echo "ncu_list rc=$?"
timeout 300 python scripts/gpu_gemm_perf.py > gpurun_out/gemm_perf.log 2>&1; cat gpurun_out/gemm_perf.log
timeout 200 python scripts/gpu_cublas_ref.py > gpurun_out/cublas.log 2>&1; grep "linear" gpurun_out/cublas.log
timeout 200 python scripts/gpu_attn_perf.py > gpurun_out/attn_perf.log 2>&1; cat gpurun_out/attn_perf.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"; cat gpurun_out/bench_ref.log
