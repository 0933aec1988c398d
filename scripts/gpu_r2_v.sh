#!/bin/bash
# end-of-round validation of the final tree: full -m gpu suite, smoke, default bench, reference arm
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "benchref rc=$?"; cut -c1-300 gpurun_out/bench_ref.log
