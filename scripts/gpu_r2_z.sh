#!/bin/bash
# last check of the final tree: full -m gpu suite + smoke (the bench lines are already recorded under profiles/)
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_records.jsonl
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/alltests.log 2>&1; echo "alltests rc=$?"; tail -n 6 gpurun_out/alltests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log
