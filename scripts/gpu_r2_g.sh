#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/gpu_e2e_probe.py 2>&1 | grep slices
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/gpu_e2e_ranks.py 2>&1 | grep "^rank"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'e2e',json.dumps(d['e2e'])); c=d['collective']; print({k:c[k] for k in ('value','us_per_call','nccl_allgather_us_per_call','bit_identical_to_single_gpu','nvlink_egress_gbs')}, c['e2e'])
PY
tail -3 gpurun_out/bench_n2.err
