#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 > gpurun_out/multigpu.log 2>&1; echo "multigpu rc=$?"; tail -n 12 gpurun_out/multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('value',round(d['value']),'e2e',round(d['e2e']['value'])); c=d['collective']; print({k:v for k,v in c.items() if k not in ('workload','reference_step','note','kernel')})
PY
tail -3 gpurun_out/bench_n2.err
