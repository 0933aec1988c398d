#!/bin/bash
# bench.py --gpus N exactly as the driver launches it (N = number of visible GPUs), plus the multi-GPU tests at that world size
set -u
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
[ -n "${SKIP_TESTS:-}" ] || timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 > gpurun_out/multigpu_n$N.log 2>&1; echo "multigpu(N=$N) rc=$?"; tail -n 4 gpurun_out/multigpu_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench_n$N rc=$?"; python - <<PY
import json
for l in open('gpurun_out/bench_n$N.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'gemm frac',round(d['roofline']['frac'],3)); c=d['collective']; print({k:v for k,v in c.items() if k not in ('reference_step','note','kernel')})
PY
grep -v -i "warn\|^\*\|OMP_NUM" gpurun_out/bench_n$N.err | tail -5
