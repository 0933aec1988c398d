#!/bin/bash
# round 2, GPU call D (2 GPUs): sharded contrastive head tests (bounded wait, B_local check), bench --gpus 2 with the collective leg,
# dual-stream towers A/B, e2e slice probe
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_parity_gpu.py -q -m gpu --timeout 600 -k "sharded or dual or clip or siglip" > gpurun_out/multigpu.log 2>&1; echo "multigpu rc=$?"; tail -n 12 gpurun_out/multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench_n2 rc=$?"; cat gpurun_out/bench_n2.log; tail -5 gpurun_out/bench_n2.err
for ds in 1 0; do for wl in clip_b32 siglip_b16; do JIMM_DUAL_STREAMS=$ds timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-extras 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('dual_streams=$ds', '$wl', 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))
"; done; done
timeout 600 python scripts/gpu_e2e_probe.py 2>&1 | grep slices
