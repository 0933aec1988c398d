#!/bin/bash
# 2-GPU box: (1) dry-run of the c5 collective leg (SigLIP2-L/16@512, bf16) at world 2 through bench.py as the driver launches it;
# (2) on one GPU: attention issuer A/B (fixed alternation vs readiness polling), the split-variant test
set -u
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
JIMM_BENCH_COLLECTIVE_WL=siglip2_l16_512 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_c5dry_n$N.log 2> gpurun_out/bench_c5dry_n$N.err; echo "bench_c5dry rc=$?"; python - <<PY
import json
for l in open('gpurun_out/bench_c5dry_n$N.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N',d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],3)); c=d['collective']; print({k:v for k,v in c.items() if k not in ('reference_step','note','kernel')})
PY
grep -v -i "warn\|^\*\|OMP_NUM" gpurun_out/bench_c5dry_n$N.err | tail -5
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu --timeout 600 2>&1 | tail -n 3
export CUDA_VISIBLE_DEVICES=0
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu --timeout 600 -k "logits or contrastive or clip or siglip" 2>&1 | tail -n 3
python - <<'PY'
import ctypes as C, sys, os
sys.path.insert(0, 'tests')
import torch
from gpu_util import check, ptr, stream
from jimm_b200 import _lib
lib = _lib.load()
for (Bi, Bt, E) in ((256, 2048, 1024), (256, 1024, 512), (256, 256, 768)):
    a = torch.nn.functional.normalize(torch.randn(Bi, E, device='cuda'), dim=-1); b = torch.nn.functional.normalize(torch.randn(Bt, E, device='cuda'), dim=-1)
    sc = torch.tensor([2.0], device='cuda'); out = torch.empty(Bi, Bt, device='cuda')
    for _ in range(3): check(lib, lib.jimm_k_logits(ptr(a), ptr(b), ptr(sc), None, ptr(out), Bi, Bt, E, Bt, stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): check(lib, lib.jimm_k_logits(ptr(a), ptr(b), ptr(sc), None, ptr(out), Bi, Bt, E, Bt, stream()))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    ref = (a.double() @ b.double().T) * float(torch.exp(sc.double()))
    print(f"logits [{Bi},{Bt}]x{E}: {us:.1f} us  {2*Bi*Bt*E/us/1e6:.2f} TFLOP/s fp32  max err {float((out.double()-ref).abs().max()):.2e}")
PY
echo "== fixed alternation"; timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
echo "== readiness polling"; JIMM_ATC_POLL=1 timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -4
echo "== fixed alternation (again)"; timeout 300 python scripts/gpu_attn_perf.py 2>&1 | grep attention | head -2
JIMM_ATC_POLL=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300 -k "attention" 2>&1 | tail -n 2
