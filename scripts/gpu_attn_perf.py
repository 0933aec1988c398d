"""Attention kernel micro-benchmark (CUDA events) for the BASELINE shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gpu_util import F16, check, ptr, stream
from jimm_b200 import _lib

lib = _lib.load()
CASES = ((256, 197, 12, 0), (256, 256, 12, 0), (256, 50, 12, 0), (256, 77, 8, 1), (128, 576, 16, 0), (32, 1024, 16, 0))
if os.environ.get("ONLY_S"):
    CASES = tuple(c for c in CASES if c[1] == int(os.environ["ONLY_S"]))
for (B, S, H, causal) in CASES:
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda").half()
    out = torch.empty(B * S, H * 64, dtype=torch.float16, device="cuda")
    for _ in range(3):
        check(lib, lib.jimm_k_attention(ptr(qkv), F16, ptr(out), F16, B, S, H, causal, stream()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3 if os.environ.get("NCU") else 20
    e0.record()
    for _ in range(reps):
        check(lib, lib.jimm_k_attention(ptr(qkv), F16, ptr(out), F16, B, S, H, causal, stream()))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 4.0 * B * H * S * S * 64 * (0.5 if causal else 1.0)
    print(f"attention B={B} S={S} H={H} causal={causal} impl={os.environ.get('JIMM_ATTN_IMPL','tc')}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s", flush=True)
    if os.environ.get("NCU"):
        break
