"""tcgen05 GEMM micro-benchmark on the encoder-block shapes with their REAL fused epilogues (CUDA events, isolated launches).

    python scripts/gpu_gemm_perf.py            # table for ViT-B/16 (T = 50432), ViT-L/16@384 (T = 73728), CLIP-B/32 towers
    ONLY=fc1 REPS=3 python scripts/gpu_gemm_perf.py   # one shape (for ncu)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gpu_util import gemm
from jimm_b200 import _lib

lib = _lib.load()
torch.manual_seed(0)
REPS = int(os.environ.get("REPS", "20"))
ONLY = os.environ.get("ONLY")
DT = torch.bfloat16 if os.environ.get("DT") == "bf16" else torch.float16


def perf(tag, M, N, K, kind):
    A = torch.randn(M, K, device="cuda").to(DT)
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).to(DT)
    bias = torch.randn(N, device="cuda")
    if kind == "res":
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        kw = dict(residual=out)
    else:
        out = torch.empty(M, N, device="cuda", dtype=DT)
        kw = dict(out=out, act={"plain": 0, "gelu": 1, "qgelu": 2}[kind])
    for _ in range(3):
        gemm(lib, A, B, bias=bias, mode=2, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(REPS):
        gemm(lib, A, B, bias=bias, mode=2, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / REPS
    print(f"gemm {tag:10s} M={M} N={N} K={K} {kind:5s}: {ms*1e3:7.1f} us  {2*M*N*K/ms/1e9:7.1f} TFLOP/s", flush=True)


CASES = [
    ("qkv", 50432, 2304, 768, "plain"), ("fc1", 50432, 3072, 768, "gelu"), ("fc2", 50432, 768, 3072, "res"), ("out", 50432, 768, 768, "res"),
    ("L.qkv", 73728, 3072, 1024, "plain"), ("L.fc1", 73728, 4096, 1024, "gelu"), ("L.fc2", 73728, 1024, 4096, "res"), ("L.out", 73728, 1024, 1024, "res"),
    ("clipv.qkv", 12800, 2304, 768, "plain"), ("clipv.fc1", 12800, 3072, 768, "qgelu"), ("clipv.fc2", 12800, 768, 3072, "res"), ("clipv.out", 12800, 768, 768, "res"),
    ("clipt.qkv", 19712, 1536, 512, "plain"), ("clipt.fc1", 19712, 2048, 512, "qgelu"), ("clipt.fc2", 19712, 512, 2048, "res"), ("clipt.out", 19712, 512, 512, "res"),
]
for c in CASES:
    if ONLY and c[0] != ONLY:
        continue
    perf(*c)
