"""Bring-up diagnostic for the tcgen05 GEMM (run on the GPU box): compares against torch and the SIMT kernel and prints a
block-wise error map so descriptor / swizzle / TMEM-lane mistakes are visible in one round trip."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gpu_util import gemm, rel_err
from jimm_b200 import _lib

lib = _lib.load()
torch.manual_seed(0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def run(M, N, K, dtype, mode):
    A = torch.randn(M, K, device="cuda").to(dtype)
    B = (torch.randn(N, K, device="cuda") / K ** 0.5).to(dtype)
    if dtype == torch.float32:
        A = (A.view(torch.int32) & ~0x1FFF).view(torch.float32)
        B = (B.view(torch.int32) & ~0x1FFF).view(torch.float32)
    ref = A.double() @ B.double().T
    out = gemm(lib, A, B, mode=mode)
    torch.cuda.synchronize()
    e = rel_err(out, ref)
    print(f"M={M} N={N} K={K} {dtype} mode={mode}: rel err {e:.3e}" + (" BAD" if e > 1e-4 else ""), flush=True)
    if e > 1e-4:
        d = (out.double() - ref).abs()
        bm, bn = min(M, 32), min(N, 32)
        blk = d[: (M // bm) * bm, : (N // bn) * bn].reshape(M // bm, bm, N // bn, bn).amax(dim=(1, 3))
        print("block max-abs-error map (32x32 blocks):")
        print((blk > 1e-3).int()[:8, :16])
        print("out[0,:8]", out[0, :8].tolist())
        print("ref[0,:8]", ref[0, :8].tolist())
        # is it a permutation of columns / rows?
        for r in (0, 1, 8, 33):
            if r < M:
                best = (ref - out[r].double().unsqueeze(0)).abs().amax(dim=1).argmin().item()
                print(f"out row {r} best matches ref row {best}")
    return e


worst = 0.0
if not os.environ.get("PERF_ONLY"):
    for dtype in (torch.float16, torch.bfloat16, torch.float32):
        for (M, N, K) in ((128, 256, 64), (128, 256, 256), (256, 512, 768), (1000, 1000, 512), (1024, 512, 256), (3000, 1000, 512)):
            for mode in (0, 1, 2):
                worst = max(worst, run(M, N, K, dtype, mode))
    print("WORST", worst)

# quick perf probe of the big shapes (CUDA events, 10 reps)
def perf(M, N, K, mode, residual):
    A = torch.randn(M, K, device="cuda").half()
    B = torch.randn(N, K, device="cuda").half()
    bias = torch.randn(N, device="cuda")
    if residual:
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32)
        kw = dict(residual=out)
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.float16)
        kw = dict(out=out)
    for _ in range(3):
        gemm(lib, A, B, bias=bias, mode=mode, **kw)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gemm(lib, A, B, bias=bias, mode=mode, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"perf M={M} N={N} K={K} {'f32 residual' if residual else 'f16 out'} mode={mode}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)


modes = [int(x) for x in os.environ.get("MODES", "0,1,2").split(",")]
for (M, N, K, res) in ((50432, 2304, 768, False), (50432, 3072, 768, False), (50432, 768, 3072, True), (50432, 768, 768, True),
                       (73728, 4096, 1024, False), (73728, 1024, 4096, True)):
    for mode in modes:
        perf(M, N, K, mode, res)
