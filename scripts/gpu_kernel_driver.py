"""Launch ONE hot-path kernel a few times at its benchmark shape (ViT-B/16, B = 256 unless stated) -- the target of the per-kernel ncu
captures under profiles/ (scripts/gpu_ncu_evidence.sh) and of quick CUDA-event timings.

    python scripts/gpu_kernel_driver.py <name> [reps]
names: gemm_qkv gemm_fc1 gemm_fc2 gemm_out gemm_patch attn_197 attn_256 attn_50 attn_77c attn_576 attn_1024 layernorm patchify
       map_attention comm_logits
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from gpu_util import F16, F32, check, gemm, ptr, stream
from jimm_b200 import _lib

lib = _lib.load()
name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
T, D, Mlp = 256 * 197, 768, 3072
dev = "cuda"


def make():
    if name.startswith("gemm_"):
        M, N, K, kind = {"gemm_qkv": (T, 3 * D, D, "plain"), "gemm_fc1": (T, Mlp, D, "gelu"), "gemm_fc2": (T, D, Mlp, "res"),
                         "gemm_out": (T, D, D, "res"), "gemm_patch": (256 * 196, D, 768, "res")}[name]
        A = torch.randn(M, K, device=dev).half()
        B = (torch.randn(N, K, device=dev) / K ** 0.5).half()
        bias = torch.randn(N, device=dev)
        if kind == "res":
            out = torch.zeros(M, N, device=dev)
            return (lambda: gemm(lib, A, B, bias=bias, mode=2, residual=out)), 2.0 * M * N * K, (M * K + N * K) * 2 + M * N * 8
        out = torch.empty(M, N, device=dev, dtype=torch.float16)
        return (lambda: gemm(lib, A, B, bias=bias, mode=2, out=out, act=1 if kind == "gelu" else 0)), 2.0 * M * N * K, (M * K + N * K + M * N) * 2
    if name.startswith("attn_"):
        B_, S, H, causal = {"attn_197": (256, 197, 12, 0), "attn_256": (256, 256, 12, 0), "attn_50": (256, 50, 12, 0), "attn_77c": (256, 77, 8, 1),
                            "attn_576": (128, 576, 16, 0), "attn_1024": (32, 1024, 16, 0)}[name]
        qkv = torch.randn(B_ * S, 3 * H * 64, device=dev).half()
        out = torch.empty(B_ * S, H * 64, dtype=torch.float16, device=dev)
        return (lambda: check(lib, lib.jimm_k_attention(ptr(qkv), F16, ptr(out), F16, B_, S, H, causal, stream()))), \
            4.0 * B_ * H * S * S * 64 * (0.5 if causal else 1.0), qkv.numel() * 2 + out.numel() * 2
    if name == "layernorm":
        x = torch.randn(T, D, device=dev) * 3 + 1.5
        sc, bi = torch.randn(D, device=dev), torch.randn(D, device=dev)
        out = torch.empty(T, D, dtype=torch.float16, device=dev)
        return (lambda: check(lib, lib.jimm_k_layernorm(ptr(x), D, 1, 0, None, ptr(sc), ptr(bi), 1e-6, ptr(out), F16, D, T, D, stream()))), 0.0, T * D * 6
    if name == "patchify":
        x = torch.randn(256, 224, 224, 3, device=dev)
        out = torch.empty(256 * 196, 768, dtype=torch.float16, device=dev)
        return (lambda: check(lib, lib.jimm_k_patchify(ptr(x), F32, 256, 224, 224, 3, 16, ptr(out), F16, stream()))), 0.0, x.numel() * 4 + out.numel() * 2
    if name == "map_attention":  # ViT-L/16@384 MAP head: B = 128, S = 576, H = 16
        B_, S, H = 128, 576, 16
        q = torch.randn(H * 64, device=dev)
        kv = torch.randn(B_ * S, 2 * H * 64, device=dev).bfloat16()
        out = torch.empty(B_, H * 64, dtype=torch.bfloat16, device=dev)
        return (lambda: check(lib, lib.jimm_k_map_attention(ptr(q), ptr(kv), 2, ptr(out), 2, B_, S, H, stream()))), 0.0, kv.numel() * 2
    if name == "comm_logits":  # single rank (world = 1): the normalise / publish / logits phases without the peer stores
        from jimm_b200.models import CLIP

        m = CLIP(64, 1, 64, 16, 8, 64, 512, 8, 1, dtype=torch.float16)
        n = m.native(256, require=True)
        h = C.create_string_buffer(64)
        _lib.check(lib.jimm_comm_init(n.handle, 0, 1, 256, h))
        _lib.check(lib.jimm_comm_connect(n.handle, bytes(h.raw)))
        n._comm = (0, 1, 256)
        ie, te = torch.randn(256, 512, device=dev), torch.randn(256, 512, device=dev)
        keep.append((m, n))
        return (lambda: n.comm_logits(ie, te)), 2.0 * 256 * 256 * 512, 2 * 256 * 512 * 4 * 2
    if name == "logits_c5":  # one rank's [256, 2048] x 1024 logits row block of BASELINE configs[4] (the shared fp32 tile, logits_tile.cuh)
        Bi, Bt, E = 256, 2048, 1024
        a = torch.nn.functional.normalize(torch.randn(Bi, E, device=dev), dim=-1)
        b = torch.nn.functional.normalize(torch.randn(Bt, E, device=dev), dim=-1)
        sc = torch.tensor([2.0], device=dev)
        out = torch.empty(Bi, Bt, device=dev)
        return (lambda: check(lib, lib.jimm_k_logits(ptr(a), ptr(b), ptr(sc), None, ptr(out), Bi, Bt, E, Bt, stream()))), 2.0 * Bi * Bt * E, (Bi + Bt) * E * 4 + Bi * Bt * 4
    raise SystemExit(f"unknown kernel {name}")


keep = []
fn, flops, bytes_ = make()
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{name}: {ms*1e3:.1f} us/launch (CUDA events, {reps} reps)" + (f"  {flops/ms/1e9:.1f} TFLOP/s" if flops else "") + f"  {bytes_/ms/1e6:.0f} GB/s algorithmic", flush=True)
