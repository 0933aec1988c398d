"""Helpers for the -m gpu tests: call the per-kernel C-ABI entry points on torch CUDA tensors."""

import ctypes as C
import math

import torch

F32, F16, BF16 = 0, 1, 2
TORCH = {F32: torch.float32, F16: torch.float16, BF16: torch.bfloat16}
CODE = {v: k for k, v in TORCH.items()}


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def check(lib, rc):
    assert rc == 0, f"rc={rc}: {lib.jimm_last_error().decode()}"


def gemm(lib, A, Bw, *, bias=None, act=0, rowadd=None, residual=None, out_dtype=torch.float32, rows=None, impl=0, mode=0, out=None):
    """C = epi(A @ Bw.T); rows=(rows_in, rows_out, row_off) remaps output rows.  Returns the output tensor."""
    M, K = A.shape
    N = Bw.shape[0]
    rows_in, rows_out, row_off = rows if rows else (0, 0, 0)
    if out is None:
        n_out_rows = M if not rows else (M // rows_in) * rows_out
        out = torch.zeros((n_out_rows, N), dtype=out_dtype, device=A.device) if residual is None else residual
    rc = lib.jimm_k_gemm(impl, CODE[A.dtype], ptr(A), A.stride(0), ptr(Bw), Bw.stride(0), M, N, K, ptr(bias), act, ptr(rowadd),
                         ptr(residual), 0 if residual is None else residual.stride(0), ptr(out), CODE[out.dtype], out.stride(0),
                         rows_in, rows_out, row_off, mode, stream())
    check(lib, rc)
    return out


def gelu_tanh(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def rel_err(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))
