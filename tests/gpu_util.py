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


# ---- parity bookkeeping: every parity assertion records (config, dtype, bound, achieved) so a pass at 1.9e-3 and a pass at
#      1e-5 do not look alike.  Records go to gpurun_out/parity_records.jsonl (merged back from the GPU box) and are turned into
#      the committed PARITY.md by scripts/make_parity_md.py.
import json
import os
import time

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARITY_LOG = os.environ.get("JIMM_PARITY_LOG", os.path.join(_ROOT, "gpurun_out", "parity_records.jsonl"))


def record_parity(case: str, what: str, dtype: str, against: str, bound, achieved: float, **extra):
    rec = dict(case=case, what=what, dtype=dtype, against=against, bound=bound, achieved=float(achieved), ok=bool(bound is None or achieved < bound),
               t=time.strftime("%Y-%m-%dT%H:%M:%S"), **extra)
    try:
        os.makedirs(os.path.dirname(PARITY_LOG), exist_ok=True)
        with open(PARITY_LOG, "a") as f:
            f.write(json.dumps(rec) + "\n")
    except OSError:
        pass
    return rec


def check_parity(case: str, what: str, dtype, against: str, out, ref, bound, **extra) -> float:
    """rel = max|out - ref| / max|ref|; recorded, then asserted against `bound` (None = report only)."""
    r = rel_err(torch.as_tensor(out).detach().cpu(), torch.as_tensor(ref).detach().cpu())
    dn = str(dtype).replace("torch.", "")
    record_parity(case, what, dn, against, bound, r, **extra)
    if bound is not None:
        assert r < bound, f"{case}/{what} [{dn}] vs {against}: rel err {r:.3e} >= bound {bound:.1e}"
    return r
