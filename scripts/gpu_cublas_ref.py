"""cuBLAS (torch.matmul / F.linear) timing on the bench GEMM shapes -- a practical ceiling for comparison only (never on the product path)."""
import torch
import torch.nn.functional as F

for (M, N, K) in ((50432, 2304, 768), (50432, 3072, 768), (50432, 768, 3072), (50432, 768, 768), (73728, 4096, 1024), (73728, 1024, 4096), (8192, 8192, 8192)):
    A = torch.randn(M, K, device="cuda").half()
    W = torch.randn(N, K, device="cuda").half()
    b = torch.randn(N, device="cuda").half()
    for name, fn in (("matmul", lambda: A @ W.T), ("linear+bias", lambda: F.linear(A, W, b))):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"cublas {name} M={M} N={N} K={K}: {ms:.3f} ms {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
