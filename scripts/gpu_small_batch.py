"""Small-batch latency (sync per call, the launch-bound regime): CUDA-graph replay vs eager launches."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

def run(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

mode = "graph" if os.environ.get("JIMM_GRAPH_MAX_BATCH", "32") != "0" else "eager"
model, img_size, _ = bench.build_model("vit_b16", "float16")
for B in (1, 4, 8, 16, 32):
    x = torch.randn(B, img_size, img_size, 3, device="cuda")
    xh = x.cpu().pin_memory()
    print(f"vit_b16 fp16 B={B:3d} {mode}: device-in {run(lambda: model(x)):7.3f} ms   host-in {run(lambda: model(xh)):7.3f} ms", flush=True)
model, img_size, text = bench.build_model("clip_b32", "float16")
for B in (1, 8, 32):
    x = torch.randn(B, img_size, img_size, 3, device="cuda")
    ids = torch.randint(1, 1000, (B, text[0]), dtype=torch.int32, device="cuda")
    print(f"clip_b32 fp16 B={B:3d} {mode}: dual {run(lambda: model(x, ids)):7.3f} ms", flush=True)
