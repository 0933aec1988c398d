"""Per-image device time of the ViT-B/16 fp16 tower against the batch size (wave quantisation and L2 residency)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

model, img_size, _ = bench.build_model(sys.argv[1] if len(sys.argv) > 1 else "vit_b16", "float16")
model.set_max_batch(256)
img = torch.randn(256, img_size, img_size, 3, device="cuda")
for _ in range(3): model(img)
torch.cuda.synchronize()
def run(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ns = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [31, 48, 62, 93, 96, 124, 127, 155, 186, 192, 217, 248, 256]
for n in ns:
    x = img[:n]
    ms = run(lambda: model(x))
    print(f"n={n:4d}  {ms:7.3f} ms  {ms / n * 1e3:6.2f} us/img", flush=True)
