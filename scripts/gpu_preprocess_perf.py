"""Front-end kernel micro-benchmark (CUDA events): images/s and achieved HBM GB/s (algorithmic bytes = input + output)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import preprocess_oracle as P
from jimm_b200.preprocess import ImagePreprocessor

for kind, size, h, w, B in (("vit", 224, 480, 640, 256), ("clip", 224, 480, 640, 256), ("siglip", 256, 480, 640, 256),
                            ("vit", 224, 224, 224, 256), ("siglip", 512, 1080, 1920, 32), ("clip", 224, 333, 500, 256)):
    proc = getattr(ImagePreprocessor, kind)(size)
    x = torch.randint(0, 256, (B, h, w, 3), dtype=torch.uint8, device="cuda")
    for dt in (torch.float16,):
        for _ in range(3):
            out = proc(x, dtype=dt)
        torch.cuda.synchronize()
        reps = 3 if os.environ.get("NCU") else 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = proc(x, dtype=dt)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = x.numel() + out.numel() * out.element_size()
        print(f"{kind:6s} {h}x{w} -> {tuple(out.shape[1:3])} B={B} {str(dt)[6:]}: {ms*1e3:8.1f} us  {B/ms*1e3:10.0f} img/s  {by/ms/1e6:7.1f} GB/s", flush=True)
    if os.environ.get("NCU"):
        break
