"""Experiment: e2e host path (pinned H2D + forward + D2H) of ViT-B/16 B=256 under different slice schedules.
Usage (GPU box): python scripts/gpu_host_path.py"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

def main():
    dev = torch.device("cuda:0")
    model, img_size, _ = bench.build_model("vit_b16", "float16")
    B = 256
    model.set_max_batch(B)
    img_host = torch.randn(B, img_size, img_size, 3).pin_memory()
    img_dev = img_host.to(dev)
    for _ in range(5):
        model(img_dev)
    torch.cuda.synchronize()
    def run(fn, n=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    print("device", round(run(lambda: model(img_dev)), 3), "ms")
    # raw copies
    print("H2D full", round(run(lambda: img_dev.copy_(img_host, non_blocking=True)), 3), "ms")
    nat = model.native(B)
    for sl in ["256", "64,192", "62,194", "33,223", "48,208", "32,64,160"]:
        for d2h in ["0"]:
            os.environ["JIMM_HOST_SLICES"] = sl
            print("slices", sl, round(run(lambda: model(img_host)), 3), "ms", flush=True)
    os.environ.pop("JIMM_HOST_SLICES")
    print("default", round(run(lambda: model(img_host)), 3), "ms")
    # old runtime behaviour: fresh pinned tensor per call, no clone
    out = torch.empty((B, nat.vision_out), dtype=torch.float32, pin_memory=True)
    lib = nat.lib
    from jimm_b200 import _lib
    def raw():
        _lib.check(lib.jimm_vit_forward_host(nat.handle, C.c_void_p(img_host.data_ptr()), 0, B, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.current_stream().synchronize()
    print("raw C call default", round(run(raw), 3), "ms")
    # device-only forward of 64 and 192 to see slice efficiency
    for n in (32, 33, 62, 64, 192, 194, 224, 256):
        x = img_dev[:n]
        print("device n=%d" % n, round(run(lambda: model(x)), 3), "ms")
main()
