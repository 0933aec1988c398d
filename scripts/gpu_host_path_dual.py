"""Experiment: where the e2e time of the dual-tower host path goes (CLIP-B/32 B=256)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

def main(workload):
    dev = torch.device("cuda:0")
    model, img_size, text = bench.build_model(workload, bench.WORKLOADS[workload][2])
    B = 256
    model.set_max_batch(B)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import jimm_oracle as O
    img_host = torch.randn(B, img_size, img_size, 3).pin_memory()
    ids_host = O.synthetic_tokens(B, text[0], text[1], text[2], seed=1).to(torch.int32).pin_memory()
    img_dev, ids_dev = img_host.to(dev), ids_host.to(dev)
    for _ in range(5):
        model(img_dev, ids_dev)
    torch.cuda.synchronize()
    def run(fn, n=20, sync_each=False):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
            if sync_each: torch.cuda.synchronize()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        return round((t1 - t0) / n * 1e3, 3), round((t2 - t0) / n * 1e3, 3)
    print(workload)
    print("device dual        (enqueue ms, total ms)", run(lambda: model(img_dev, ids_dev)))
    print("device dual sync/step                    ", run(lambda: model(img_dev, ids_dev), sync_each=True))
    print("device encode_image                      ", run(lambda: model.encode_image(img_dev)))
    print("device encode_text                       ", run(lambda: model.encode_text(ids_dev)))
    print("H2D images                               ", run(lambda: img_dev.copy_(img_host, non_blocking=True)))
    for sl in ["256", "64,192", "64,64,128", "32,64,160"]:
        os.environ["JIMM_HOST_SLICES"] = sl
        print("host dual slices", sl, run(lambda: model(img_host, ids_host)), flush=True)
    os.environ.pop("JIMM_HOST_SLICES")
    print("host dual default", run(lambda: model(img_host, ids_host)))
main(sys.argv[1] if len(sys.argv) > 1 else "clip_b32")
