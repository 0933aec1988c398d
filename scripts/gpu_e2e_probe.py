"""e2e (host uint8 frames -> logits on host) of ViT-B/16 B=256 under different host-slice layouts (JIMM_HOST_SLICES) vs the device-resident step."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch

    import bench
    from jimm_b200 import _lib, build

    build.build()
    lib = _lib.load()
    torch.cuda.set_device(0)
    bw = bench.Bench("vit_b16", 0, 0, 1, 0, lib)
    for _ in range(5):
        bw.step_dev()
    ms, _ = bw.timed(bw.step_dev, 20)
    for _ in range(3):
        bw.step_host_u8()
    ms8, _ = bw.timed(bw.step_host_u8, 20)
    for _ in range(3):
        bw.step_host()
    ms32, _ = bw.timed(bw.step_host, 20)
    ms_again, _ = bw.timed(bw.step_dev, 20)
    state = {"pending": None}

    def step_async():
        nxt = bw.model.forward_async(bw.u8_host)
        res = state["pending"].result() if state["pending"] is not None else None
        state["pending"] = nxt
        return res

    for _ in range(3):
        step_async()
    msp, _ = bw.timed(step_async, 20)
    state["pending"].result()
    ms_last, _ = bw.timed(bw.step_dev, 20)
    print(f"slices={os.environ.get('JIMM_HOST_SLICES','auto'):8s} device {ms/20:.3f} ms  e2e u8 {ms8/20:.3f} ms  e2e f32 {ms32/20:.3f} ms  device again {ms_again/20:.3f}  "
          f"u8 async depth 2 {msp/20:.3f}  device last {ms_last/20:.3f}", flush=True)
else:
    for sl in (None,):
        env = dict(os.environ)
        if sl:
            env["JIMM_HOST_SLICES"] = sl
        subprocess.run([sys.executable, __file__, "child"], env=env)
