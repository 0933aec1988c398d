"""e2e diagnostics for ViT-B/16 B=256: device-resident step interleaved with the host paths (uint8 sync / fp32 sync / uint8 async depth 2),
20 steps each, in one process -- separates clock drift under the power cap from path effects."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from jimm_b200 import _lib, build

build.build()
lib = _lib.load()
torch.cuda.set_device(0)
bw = bench.Bench("vit_b16", 0, 0, 1, 0, lib)
state = {"pending": None}


def step_async():
    nxt = bw.model.forward_async(bw.u8_host)
    res = state["pending"].result() if state["pending"] is not None else None
    state["pending"] = nxt
    return res


def run(name, fn, n=20):
    for _ in range(3):
        fn()
    if state["pending"] is not None and fn is not step_async:
        state["pending"].result()
        state["pending"] = None
    ms, _ = bw.timed(fn, n)
    if fn is step_async:
        state["pending"].result()
        state["pending"] = None
    print(f"{name:10s} {ms/n:.3f} ms", flush=True)


order = os.environ.get("ORDER", "dev,u8,dev,u8,f32,dev,u8,u8,async,dev,f32,u8,dev").split(",")
fns = {"dev": bw.step_dev, "u8": bw.step_host_u8, "f32": bw.step_host, "async": step_async}
for o in order:
    run(o, fns[o])
