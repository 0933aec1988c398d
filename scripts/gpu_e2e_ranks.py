"""Per-step e2e times of the uint8 / fp32 host paths on every rank (diagnostic for bench --gpus N: is one rank or one path slow?)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
from jimm_b200 import _lib, build
from jimm_b200 import dist as jd

rank, world, local = jd.init_from_env("nccl")
torch.cuda.set_device(local)
build.build()
lib = _lib.load()
bw = bench.Bench("vit_b16", 0, rank, world, local, lib)
for _ in range(3):
    bw.step_dev()
for name, fn in (("u8", bw.step_host_u8), ("f32", bw.step_host), ("u8", bw.step_host_u8)):
    for _ in range(2):
        fn()
    bw.barrier()
    ts = []
    for _ in range(8):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    bw.barrier()
    print(f"rank {rank} {name}: " + " ".join(f"{t:.2f}" for t in ts), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
