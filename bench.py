#!/usr/bin/env python
"""bench.py -- images/sec of the ViT forward hot path on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload vit_b16|vit_l16_map|clip_b32|siglip_b16|siglip2_l16_512]

A "step" is one forward pass of the hot path over one synthetic batch.  The default workload is BASELINE.json configs[1]:
ViT-B/16 @224, batch 256 per GPU, fp16 tensor-core operands (fp32 accumulate / residual / LN / softmax), random-init weights.
  value  : whole-job images/sec with the inputs already resident in HBM (CUDA events, barrier + synchronize both sides,
           max over ranks).  Weak scaling: every rank runs its own 256-image batch; no data-path collective for ViT.
  e2e    : the same metric through the public Python API with HOST (pinned) inputs: H2D copy + forward + D2H of the logits
           inside the timed region (jimm_vit_forward_host), with a synchronisation every step.  `pipelined_depth2_value` is the
           same loop through `forward_async` with two calls in flight (every step still copies its inputs in and its result out);
           it is an extra, not the headline.
  roofline: the dominant kernel (tcgen05 GEMM) timed live with CUDA events around every launch of the timed steps.
  cpu_baseline: the CPU oracle (torch fp32, jimm semantics -- the stand-in for the reference's JAX-CPU path, which cannot
           be installed here) on a bounded sample, rank 0 / N=1 only.
--impl reference times that CPU path alone, on the same metric / config (see the tier's reference-arm contract).
"""

from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_IMG = {"vit_b16": 35.128, "vit_l16_map": 383.85, "clip_b32": 14.778, "siglip_b16": 57.85, "siglip2_l16_512": 766.55}

WORKLOADS = {
    # name: (description, per-GPU batch, dtype)
    "vit_b16": ("ViT-B/16 @224, batch 256/GPU, fp16 operands, random-init (BASELINE configs[1])", 256, "float16"),
    "vit_l16_map": ("ViT-L/16 @384 MAP head, batch 128/GPU, bf16 operands (BASELINE configs[2])", 128, "bfloat16"),
    "clip_b32": ("CLIP ViT-B/32 dual tower, batch 256 pairs/GPU, fp16 (BASELINE configs[3])", 256, "float16"),
    "siglip_b16": ("SigLIP-B/16 @256 dual tower, batch 256 pairs/GPU, fp16 (north-star extra)", 256, "float16"),
    "siglip2_l16_512": ("SigLIP2-L/16 @512 dual tower (S=1024), batch 256 pairs/GPU, bf16, vocab 32000 stand-in (BASELINE configs[4] per-GPU share)",
                        256, "bfloat16"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1402.5), d.get("hbm_gbs", 6568.4), "measured (MEASURED_PEAKS.json, sustained bf16 GEMM)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md 'clocks line')."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []   # (arrival time, csv line)
        self.proc = None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def in_window(self) -> int:
        """Samples that arrived while the timed region ran (a sample describes the ~100 ms before it arrives)."""
        if self.t0 is None or self.t1 is None:
            return len(self.rows)
        return sum(1 for t, _ in self.rows if self.t0 <= t <= self.t1 + 0.12)

    def selected_rows(self):
        """Samples of the timed region; if the region was too short for one, those of the post-region load the caller kept running
        (single GPU) or, failing that, the last warm-up samples (multi-GPU runs cannot extend the load unilaterally)."""
        if self.t0 is None or self.t1 is None:
            return list(self.rows)
        lo = self.t0 if self.in_window() else self.t1
        return [(t, r) for t, r in self.rows if t >= lo] or self.rows[-3:]

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for _, r in self.selected_rows():
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def build_model(workload: str, dtype_name: str):
    """Random-init weights of the named architecture (no network for checkpoints): the reference's init distributions."""
    import torch

    from jimm_b200 import Rngs
    from jimm_b200.common.vit import VisionTransformerBase
    from jimm_b200.models import CLIP, SigLIP, VisionTransformer

    dt = getattr(torch, dtype_name)
    if workload == "vit_b16":
        return VisionTransformer(dtype=dt, rngs=Rngs(0)).eval(), 224, None
    if workload == "vit_l16_map":
        return VisionTransformerBase(img_size=384, patch_size=16, in_channels=3, hidden_size=1024, num_layers=24, num_heads=16, mlp_dim=4096,
                                     pooling_type="MAP", layernorm_epsilon=1e-6, dtype=dt, rngs=Rngs(0)), 384, None
    if workload == "clip_b32":
        return CLIP(224, 12, 768, 32, 77, 49408, 512, 8, 12, dtype=dt, rngs=Rngs(0)), 224, (77, 49408, "clip")
    if workload == "siglip_b16":
        return SigLIP(256, 12, 768, 16, 64, 32000, 768, 12, 12, dtype=dt, rngs=Rngs(0)), 256, (64, 32000, "siglip")
    if workload == "siglip2_l16_512":
        # vision 1024/24L/16H patch 16 @512 (S = 1024, MAP head), text 1024/24L/16H/T64; the embedding gather
        # does not depend on the vocabulary size, 32000 keeps the random init short
        return SigLIP(512, 24, 1024, 16, 64, 32000, 1024, 16, 24, dtype=dt, rngs=Rngs(0)), 512, (64, 32000, "siglip")
    raise SystemExit(f"unknown workload {workload}")


def cpu_threads() -> int:
    """Threads for the CPU arm: every core this process may use (affinity mask and cgroup CPU quota honoured), capped at 32 --
    torch's intra-op pool stops scaling, then regresses, beyond that for this model size; JIMM_CPU_THREADS overrides."""
    env = os.environ.get("JIMM_CPU_THREADS")
    if env:
        return max(1, int(env))
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 32))


def oracle_step_fn(workload: str, B: int):
    """The CPU arm: oracle restatement in torch fp32 with jimm semantics, same architecture / synthetic inputs."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import jimm_oracle as O

    torch.set_num_threads(cpu_threads())
    if workload == "vit_b16":
        cfg = O.ViTCfg()
        p = O.random_vit_params(cfg, seed=0)
        img = O.synthetic_images(B, 224)
        return lambda: O.vit_forward(p, cfg, img)
    if workload == "vit_l16_map":
        t = O.TowerCfg(384, 16, 3, 1024, 24, 16, 4096, "MAP", layernorm_epsilon=1e-6)
        p = O.random_tower_params(t, seed=0)
        img = O.synthetic_images(B, 384)
        return lambda: O.vision_tower(p, "", img, t)
    if workload == "clip_b32":
        cfg = O.DualCfg(224, 12, 768, 32, 77, 49408, 512, 8, 12)
        p = O.random_dual_params(cfg, "clip", seed=0)
        img, txt = O.synthetic_images(B, 224), O.synthetic_tokens(B, 77, 49408, "clip")
        return lambda: O.clip_forward(p, cfg, img, txt)
    if workload == "siglip2_l16_512":
        cfg = O.DualCfg(512, 24, 1024, 16, 64, 32000, 1024, 16, 24)
        p = O.random_dual_params(cfg, "siglip", seed=0)
        img, txt = O.synthetic_images(B, 512), O.synthetic_tokens(B, 64, 32000, "siglip")
        return lambda: O.siglip_forward(p, cfg, img, txt)
    cfg = O.DualCfg(256, 12, 768, 16, 64, 32000, 768, 12, 12)
    p = O.random_dual_params(cfg, "siglip", seed=0)
    img, txt = O.synthetic_images(B, 256), O.synthetic_tokens(B, 64, 32000, "siglip")
    return lambda: O.siglip_forward(p, cfg, img, txt)


def time_cpu(workload: str, B: int, steps: int, warmup: int):
    import torch

    fn = oracle_step_fn(workload, B)
    with torch.no_grad():
        for _ in range(warmup):
            fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps * 1e3


# ----------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the CPU restatement of the reference's forward (oracle/jimm_oracle.py, torch fp32, jimm semantics;
    the reference's own JAX-CPU path cannot be installed in this image) on the host cores, same metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # bounded sample: size the per-step batch from one probe image so a step stays around 3 s of host time
    probe_ips, _ = time_cpu(args.workload, 1, 1, 1)
    B = args.cpu_batch if args.cpu_batch_fixed else max(1, min(16, int(probe_ips * 3.0)))
    ips, ms = time_cpu(args.workload, B, args.steps, args.warmup)
    cores = cpu_threads()
    sample = f"{B} images/step x {args.steps} steps of {WORKLOADS[args.workload][0]}"
    line = {
        "impl": "reference", "metric": "images/sec", "value": ips, "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOADS[args.workload][0], "cpu_sample_batch": B},
        "cpu_baseline": {"value": ips, "unit": "images/sec", "cores": cores, "host_cores": os.cpu_count() or 1, "kind": "port", "sample": sample,
                         "note": "oracle/jimm_oracle.py (torch CPU fp32, jimm semantics); the reference's JAX-CPU path is not installable here"},
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch

    from jimm_b200 import _lib, build
    from jimm_b200 import dist as jd

    rank, world, local = jd.init_from_env("nccl")
    if world != args.gpus and world > 1:
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    build.build()
    lib = _lib.load()
    desc, B, dtype_name = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    model, img_size, text = build_model(args.workload, dtype_name)
    model.set_max_batch(B)
    dual = text is not None

    g = torch.Generator().manual_seed(1234 + rank)
    img_host = torch.randn(B, img_size, img_size, 3, generator=g, dtype=torch.float32).pin_memory()
    img_dev = img_host.to(dev)
    if dual:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import jimm_oracle as O

        ids_host = O.synthetic_tokens(B, text[0], text[1], text[2], seed=4321 + rank).to(torch.int32).pin_memory()
        ids_dev = ids_host.to(dev)
        step_dev = lambda: model(img_dev, ids_dev)
        step_host = lambda: model(img_host, ids_host)
    else:
        step_dev = lambda: model(img_dev)
        step_host = lambda: model(img_host)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        return jd.max_over_ranks(e0.elapsed_time(e1)), out

    # ---- warm-up (also builds the native handle); the clock sampler is started first so that nvidia-smi is already streaming
    #      when the timed region begins (its start-up can take longer than a short timed region) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_dev()
    torch.cuda.synchronize(dev)
    native = model.native(B)

    # ---- value: device-resident inputs ----
    l0 = lib.jimm_launch_count()
    sampler.mark_begin()
    ms_total, out = timed(step_dev, args.steps)
    sampler.mark_end()
    launches = lib.jimm_launch_count() - l0
    clocks = None
    if rank == 0:
        if sampler.proc is not None and sampler.in_window() == 0 and world == 1:
            # the region was shorter than the sampling period: keep the same load running (untimed) until a sample lands
            t_wait = time.time()
            while len([1 for t, _ in sampler.rows if t >= sampler.t1]) < 2 and time.time() - t_wait < 3.0:
                step_dev()
                torch.cuda.synchronize(dev)
        clocks = sampler.stop()
    ms_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- roofline: tcgen05 GEMM launches of the same steps, bracketed by events on the launch stream ----
    _lib.check(lib.jimm_profile_begin(native.handle))
    for _ in range(args.steps):
        step_dev()
    g_ms, g_fl, g_n = C.c_double(), C.c_double(), C.c_longlong()
    _lib.check(lib.jimm_profile_end(native.handle, C.byref(g_ms), C.byref(g_fl), C.byref(g_n)))
    peak_tf, peak_bw, peak_src = peaks()
    achieved = g_fl.value / (g_ms.value * 1e-3) / 1e12 if g_ms.value > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "traffic": traffic, "peak_source": peak_src, "launches": g_n.value,
                "avg_launch_ms": g_ms.value / max(g_n.value, 1), "gemm_share_of_step": g_ms.value / args.steps / ms_step}

    # ---- e2e: host buffers through the public API (H2D + forward + D2H inside the timed region) ----
    for _ in range(2):
        step_host()
    ms_e2e, out_h = timed(step_host, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    # the same loop with asynchronous dispatch, two calls in flight (every step still copies its inputs in and its result out)
    pipelined = None
    if not dual and hasattr(model, "forward_async"):
        state = {"pending": None}

        def step_async():
            nxt = model.forward_async(img_host)
            out = state["pending"].result() if state["pending"] is not None else None
            state["pending"] = nxt
            return out

        for _ in range(2):
            step_async()
        ms_pipe, _ = timed(step_async, args.steps)
        state["pending"].result()
        pipelined = world * B * args.steps / (ms_pipe * 1e-3)
    h2d = img_host.numel() * 4 + (ids_host.numel() * 4 if dual else 0)
    d2h = out_h.numel() * 4

    # ---- CPU baseline (rank 0, N=1 only; bounded sample) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # bounded sample of the same workload: ~12 s of host time, sized from a one-image probe
        probe_ips, _ = time_cpu(args.workload, 1, 1, 1)
        cb = max(1, min(16, int(probe_ips * 2.0)))
        csteps = args.cpu_steps if args.cpu_steps > 0 else max(2, min(60, int(12.0 * probe_ips / cb)))
        ips, _ = time_cpu(args.workload, cb, csteps, 1)
        cpu = {"value": ips, "unit": "images/sec", "cores": cpu_threads(), "host_cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{cb} images/step x {csteps} steps, oracle/jimm_oracle.py torch-CPU fp32 (jimm semantics)"}

    if rank == 0:
        gflop = GFLOP_PER_IMG[args.workload]
        line = {
            "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"float16": "f16", "bfloat16": "bf16", "float32": "tf32"}[dtype_name], "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": B, "global_batch": world * B, "parallelism": f"dp{world}",
                       "l2_policy": "inputs_larger_than_L2 (154 MB fp32 images; >1 GB of activations streamed per step)",
                       "gflop_per_image": gflop},
            "model_tflops": value * gflop / 1e3, "model_frac_of_peak": value * gflop / 1e3 / (peak_tf * world),
            "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                    "sync": "every step (value above)", "pipelined_depth2_value": pipelined},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vit_b16", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--cpu-steps", type=int, default=0, help="0 = size the CPU sample for ~12 s")
    ap.add_argument("--cpu-batch-fixed", action="store_true", help="reference arm: use --cpu-batch instead of sizing it from a probe")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
