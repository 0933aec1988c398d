"""GPU (needs >= 2 B200s; skipped otherwise): the sharded contrastive head -- one process per GPU, embeddings exchanged by
the fused normalise + NVLink peer-store + logits kernel (csrc/comm.cu) -- against the oracle's full logits, plus the
torch.distributed all_gather baseline and repeated calls (epoch parity buffers)."""

import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kind, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    import jimm_oracle as O
    from jimm_b200 import dist as jd
    from jimm_b200.models import CLIP, SigLIP

    jd.init_from_env("nccl")
    tw = 128 if kind == "clip" else 256
    cfg = O.DualCfg(64, 2, 256, 16, 20, 300, tw, tw // 64, 2)
    p = O.random_dual_params(cfg, kind, seed=11)
    Bg = 8 * world
    img, txt = O.synthetic_images(Bg, 64), O.synthetic_tokens(Bg, 20, 300, kind)
    with torch.no_grad():
        ref = (O.clip_forward if kind == "clip" else O.siglip_forward)(p, cfg, img, txt)
    cls = CLIP if kind == "clip" else SigLIP
    m = cls(64, 2, 256, 16, 20, 300, tw, tw // 64, 2, dtype=torch.float16)
    for k, v in p.items():
        m.set_flat_param(k, v)
    lo, hi = jd.shard_range(Bg, rank, world)
    # single-GPU result of the whole batch on this rank: the sharded path must reproduce its row block bit for bit
    m.set_comm("off")
    full = m(img.cuda(), txt.cuda()).cpu()
    errs = []
    for mode in ("peer", "peer", "peer", "nccl"):  # repeated peer calls exercise both parity buffers
        m.set_comm(mode)
        out = m(img[lo:hi].cuda(), txt[lo:hi].cuda())
        assert out.shape == (hi - lo, Bg)
        if mode == "peer":
            assert torch.equal(out.cpu(), full[lo:hi]), "fused peer-memory head differs from the single-GPU head"
        else:
            assert float((out.cpu() - full[lo:hi]).abs().max()) < 1e-4
        errs.append(float((out.cpu().double() - ref[lo:hi].double()).abs().max() / ref.abs().max()))
    # host inputs: ids first, images on a side stream under the text tower, pinned host result
    m.set_comm("peer")
    out_h = m(img[lo:hi].pin_memory(), txt[lo:hi].to(torch.int32).pin_memory())
    assert not out_h.is_cuda and torch.equal(out_h, full[lo:hi])
    # raw uint8 frames (model.set_preprocessor): device frames and pinned host frames equal front-end-then-model bit for bit
    from jimm_b200.preprocess import ImagePreprocessor

    proc = ImagePreprocessor.clip(64) if kind == "clip" else ImagePreprocessor.siglip(64)
    m.set_preprocessor(proc)
    frames = torch.randint(0, 256, (Bg, 64, 64, 3), generator=torch.Generator().manual_seed(5), dtype=torch.uint8)[lo:hi].contiguous()
    ref_u8 = m(proc(frames.cuda(), dtype=torch.float16), txt[lo:hi].cuda())
    assert torch.equal(m(frames.cuda(), txt[lo:hi].cuda()), ref_u8)
    out_u8 = m(frames.pin_memory(), txt[lo:hi].to(torch.int32).pin_memory())
    assert not out_u8.is_cuda and torch.equal(out_u8, ref_u8.cpu())
    torch.cuda.synchronize()
    dist.barrier()
    q.put((rank, errs))
    dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["clip", "siglip"])
@pytest.mark.timeout(600)
def test_sharded_contrastive_head(kind):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=500) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    from gpu_util import record_parity

    for rank, errs in res:
        # the sharded head is asserted BIT-IDENTICAL to the single-GPU head inside the workers; against the fp32 oracle the logits carry the
        # fp16 towers' error amplified by exp(logit_scale) (tests/test_parity_gpu.py LOGITS_TOL), recorded here
        record_parity(f"multi-GPU {kind} head, world {world}, rank {rank}", "logits row block", "float16", "fp32", 2e-3, max(errs))
        assert all(e < 2e-3 for e in errs), (rank, errs)


def _worker_faults(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      JIMM_COMM_TIMEOUT_MS="300")
    import torch.distributed as dist

    from jimm_b200 import _lib
    from jimm_b200 import dist as jd
    from jimm_b200.models import CLIP

    jd.init_from_env("nccl")
    m = CLIP(64, 1, 64, 16, 8, 64, 64, 1, 1, dtype=torch.float16)
    n = m.native(8, require=True)
    n.comm_setup(8)
    res = {}
    ie, te = torch.randn(8, 64, device="cuda"), torch.randn(8, 64, device="cuda")
    ok = n.comm_logits(ie, te)
    torch.cuda.synchronize()
    res["first_ok"] = bool(torch.isfinite(ok).all()) and n.lib.jimm_comm_status(n.handle) == 0
    dist.barrier()
    # 1) ranks disagree on B_local: every rank must see NaN logits and a sticky error naming the mismatch (never stale columns)
    b = 8 if rank == 0 else 6
    out = n.comm_logits(ie[:b], te[:b])
    torch.cuda.synchronize()
    res["mismatch_nan"] = bool(torch.isnan(out).all())
    res["mismatch_rc"] = n.lib.jimm_comm_status(n.handle)
    res["mismatch_msg"] = _lib.last_error()
    try:
        n.comm_logits(ie, te)
        res["next_call_raises"] = False
    except _lib.JimmError:
        res["next_call_raises"] = True
    dist.barrier()
    q.put((rank, res))
    dist.destroy_process_group()


def _worker_timeout(rank, world, port, q):
    import sys
    import time

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                      JIMM_COMM_TIMEOUT_MS="300")
    import torch.distributed as dist

    from jimm_b200 import _lib
    from jimm_b200 import dist as jd
    from jimm_b200.models import CLIP

    jd.init_from_env("nccl")
    m = CLIP(64, 1, 64, 16, 8, 64, 64, 1, 1, dtype=torch.float16)
    n = m.native(8, require=True)
    n.comm_setup(8)
    ie, te = torch.randn(8, 64, device="cuda"), torch.randn(8, 64, device="cuda")
    res = {}
    if rank == 0:  # rank 1 never joins this exchange: the kernel must give up after the timeout instead of spinning forever
        t0 = time.time()
        out = n.comm_logits(ie, te)
        torch.cuda.synchronize()
        res["seconds"] = time.time() - t0
        res["nan"] = bool(torch.isnan(out).all())
        res["rc"] = n.lib.jimm_comm_status(n.handle)
        res["msg"] = _lib.last_error()
    dist.barrier()
    q.put((rank, res))
    dist.destroy_process_group()


def _run(worker, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    return res


@pytest.mark.timeout(400)
def test_comm_rejects_mismatched_rows():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    res = _run(_worker_faults)
    for rank, r in res.items():
        assert r["first_ok"], (rank, r)
        assert r["mismatch_nan"] and r["mismatch_rc"] != 0 and "different number of rows" in r["mismatch_msg"], (rank, r)
        assert r["next_call_raises"], (rank, r)


@pytest.mark.timeout(400)
def test_comm_peer_timeout_is_bounded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    r = _run(_worker_timeout)[0]
    assert r["nan"] and r["rc"] != 0 and "did not publish" in r["msg"], r
    assert r["seconds"] < 30, r
