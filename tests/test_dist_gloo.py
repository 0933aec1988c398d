"""CPU, world_size 2, gloo: the host-side logic of the N>1 path -- rendezvous from env, batch sharding, IPC-handle
exchange order, max-over-ranks timing reduction, and that the sharded contrastive head (each rank: local image rows x
all-gathered text rows) reassembles to the oracle's full logits."""

import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "oracle"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist

    import jimm_oracle as O
    from jimm_b200 import dist as jd

    r, w, _ = jd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    lo, hi = jd.shard_range(8, r, w)
    blob = jd.exchange_handles(bytes([r]) * 64)
    assert blob == b"".join(bytes([i]) * 64 for i in range(w))
    t = jd.max_over_ranks(1.0 + r)
    # sharded contrastive head vs oracle
    oc = O.DualCfg(32, 1, 64, 8, 8, 50, 64, 1, 1)
    p = O.random_dual_params(oc, "siglip", seed=5, dtype=torch.float64)
    img = O.synthetic_images(8, 32, dtype=torch.float64)
    txt = O.synthetic_tokens(8, 8, 50, "siglip")
    full = O.siglip_forward(p, oc, img, txt)
    ie = O.siglip_encode_image(p, oc, img[lo:hi])
    te = O.siglip_encode_text(p, oc, txt[lo:hi])
    i_n = ie / torch.linalg.norm(ie, dim=-1, keepdim=True)
    t_n = te / torch.linalg.norm(te, dim=-1, keepdim=True)
    gathered = [torch.empty_like(t_n) for _ in range(w)]
    dist.all_gather(gathered, t_n)
    block = torch.exp(p["logit_scale"]) * i_n @ torch.cat(gathered).T + p["logit_bias"]
    err = float((block - full[lo:hi]).abs().max())
    q.put((rank, lo, hi, t, err))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 4), (4, 8)]
    assert all(r[3] == 2.0 for r in res)  # max over ranks
    assert all(r[4] < 1e-12 for r in res)


def test_shard_range_errors():
    from jimm_b200 import dist as jd

    with pytest.raises(ValueError):
        jd.shard_range(7, 0, 2)
    assert jd.shard_range(8, 1, 4) == (2, 4)
    assert jd.max_over_ranks(3.0) == 3.0
