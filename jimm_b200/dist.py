"""torch.distributed plumbing for the one-process-per-GPU layout (control plane only: rendezvous, batch sharding,
exchange of the 64-byte CUDA-IPC handles of the NVLink gather buffers)."""

from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group from torchrun's env (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*).
    Returns (rank, world, local_rank).  No-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            # binding the process group to its device avoids NCCL's rank -> GPU guess (and the hang it warns about)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [lo, hi) of n samples for `rank` (the reference shards the batch axis: P("batch", ...))."""
    if n % world != 0:
        raise ValueError(f"global batch {n} is not divisible by the number of ranks {world}")
    per = n // world
    return rank * per, (rank + 1) * per


def exchange_handles(handle: bytes, group=None) -> bytes:
    """All-gather the per-rank 64-byte IPC handles, concatenated in rank order."""
    if len(handle) != 64:
        raise ValueError("IPC handle must be 64 bytes")
    world = dist.get_world_size(group)
    out: List[bytes | None] = [None] * world
    dist.all_gather_object(out, handle, group=group)
    return b"".join(out)  # type: ignore[arg-type]


def max_over_ranks(value: float) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
