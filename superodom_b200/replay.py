"""Batched scan replay against a frozen map, sharded over GPUs (BASELINE cfg4).

Each scan's registration depends only on (scan, prior, map): the units are independent, so ranks take contiguous
chunks of the scan list with NO data-path collective.  The only exchange is the gather of the 7-double poses at
the end of a step (NCCL over NVLink on GPUs; gloo in the CPU tests).  torch.distributed is plumbing only.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of scan indices owned by `rank` (first n_total % world ranks get one more)."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_poses(local_poses, n_total: int, rank: int, world: int, device=None):
    """All-gather per-rank [n_local, 7] float64 pose blocks into the global [n_total, 7] array, in scan order.
    Uses torch.distributed (backend chosen by the caller: nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    local = torch.as_tensor(np.ascontiguousarray(local_poses, dtype=np.float64))
    if world == 1 or not dist.is_initialized():
        return local.numpy().copy()
    cap = (n_total + world - 1) // world                 # ragged shards are padded to the largest one
    buf = torch.zeros((cap, 7), dtype=torch.float64)
    buf[: local.shape[0]] = local
    if device is not None:
        buf = buf.to(device)
    out = torch.empty((world * cap, 7), dtype=torch.float64, device=buf.device)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, cap, 7)
    parts = []
    for r in range(world):
        b, e = shard_range(n_total, r, world)
        parts.append(out[r, : e - b])
    return np.concatenate(parts, 0)


def replay(register_fn, n_total: int, rank: int, world: int, batch: int, device=None):
    """Run `register_fn(begin, end) -> [end-begin, 7] poses` over this rank's shard in batches, then gather."""
    b, e = shard_range(n_total, rank, world)
    poses = []
    for s in range(b, e, batch):
        poses.append(np.asarray(register_fn(s, min(s + batch, e)), dtype=np.float64).reshape(-1, 7))
    local = np.concatenate(poses, 0) if poses else np.zeros((0, 7))
    return gather_poses(local, n_total, rank, world, device)
