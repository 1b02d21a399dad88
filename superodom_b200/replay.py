"""Batched scan replay against a frozen map, sharded over GPUs (BASELINE cfg4).

Each scan's registration depends only on (scan, prior, map): the units are independent, so ranks take contiguous
chunks of the scan list with NO data-path collective.  The only exchange is ONE gather of the per-scan result rows at
the end of a replay: every rank's context appends {pose_opt[7], status + 256 n_iterations} rows to a device buffer
(`so_set_pose_sink`), and `gather_rows` enqueues the all-gather on the stream the registrations ran on -- no host hop
between the last optimiser step and the collective (NCCL over NVLink on GPUs; gloo on CPU tensors in the tests).
torch.distributed is plumbing only.
"""
from __future__ import annotations

import numpy as np

ROW = 8      # doubles per scan in a pose-sink row


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) of scan indices owned by `rank` (first n_total % world ranks get one more)."""
    base, rem = divmod(n_total, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_cap(n_total: int, world: int) -> int:
    """Rows every rank's sink holds: the largest shard (ragged shards are padded to it for the all-gather)."""
    return (n_total + world - 1) // world


def gather_rows(local_rows, out=None):
    """All-gather equally sized per-rank row blocks [cap, ROW] (torch tensors, device or CPU) into [world * cap, ROW].
    Asynchronous with respect to the host on a CUDA tensor: the collective is enqueued on the current stream."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        if out is None:
            return local_rows
        out.copy_(local_rows)
        return out
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local_rows.shape[0], local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, local_rows)
    return out


def unpack_rows(all_rows, n_total: int, world: int):
    """[world * cap, ROW] gathered rows (numpy) -> (poses [n_total, 7], status [n_total], n_iterations [n_total]) in scan order."""
    a = np.asarray(all_rows, dtype=np.float64).reshape(world, -1, ROW)
    parts = []
    for r in range(world):
        b, e = shard_range(n_total, r, world)
        parts.append(a[r, : e - b])
    rows = np.concatenate(parts, 0) if parts else np.zeros((0, ROW))
    code = rows[:, 7].astype(np.int64)
    return rows[:, :7].copy(), code % 256, code // 256


def gather_poses(local_poses, n_total: int, rank: int, world: int, device=None):
    """Host-array convenience form: all-gather per-rank [n_local, 7] float64 pose blocks into the global [n_total, 7] array."""
    import torch

    local = np.ascontiguousarray(local_poses, dtype=np.float64).reshape(-1, 7)
    cap = shard_cap(n_total, world)
    buf = torch.zeros((cap, ROW), dtype=torch.float64)
    buf[: local.shape[0], :7] = torch.from_numpy(local)
    if device is not None:
        buf = buf.to(device)
    out = gather_rows(buf)
    return unpack_rows(out.cpu().numpy(), n_total, world)[0]


def replay(register_fn, n_total: int, rank: int, world: int, batch: int, device=None):
    """Run `register_fn(begin, end) -> [end-begin, 7] poses` over this rank's shard in batches, then gather ONCE."""
    b, e = shard_range(n_total, rank, world)
    poses = []
    for s in range(b, e, batch):
        poses.append(np.asarray(register_fn(s, min(s + batch, e)), dtype=np.float64).reshape(-1, 7))
    local = np.concatenate(poses, 0) if poses else np.zeros((0, 7))
    return gather_poses(local, n_total, rank, world, device)
