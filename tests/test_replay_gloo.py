"""world_size-2 gloo test of the multi-GPU replay plumbing (sharding + pose gather), on CPU."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from superodom_b200 import replay


def test_shard_range_partitions():
    for n in (0, 1, 7, 1024, 1027):
        for w in (1, 2, 3, 8):
            r = [replay.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def _fake_pose(i):
    return np.array([i, 2 * i, -i, 0, 0, 0, 1], dtype=np.float64)


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        out = replay.replay(lambda b, e: np.stack([_fake_pose(i) for i in range(b, e)]), n_total, rank, world, batch=3)
        # the pose-sink form: every rank holds [cap, 8] rows {pose7, status + 256 n_iter}; one all-gather; unpack in scan order
        import torch
        b, e = replay.shard_range(n_total, rank, world)
        rows = torch.zeros((replay.shard_cap(n_total, world), replay.ROW), dtype=torch.float64)
        for k, i in enumerate(range(b, e)):
            rows[k, :7] = torch.from_numpy(_fake_pose(i))
            rows[k, 7] = float((i % 3) + 256 * (i + 1))
        poses, status, iters = replay.unpack_rows(replay.gather_rows(rows).numpy(), n_total, world)
        assert np.array_equal(poses, out) and np.array_equal(status, np.arange(n_total) % 3) and np.array_equal(iters, np.arange(n_total) + 1)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_replay_gather_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    n_total = 11          # ragged: 6 + 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([_fake_pose(i) for i in range(n_total)])
    for r in range(2):
        assert np.array_equal(got[r], want)
