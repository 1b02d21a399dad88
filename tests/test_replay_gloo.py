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
