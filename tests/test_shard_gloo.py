"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous frame-pair partitioning + pose all-gather."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rgbd_odometry_b200.shard import gather_poses, partition


def test_partition_covers_everything_once():
    for total in (0, 1, 7, 16, 1024, 16384, 1001):
        for world in (1, 2, 3, 4, 8):
            blocks = [partition(total, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = partition(total, world, rank)
    local = np.arange(start, start + count, dtype=np.float64)[:, None] * np.ones((1, 12)) + np.arange(12) * 1e-3
    allp = gather_poses(local, total)
    if rank == 0:
        q.put(allp.numpy())
    dist.destroy_process_group()


def test_pose_gather_gloo_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    total = 11                      # ragged: 6 + 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(total, dtype=np.float64)[:, None] * np.ones((1, 12)) + np.arange(12) * 1e-3
    assert np.array_equal(got, want)
