"""World-size-2 gloo test of the multi-rank host logic: disjoint parameter-point blocks, gather in point order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosmomc_galileon_b200.sharding import gather_cls, gather_cls_to_root, shard_range


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    # stand-in for the GPU stage: a deterministic "spectrum" per global point index
    local = np.stack([np.full((3, 5), float(i)) + np.arange(5) for i in range(lo, hi)]) if hi > lo else np.zeros((0, 3, 5))
    full = gather_cls(local, n_total)
    at_root = gather_cls_to_root(local, n_total)
    assert (at_root is None) == (rank != 0)
    if rank == 0:
        np.testing.assert_array_equal(at_root, full)
    q.put((rank, full))
    dist.destroy_process_group()


def test_two_rank_gather_keeps_point_order():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_total = 7  # uneven split: 4 + 3
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.stack([np.full((3, 5), float(i)) + np.arange(5) for i in range(n_total)])
    for _, full in results:
        np.testing.assert_array_equal(full, expect)
