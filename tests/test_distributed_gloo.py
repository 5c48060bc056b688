"""World-size-2 test of the multi-GPU host logic on CPU (gloo): sharding covers the batch exactly once with no
overlap, per-rank seeds differ, and the throughput aggregation is sum(frames) / max(elapsed)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vi_depth_completion_b200 import sharding as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests import common as C
        total = 2 * 256 + 1                     # ragged on purpose
        b, e = S.shard_range(total, rank, world)
        I_g, _ = C.random_gravity(e - b, seed=S.rank_seed(1234, rank))
        elapsed = 10.0 * (rank + 1)             # rank 1 is the slow one
        frames, ms, fps = S.aggregate_throughput(e - b, elapsed)
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, e, float(I_g[0, 0])))
        dist.barrier()
        q.put((rank, frames, ms, fps, gathered))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_timing_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, frames, ms, fps, gathered in res:
        assert frames == 513 and ms == 20.0 and fps == 513 / 0.020
        (b0, e0, g0), (b1, e1, g1) = gathered
        assert b0 == 0 and e0 == b1 and e1 == 513 and (e0 - b0) - (e1 - b1) in (0, 1)
        assert g0 != g1                          # different per-rank seeds -> different frames


def test_shard_range_properties():
    for total in (0, 1, 7, 256, 2048, 2049):
        for world in (1, 2, 4, 8):
            cuts = [S.shard_range(total, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == total
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1
    frames, ms, fps = S.aggregate_throughput(256, 1.0)      # not initialised -> identity
    assert (frames, ms, fps) == (256.0, 1.0, 256000.0)
