"""N > 1 host logic on CPU with the gloo backend (world_size 2): shard ranges tile the volume list, the timing
reduction is a max over ranks, detections gather back in global order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mslesions3d_b200 import parallel


def test_shard_ranges_tile_the_range():
    for n in (0, 1, 7, 8, 64, 1001):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert parallel.rank_world() == (rank, rank, world)
        slow = parallel.max_over_ranks(10.0 + rank)
        a, b = parallel.shard_range(5, rank, world)
        local = [("vol%d" % i, torch.full((i + 1, 6), float(i))) for i in range(a, b)]
        merged = parallel.gather_detections(local, dst=0)
        ret[rank] = (slow, [m[0] for m in merged], [tuple(m[1].shape) for m in merged])
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0][0] == 11.0 and ret[1][0] == 11.0            # max over ranks on every rank
    assert ret[0][1] == ["vol0", "vol1", "vol2", "vol3", "vol4"]   # global order on rank 0
    assert ret[0][2] == [(1, 6), (2, 6), (3, 6), (4, 6), (5, 6)]
    assert ret[1][1] == []
