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


# ---------------------------------------------------------------------------------------------------
# data-parallel training plumbing: flat parameter / gradient buffers and their all-reduce
# ---------------------------------------------------------------------------------------------------
def _train_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mslesions3d_b200.ssd3d import LSSD3D
        from mslesions3d_b200.training import FlatParams, cosine_lr
        torch.manual_seed(0)
        model = LSSD3D(n_classes=2, input_channels=1, input_size=(64, 64, 64))
        before = {k: v.clone() for k, v in model.named_parameters()}
        flat = FlatParams(model)
        # parameters are now views of the flat buffer, values unchanged, weights first and biases last
        ok = all(torch.equal(before[k], p) for k, p in model.named_parameters())
        names = list(flat.offsets)
        first_bias = min(i for i, n in enumerate(names) if n.endswith(".bias"))
        ok = ok and all(n.endswith(".bias") for n in names[first_bias:]) and flat.offsets[names[first_bias]] == flat.bias_start
        ok = ok and "rescale_factors" not in flat.offsets and all(o % 8 == 0 for o in flat.offsets.values())
        p = dict(model.named_parameters())["base.features.3.conv2.weight"]
        flat.param[flat.offsets["base.features.3.conv2.weight"]] = 123.0
        ok = ok and float(p.flatten()[0]) == 123.0
        # every rank writes its own gradient; one all-reduce of the flat buffer sums them (fit_step's call)
        for n, g in flat.views_grad.items():
            g.fill_(float(rank + 1))
        dist.all_reduce(flat.grad)
        want = float(sum(r + 1 for r in range(world)))
        ok = ok and all(bool((g == want).all()) for g in flat.views_grad.values())
        ret[rank] = (ok, flat.numel, flat.bias_start, cosine_lr(1.0, 0), cosine_lr(1.0, 40), cosine_lr(1.0, 20))
    finally:
        dist.destroy_process_group()


def test_gloo_flat_gradient_allreduce():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]
    assert ret[0][1:3] == ret[1][1:3]                    # identical layout on every rank
    assert ret[0][1] >= 949936 - 128 and ret[0][1] < 949936 + 8 * 103   # all trainable parameters + padding
    assert ret[0][3] == 1.0 and abs(ret[0][4]) < 1e-12 and abs(ret[0][5] - 0.5) < 1e-12


def _replica_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mslesions3d_b200.ssd3d import LSSD3D
        from mslesions3d_b200.training import (FlatParams, broadcast_replica_state, replica_checksum_matches,
                                               sync_batchnorm_buffers)
        torch.manual_seed(100 + rank)                       # every rank builds a DIFFERENT model
        model = LSSD3D(n_classes=2, input_channels=1, input_size=(64, 64, 64))
        with torch.no_grad():
            for _, b in model.named_buffers():
                if b.is_floating_point():
                    b.add_(float(rank))                      # and different BatchNorm statistics
            model.rescale_factors.fill_(20.0 + rank)
        flat = FlatParams(model)
        flat.exp_avg.fill_(float(rank))
        flat.status[2] = 7 * rank
        differ = not replica_checksum_matches(flat)
        broadcast_replica_state(model, flat, src=0)
        same = replica_checksum_matches(flat)
        digest = float(flat.param.double().sum())
        bn = float(model.base.features[3].bn2.running_mean.double().sum())
        ok = bool((flat.exp_avg == 0).all()) and int(flat.status[2]) == 0 and float(model.rescale_factors.mean()) == 20.0
        # parameters still alias the flat buffer after the broadcast
        p = dict(model.named_parameters())["base.features.0.0.weight"]
        ok = ok and p.data_ptr() == flat.param.data_ptr() + 4 * flat.offsets["base.features.0.0.weight"]
        # rank-local statistics drift apart during training; the on-demand average brings them together again
        with torch.no_grad():
            model.base.features[3].bn2.running_mean.add_(float(rank))
        sync_batchnorm_buffers(model)
        bn2 = float(model.base.features[3].bn2.running_mean.double().sum())
        ret[rank] = (differ, same, digest, bn, ok, bn2)
    finally:
        dist.destroy_process_group()


def test_gloo_replicas_start_identical_from_different_seeds():
    """multi-rank fit_step only all-reduced gradients; ranks built from different RNG states must be
    made replicas of rank 0 (parameters, Adam state, BatchNorm buffers) before the first step."""
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_replica_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert ret[0][0] and ret[1][0]                 # they really started different
    assert ret[0][1] and ret[1][1]                 # and are bit-identical afterwards
    assert ret[0][2] == ret[1][2] and ret[0][3] == ret[1][3]
    assert ret[0][4] and ret[1][4]
    assert ret[0][5] == ret[1][5] and abs(ret[0][5] - (ret[0][3] + 0.5 * 128)) < 1e-3
