#!/usr/bin/env python
"""Training-step throughput (BASELINE.json configs[2]): SSD3D fit_step = forward + IoU matching + MultiBox loss
+ backward + (NCCL gradient all-reduce) + Adam, 1ch 96^3, batch 16 per GPU, data parallel.

    python scripts/bench_train.py [--steps K] [--warmup W] [--batch B] [--size S] [--cpu-baseline]
    python -m torch.distributed.run --nproc-per-node N ... scripts/bench_train.py

Prints one JSON line (rank 0): volumes/s over all ranks, ms/step (CUDA events, max over ranks), launches/step.
``--cpu-baseline`` also times the oracle's torch-CPU training step (the reference's mechanism) on a bounded
sample.  ``--profile`` prints a per-phase breakdown (forward / loss / backward / adam) measured with events.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=96)
    ap.add_argument("--channels", type=int, default=1)
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph (for ncu launch lists)")
    args = ap.parse_args()

    import torch.distributed as dist
    from mslesions3d_b200 import _lib, ops, synthetic, training
    from mslesions3d_b200.parallel import rank_world, max_over_ranks
    from mslesions3d_b200.ssd3d import LSSD3D
    from oracle import ssd3d_oracle as O

    rank, local_rank, world = rank_world()
    _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    size = (args.size,) * 3
    sd = synthetic.random_state_dict(args.channels, seed=0)
    model = LSSD3D(n_classes=2, input_channels=args.channels, input_size=size, threshold=[0.1, 0.2], lr=1e-4)
    model.load_state_dict(sd)
    model = model.to(dev).train()
    model.use_cuda_graph = not args.eager
    n_rot = 3
    batches = []
    for r in range(n_rot):
        x, b, l = synthetic.make_batch(args.batch, args.channels, size, first_idx=(rank * n_rot + r) * args.batch,
                                       with_boxes=True)
        batches.append({"img": torch.from_numpy(x).to(dev), "boxes": [torch.from_numpy(v).to(dev) for v in b],
                        "labels": [torch.from_numpy(v).to(dev) for v in l]})

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        model.fit_step(batches[i % n_rot])
    barrier()
    ops.LAUNCHES[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    last = None
    for i in range(args.steps):
        last = model.fit_step(batches[i % n_rot])
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    launches = ops.LAUNCHES[0]
    line = {"metric": "SSD3D training volumes/sec (fwd + matching + loss + bwd + Adam)",
            "value": world * args.batch * args.steps / (ms / 1000.0), "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "host_ms_per_step": 1000.0 * wall / args.steps, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "SSD3D training step, %dch %d^3, batch %d per GPU, threshold [0.1,0.2], Adam + "
                       "cosine schedule, flat-gradient NCCL all-reduce" % (args.channels, args.size, args.batch),
                       "global_batch": args.batch * world, "parallelism": "dp%d" % world},
            "gpu_launches_per_step": launches / args.steps, "loss": [float(v) for v in last.cpu()]}
    if args.profile and rank == 0:
        # per-phase device time of one step (events around the phases of training.fit_step, re-run eagerly)
        eng = model.train_engine()
        flat = eng.flatten()
        b = batches[0]
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.no_grad():
            ev[0].record()
            locs, scores = eng.forward(b["img"])
            ev[1].record()
            m = model.loss_fn.match(b["boxes"], b["labels"])
            out, n_pos, gl, gs = ops.multibox_loss(locs, scores, m["true_classes"], m["true_locs"])
            ev[2].record()
            grads = training._Grads(((n, p) for n, p in model.named_parameters() if n != "rescale_factors"), flat)
            eng.backward(gl, gs, grads)
            ev[3].record()
            ops.adam_step(flat.param, flat.grad, flat.exp_avg, flat.exp_avg_sq, flat.bias_start, 1e-4, 2e-4, flat.step + 1)
            ev[4].record()
        torch.cuda.synchronize()
        line["phases_ms"] = {k: ev[i].elapsed_time(ev[i + 1]) for i, k in
                             enumerate(["forward", "match+loss", "backward", "adam"])}
    if args.cpu_baseline and rank == 0 and world == 1:
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        nb = 2
        x, b, l = synthetic.make_batch(nb, args.channels, size, with_boxes=True)
        pri = O.prior_boxes_fast(size, in_channels=args.channels)
        bx, lb = [torch.from_numpy(v) for v in b], [torch.from_numpy(v) for v in l]
        t0 = time.perf_counter()
        O.fit_steps(sd, [(torch.from_numpy(x), bx, lb)] * 2, pri, [0.1, 0.2], 1e-4)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 2 * nb / dt, "unit": "volumes/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "2 steps of batch %d (same shape), torch-CPU autograd + torch.optim.Adam" % nb}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
