"""Batch-of-volumes data parallelism: every volume is independent in forward, decode
and NMS, so N GPUs = N processes (one per GPU, ``torchrun``), each taking a contiguous shard of the volumes.
No collective sits on the inference data path; ``torch.distributed`` (NCCL on GPUs, gloo in CPU tests) is used
only for the barrier / max-over-ranks timing reduction and for gathering variable-length detection lists to
rank 0 on the host."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def rank_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) share of ``n_items`` volumes for ``rank``; sizes differ by at most one and
    the shards tile [0, n_items) in rank order."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard request: n=%d rank=%d world=%d" % (n_items, rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (elapsed time): the job is as slow as its slowest rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_detections(local: Sequence, dst: int = 0) -> List:
    """Gather per-rank lists of per-volume detection tuples to ``dst`` in global volume order (host side,
    variable length: an object gather, not a tensor collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local)
    world = dist.get_world_size()
    out = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(list(local), out, dst=dst)
    if dist.get_rank() != dst:
        return []
    merged: List = []
    for part in out:
        merged.extend(part)
    return merged


def bind_to_gpu_cpus(local_rank: int) -> Sequence[int]:
    """Pin this process to the CPUs NVML reports as local to its GPU (same NUMA node / PCIe root), so that pinned
    host staging buffers allocated afterwards are first-touched next to the GPU they feed: with 8 ranks copying
    67 MB per step each, cross-socket traffic otherwise halves the per-GPU host->device bandwidth.  Best effort:
    returns the CPU list it bound to, or () when NVML / affinity control is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = set(os.sched_getaffinity(0))
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return ()
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return ()
