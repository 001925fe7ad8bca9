#!/usr/bin/env python
"""BASELINE.json configs[4], NMS part WITHOUT the 10*top_k truncation: greedy 3-D NMS over ALL n score-sorted
candidates (n = 1 k ... 2.5 M cubic boxes of side 0.02-0.1 at uniform centres, threshold 0.5),
plus the 64-bit key sort that orders them.  The reference materialises an n x n IoU matrix (ssd3d.py:407):
4 n^2 bytes = 1 GB at n = 16 k, 25 TB at 2.5 M, so its torch-CPU port is timed up to n = 16 k only.

    python scripts/sweep_nms_long.py [--json out.json] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import ops  # noqa: E402
from oracle import ssd3d_oracle as O  # noqa: E402  (CPU reference timing only)

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--max-n", type=int, default=2500000)
args = ap.parse_args()
dev = torch.device("cuda")
peak = 6439.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]


def gpu_us(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {"peak_gbs": peak, "nms_full": [], "sort": []}
torch.set_num_threads(min(16, os.cpu_count() or 1))
print("== greedy 3-D NMS over all n sorted candidates (no truncation), threshold 0.5 ==")
for n in (1000, 4000, 16000, 64000, 256000, 1000000, 2500000):
    if n > args.max_n:
        break
    gg = torch.Generator().manual_seed(n)
    side = 0.02 + 0.08 * torch.rand(n, 1, generator=gg)
    ctr = torch.rand(n, 3, generator=gg)
    boxes_cpu = torch.cat([ctr - side / 2, ctr + side / 2], 1).contiguous()
    boxes = boxes_cpu.to(dev)
    keep, cnt = ops.nms3d_sorted_chunked(boxes, 0.5, return_count=True)
    kept = int(cnt.item())
    us = gpu_us(lambda: ops.nms3d_sorted_chunked(boxes, 0.5), args.reps)
    row = dict(n=n, us=us, kept=kept, candidates_per_s=n / us * 1e6, algorithmic_MB=n * 36 / 1e6,
               GBs=n * 36 / us / 1e3, frac_hbm=n * 36 / us / 1e3 / peak,
               iou_tests_upper_bound=float(n) * kept / 2, reference_iou_matrix_GB=4.0 * n * n / 1e9)
    if n >= 64000:
        row["dense_cross_us"] = gpu_us(lambda: ops.nms3d_sorted_chunked(boxes, 0.5, 4096, use_grid=False), 1)
        for ch in (4096, 8192, 16384):
            row["grid_chunk%d_us" % ch] = gpu_us(lambda: ops.nms3d_sorted_chunked(boxes, 0.5, ch), args.reps)
    if n <= 131072:
        row["bit_matrix_us"] = gpu_us(lambda: ops.nms3d_sorted(boxes, 0.5), args.reps)
    if not args.no_cpu and n <= 16000:
        t0 = time.perf_counter()
        want = O.greedy_nms(boxes_cpu, ops.f32(0.5))
        row["cpu_us"] = (time.perf_counter() - t0) * 1e6
        row["exact_vs_cpu"] = bool(torch.equal(want, keep.cpu()))
    out["nms_full"].append(row)
    print("n=%8d  %11.1f us  kept %8d  %.3g cand/s  bit-matrix %s us  cpu %s us  %s  %s" % (
        n, us, kept, row["candidates_per_s"], "%.0f" % row["bit_matrix_us"] if "bit_matrix_us" in row else "-",
        "%.0f" % row["cpu_us"] if "cpu_us" in row else "-", row.get("exact_vs_cpu", ""),
        " ".join("%s=%.0f" % (k, v) for k, v in row.items() if k.startswith(("grid_chunk", "dense_cross")))), flush=True)
    del boxes, keep

print("== 64-bit key sort (block bitonic + merge passes) ==")
for n in (16384, 262144, 2500000):
    if n > args.max_n:
        break
    keys = torch.randint(0, 1 << 62, (n,), dtype=torch.int64, device=dev)
    work = torch.empty_like(keys)

    def run():
        work.copy_(keys)
        ops.sort_keys_u64(work)
    us = gpu_us(run, args.reps)
    us_copy = gpu_us(lambda: work.copy_(keys), args.reps)
    us_torch = gpu_us(lambda: torch.sort(keys), args.reps)
    row = dict(n=n, us=us - us_copy, torch_sort_us=us_torch, keys_per_s=n / max(us - us_copy, 1e-3) * 1e6)
    out["sort"].append(row)
    print("n=%8d  %9.1f us  (torch.sort %9.1f us)  %.3g keys/s" % (n, row["us"], us_torch, row["keys_per_s"]), flush=True)
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
