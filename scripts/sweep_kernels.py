#!/usr/bin/env python
"""Kernel sweeps of BASELINE.json configs[4]: depthwise 3-D conv over channels / volume size / stride, and 3-D
NMS (detect_objects: filter + exact top-10*top_k + greedy NMS) over the anchor count, each against the HBM
roofline and against the reference's torch-CPU ops on the same shapes (bounded: batch 1 / small n for the CPU).

    python scripts/sweep_kernels.py [--json out.json] [--no-cpu]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import ops  # noqa: E402
from oracle import ssd3d_oracle as O  # noqa: E402  (CPU reference timing only)

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--no-cpu", action="store_true")
ap.add_argument("--reps", type=int, default=10)
args = ap.parse_args()
dev = torch.device("cuda")
peak = 6439.5
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(pk):
    peak = json.load(open(pk))["hbm_gbs"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def gpu_us(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


out = {"peak_gbs": peak, "depthwise": [], "nms": []}
torch.set_num_threads(min(16, os.cpu_count() or 1))
print("== depthwise 3x3x3 + BN + ReLU, batch 8, bf16 channels-last ==")
g = torch.Generator().manual_seed(0)
for stride in (1, 2):
    for c in (32, 64, 128, 256, 512):
        for vo in (4, 8, 16, 32, 64):
            vi = vo * stride
            n = 8
            in_bytes = n * c * vi ** 3 * 2
            if in_bytes > 6e9:
                continue
            x = torch.randn((n, vi, vi, vi, c), device=dev, dtype=torch.bfloat16).permute(0, 4, 1, 2, 3)
            w = (torch.randn((c, 1, 3, 3, 3), generator=g) * 0.3).to(dev)
            wd = ops.pack_dw_weight(w)
            sc, sh = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            us = gpu_us(lambda: ops.dwconv3d_bn_relu(x, wd, sc, sh, stride), args.reps)
            nbytes = in_bytes + n * c * vo ** 3 * 2 + 54 * c
            row = dict(C=c, out=vo, stride=stride, us=us, MB=nbytes / 1e6, GBs=nbytes / us / 1e3,
                       frac_hbm=nbytes / us / 1e3 / peak)
            if not args.no_cpu and in_bytes / n <= 3e8:
                xc = x[:1].float().cpu().contiguous()
                wc = w.cpu()
                F.conv3d(xc, wc, None, stride, 1, 1, c)
                t0 = time.perf_counter()
                F.relu(F.conv3d(xc, wc, None, stride, 1, 1, c))
                row["cpu_us_per_volume"] = (time.perf_counter() - t0) * 1e6
                row["gpu_us_per_volume"] = us / n
            out["depthwise"].append(row)
            print("C=%3d out %2d^3 s%d  %8.1f us  %8.1f MB  %6.0f GB/s  %5.1f%% HBM  cpu/vol %s us" % (
                c, vo, stride, us, nbytes / 1e6, row["GBs"], 100 * row["frac_hbm"],
                "%.0f" % row["cpu_us_per_volume"] if "cpu_us_per_volume" in row else "-"), flush=True)
            del x

print("== detect_objects (decode + filter + exact top-8000 + greedy 3-D NMS), 1 image, min_score 0, top_k 800 ==")
for n in (1000, 4000, 16000, 64000, 256000, 1000000, 2500000):
    gg = torch.Generator().manual_seed(n)
    side = 0.02 + 0.08 * torch.rand(n, 1, generator=gg)
    ctr = torch.rand(n, 3, generator=gg)
    priors = torch.cat([ctr, side.expand(n, 3)], 1).clamp(0, 1).to(dev)
    locs = (torch.randn(1, n, 6, generator=gg) * 0.2).to(dev)
    scores = (torch.randn(1, n, 2, generator=gg) * 1.5).to(dev)
    top_k = 800
    us = gpu_us(lambda: ops.detect_objects_padded(locs, scores, priors, 0.0, 0.5, top_k), args.reps)
    nbytes = n * (24 + 8 + 24) + n * 28
    row = dict(n=n, us=us, MB=nbytes / 1e6, GBs=nbytes / us / 1e3, frac_hbm=nbytes / us / 1e3 / peak,
               candidates_per_s=n / us * 1e6, nms_boxes=min(n, 10 * top_k))
    if not args.no_cpu and n <= 64000:
        t0 = time.perf_counter()
        O.detect_objects(locs.cpu(), scores.cpu(), priors.cpu(), 0.0, 0.5, top_k)
        row["cpu_us"] = (time.perf_counter() - t0) * 1e6
    out["nms"].append(row)
    print("n=%8d  %9.1f us  %7.1f MB  %6.1f GB/s  %5.2f%% HBM  %.3g cand/s  cpu %s us" % (
        n, us, nbytes / 1e6, row["GBs"], 100 * row["frac_hbm"], row["candidates_per_s"],
        "%.0f" % row["cpu_us"] if "cpu_us" in row else "-"), flush=True)
if args.json:
    json.dump(out, open(args.json, "w"), indent=1)
