#!/usr/bin/env python
"""Fused stem + first depthwise conv (csrc/conv_stem_dw.cu) vs the two stand-alone kernels at the benchmark shape
(2ch 128^3, batch 8): CUDA-event time of back-to-back launches over rotating inputs (> L2).  With --nvtx the timed
fused launches sit in an NVTX range "prof"."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mslesions3d_b200 import ops  # noqa: E402


def timed(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return 1000.0 * a.elapsed_time(b) / n


def main():
    nvtx = "--nvtx" in sys.argv
    g = torch.Generator(device="cuda").manual_seed(0)
    batch, cin, size, sd = 8, 2, (128, 128, 128), 2
    n_rot = 4
    xs = [torch.randn((batch, cin) + size, device="cuda", generator=g).to(torch.bfloat16) for _ in range(n_rot)]
    ws = ops.pack_stem_weight(torch.randn((32, cin, 3, 3, 3), device="cuda", generator=g) * 0.2)
    wd = ops.pack_dw_weight(torch.randn((32, 1, 3, 3, 3), device="cuda", generator=g) * 0.3)
    sc0, sh0 = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda") * 0.1
    sc1, sh1 = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda") * 0.1
    mids = [ops.stem_conv_bn_relu(x, ws, sc0, sh0, sd, kernel="tc") for x in xs]
    outs = [ops.stem_dw_bn_relu(x, ws, sc0, sh0, wd, sc1, sh1, sd) for x in xs]
    n = 10 if nvtx else 40
    res = {}
    if not nvtx:
        res["stem_us"] = timed(lambda i: ops.stem_conv_bn_relu(xs[i % n_rot], ws, sc0, sh0, sd, out=mids[i % n_rot], kernel="tc"), n)
        res["dw_us"] = timed(lambda i: ops.dwconv3d_bn_relu(mids[i % n_rot], wd, sc1, sh1, 2), n)
    if nvtx:
        torch.cuda.nvtx.range_push("prof")
    res["fused_us"] = timed(lambda i: ops.stem_dw_bn_relu(xs[i % n_rot], ws, sc0, sh0, wd, sc1, sh1, sd, out=outs[i % n_rot]), n)
    if nvtx:
        torch.cuda.nvtx.range_pop()
        return
    vin = batch * size[0] * size[1] * size[2]
    nbytes = 2 * (cin * vin + 32 * vin // 64)
    res.update(algorithmic_MB=nbytes / 1e6, fused_GBs=nbytes / res["fused_us"] / 1e3,
               unfused_MB=2 * (cin * vin + 2 * 32 * vin // 8 + 32 * vin // 64) / 1e6)
    print("stem %.1f us + dw %.1f us = %.1f us; fused %.1f us (%.0f GB/s of its %.1f MB)" % (
        res["stem_us"], res["dw_us"], res["stem_us"] + res["dw_us"], res["fused_us"], res["fused_GBs"], res["algorithmic_MB"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "bench_stem_dw.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
