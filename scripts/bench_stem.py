#!/usr/bin/env python
"""Stem kernels side by side: banded-B tcgen05 GEMM on raw TMA rows ("tz", csrc/conv_stem_tz.cu) vs the
gather-based tcgen05 implicit GEMM ("tc", csrc/conv_stem_tc.cu) at the benchmark (2ch 128^3, batch 8) and training
(1ch 96^3, batch 16) shapes.  CUDA-event time of back-to-back launches over rotating inputs and outputs (> L2),
max |tz - tc| over the outputs (both round the same fp32 sums: expect 0 or 1 bf16 ulp on a few elements).
With --nvtx the timed tz launches sit in an NVTX range "prof" (ncu --nvtx --nvtx-include "prof/")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mslesions3d_b200 import ops  # noqa: E402

SHAPES = [("C2", 8, 2, (128, 128, 128), 2), ("C3", 16, 1, (96, 96, 96), 2)]


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
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    out = {}
    g = torch.Generator(device="cuda").manual_seed(0)
    for name, batch, cin, size, sd in SHAPES:
        if only and name not in only:
            continue
        vin = batch * size[0] * size[1] * size[2]
        vout = vin // (sd * 4)
        n_rot = 4
        xs = [torch.randn((batch, cin) + size, device="cuda", generator=g).to(torch.bfloat16) for _ in range(n_rot)]
        w = ops.pack_stem_weight(torch.randn((32, cin, 3, 3, 3), device="cuda", generator=g) * 0.2)
        sc, sh = torch.rand(32, device="cuda") + 0.5, torch.randn(32, device="cuda") * 0.1
        ys = [ops.stem_conv_bn_relu(xs[i], w, sc, sh, sd, kernel="tc") for i in range(n_rot)]
        n = 10 if nvtx else 40
        res = {}
        if not nvtx:
            res["tc_us"] = timed(lambda i: ops.stem_conv_bn_relu(xs[i % n_rot], w, sc, sh, sd, out=ys[i % n_rot], kernel="tc"), n)
            ref = [y.clone() for y in ys]
        if nvtx:
            torch.cuda.nvtx.range_push("prof")
        res["tz_us"] = timed(lambda i: ops.stem_conv_bn_relu(xs[i % n_rot], w, sc, sh, sd, out=ys[i % n_rot], kernel="tz"), n)
        if nvtx:
            torch.cuda.nvtx.range_pop()
            continue
        diff = max(float((ys[i].float() - ref[i].float()).abs().max()) for i in range(n_rot))
        nbytes = 2 * (cin * vin + 32 * vout)
        res.update(algorithmic_MB=nbytes / 1e6, tz_GBs=nbytes / res["tz_us"] / 1e3, tc_GBs=nbytes / res["tc_us"] / 1e3,
                   max_abs_diff_tz_vs_tc=diff)
        out[name] = res
        print("%s: tc %.1f us  tz %.1f us  (%.0f / %.0f GB/s)  max |tz - tc| %.4g" % (
            name, res["tc_us"], res["tz_us"], res["tc_GBs"], res["tz_GBs"], diff), flush=True)
    if not nvtx:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_stem.json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
