#!/usr/bin/env python
"""Fused Block kernel vs the two stand-alone kernels at the benchmark's f1 / f2 / f3 shapes (batch 8 of 128^3):
CUDA-event time of back-to-back launches over rotating inputs (> L2).  With --nvtx the timed launches of the fused
kernel sit in an NVTX range "prof" (for ncu --nvtx --nvtx-include "prof/")."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mslesions3d_b200 import ops  # noqa: E402

SHAPES = [("f1", 32, 64, 2, (64, 64, 64)), ("f2", 64, 128, 2, (32, 32, 32)), ("f3", 128, 128, 1, (16, 16, 16))]


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
    for name, cin, cout, s, size in SHAPES:
        if only and name not in only:
            continue
        n_rot = max(2, int(300e6 // (8 * cin * size[0] * size[1] * size[2] * 2)) + 1)
        xs = [torch.randn((8,) + size + (cin,), device="cuda", generator=g).to(torch.bfloat16).permute(0, 4, 1, 2, 3)
              for _ in range(n_rot)]
        wd = (torch.randn((27, cin), device="cuda", generator=g) * 0.25).to(torch.bfloat16)
        wp = (torch.randn((cout, cin), device="cuda", generator=g) / cin ** 0.5).to(torch.bfloat16)
        s1, b1 = torch.rand(cin, device="cuda") + 0.5, torch.randn(cin, device="cuda") * 0.1
        s2, b2 = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda") * 0.1
        flag = torch.zeros(1, dtype=torch.int32, device="cuda")
        n = 10 if nvtx else 50
        mids = [ops.dwconv3d_bn_relu(x, wd, s1, b1, s) for x in xs]
        if not nvtx:
            t_dw = timed(lambda i: ops.dwconv3d_bn_relu(xs[i % n_rot], wd, s1, b1, s), n)
            t_pw = timed(lambda i: ops.pwconv_bn_relu(mids[i % n_rot], wp, s2, b2, flag), n)
        if nvtx:
            torch.cuda.nvtx.range_push("prof")
        t_f = timed(lambda i: ops.block_dwpw_bn_relu(xs[i % n_rot], wd, s1, b1, wp, s2, b2, s, flag), n)
        if nvtx:
            torch.cuda.nvtx.range_pop()
            continue
        vin = 8 * size[0] * size[1] * size[2]
        vout = vin // (s ** 3)
        fused_bytes = 2 * (cin * vin + cout * vout)
        out[name] = {"dw_us": t_dw, "pw_us": t_pw, "fused_us": t_f, "fused_GBs": fused_bytes / t_f / 1e3,
                     "fused_frac_hbm": fused_bytes / t_f / 1e3 / 6439.5, "algorithmic_MB": fused_bytes / 1e6}
        print(name, json.dumps(out[name]), flush=True)
    if not nvtx:
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_block.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
