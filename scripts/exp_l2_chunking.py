#!/usr/bin/env python
"""Experiment: does running stem -> first depthwise conv per sub-batch keep the stem output in L2?
Times stem(8) + dw(8) against [stem(k) + dw(k)] x (8/k) in a captured graph, inputs rotating over > L2."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import _lib, ops, synthetic  # noqa: E402


def main():
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    n, cin, s = 8, 2, 128
    xs = [torch.randn((n, cin, s, s, s), device=dev).to(torch.bfloat16) for _ in range(4)]
    w = ops.pack_stem_weight(torch.randn(32, cin, 3, 3, 3, device=dev) * 0.1)
    sc, sh = torch.rand(32, device=dev) + 0.5, torch.randn(32, device=dev) * 0.1
    wd = ops.pack_dw_weight(torch.randn(32, 1, 3, 3, 3, device=dev) * 0.2)
    so = s // 2
    y = torch.empty((n, so, so, so, 32), dtype=torch.bfloat16, device=dev)
    z = torch.empty((n, so // 2, so // 2, so // 2, 32), dtype=torch.bfloat16, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run(x, k):
        for i in range(0, n, k):
            xi, yi, zi = x[i:i + k], y[i:i + k], z[i:i + k]
            rc = lib.ssd3d_stem_conv_bn_relu(xi.data_ptr(), 1, w.data_ptr(), sc.data_ptr(), sh.data_ptr(), yi.data_ptr(),
                                             k, cin, s, s, s, 2, st)
            assert rc == 0
            rc = lib.ssd3d_dwconv3d_bn_relu(yi.data_ptr(), wd.data_ptr(), sc.data_ptr(), sh.data_ptr(), zi.data_ptr(),
                                            k, 32, so, so, so, 2, st)
            assert rc == 0

    ref = None
    for k in (8, 4, 2, 1):
        for x in xs:
            run(x, k)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 25
        e0.record()
        for r in range(reps):
            for x in xs:
                run(x, k)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1000 / (reps * len(xs))
        chk = float(z.float().sum())
        ref = chk if ref is None else ref
        print("sub-batch %d: %.1f us per batch of %d  (checksum %s)" % (k, us, n, "same" if chk == ref else "DIFFERENT"))


if __name__ == "__main__":
    main()
