#!/usr/bin/env python
"""HBM probes that bound the streaming kernels: pure write (memset), pure read (sum), copy, and a 1:2 read:write mix
(the stem's ratio: 67 MB in, 134 MB out), CUDA events, buffers >> L2."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e-3


def main():
    n = 512 * 1024 * 1024          # elements
    a = torch.empty(n, dtype=torch.bfloat16, device="cuda")   # 1 GiB
    b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    a.normal_()
    out = {}
    t = timed(lambda: b.zero_())
    out["memset_write_GBs"] = 2 * n / t / 1e9
    t = timed(lambda: b.fill_(1.5))
    out["fill_write_GBs"] = 2 * n / t / 1e9
    t = timed(lambda: b.copy_(a))
    out["copy_GBs"] = 4 * n / t / 1e9
    a32 = a.view(torch.float32)
    t = timed(lambda: a32.sum())
    out["read_sum_GBs"] = 2 * n / t / 1e9
    # 1 : 2 mix: read half of a, write all of b (repeat_interleave-free: two copies from the same half)
    h = n // 2
    def mix():
        b[:h].copy_(a[:h]); b[h:].fill_(0.5)
    t = timed(mix)
    out["mix_1r_2w_GBs"] = (2 * h + 2 * n) / t / 1e9
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_hbm.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
