#!/usr/bin/env python
"""One any-length NMS call for an ncu launch list (per-kernel share of the chunked NMS):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python scripts/prof_nms_long.py [n]
Same boxes as scripts/sweep_nms_long.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mslesions3d_b200 import ops  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2500000
gg = torch.Generator().manual_seed(n)
side = 0.02 + 0.08 * torch.rand(n, 1, generator=gg)
ctr = torch.rand(n, 3, generator=gg)
boxes = torch.cat([ctr - side / 2, ctr + side / 2], 1).contiguous().cuda()
keep, cnt = ops.nms3d_sorted_chunked(boxes, 0.5, return_count=True)
print("n", n, "kept", int(cnt.item()))
