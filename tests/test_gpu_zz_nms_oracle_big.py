"""Any-length greedy NMS at the C4 / C5 sizes (1 M and 2.5 M candidates) against an INDEPENDENT implementation:
the C restatement of the reference's loop (oracle/nms_oracle.c, itself pinned against the n x n restatement and
the reference's own loop on the CPU), and a run-to-run determinism check -- the kernel's cross test reads kept
bits while another CTA sets them, so identical masks over repeated runs are demanded, not
assumed.  Collected last (file name); the CPU oracle needs ~15 s at 1 M and ~60 s at 2.5 M.
Set SSD3D_SKIP_SLOW=1 to skip the 2.5 M cases while iterating."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.slow]


def _boxes(n):
    g = torch.Generator().manual_seed(n)
    c = torch.rand(n, 3, generator=g)
    s = 0.02 + 0.08 * torch.rand(n, 1, generator=g)
    return torch.cat([c - s / 2, c + s / 2], 1).contiguous()       # SURVEY 8d C5 boxes


@pytest.mark.parametrize("n", [1000000, 2500000])
def test_chunked_nms_exact_vs_c_oracle(n):
    from mslesions3d_b200 import ops
    from oracle import nms_oracle
    if n > 1000000 and os.environ.get("SSD3D_SKIP_SLOW") == "1":
        pytest.skip("SSD3D_SKIP_SLOW=1")
    boxes = _boxes(n)
    want = torch.from_numpy(nms_oracle.greedy_nms(boxes.numpy(), 0.5))
    keep, count = ops.nms3d_sorted_chunked(boxes.cuda(), 0.5, return_count=True)
    keep = keep.cpu()
    assert torch.equal(keep, want), "keep masks differ at %d of %d positions" % (int((keep != want).sum()), n)
    assert int(count.item()) == int(want.sum()) and 0.05 * n < int((~want).sum()) < 0.95 * n


def test_chunked_nms_is_deterministic_at_2_5m():
    """Five runs over the same 2.5 M candidates, interleaved with other work that perturbs CTA scheduling: the keep
    masks must be bit-identical (and equal to the run the oracle test above compares)."""
    from mslesions3d_b200 import ops
    n = 2500000 if os.environ.get("SSD3D_SKIP_SLOW") != "1" else 600000
    boxes = _boxes(n).cuda()
    first = None
    noise = torch.empty((64 << 20,), dtype=torch.float32, device="cuda")
    side = torch.cuda.Stream()
    for it in range(5):
        if it % 2:                         # a competing memory-bound kernel on another stream
            with torch.cuda.stream(side):
                noise.fill_(float(it))
        keep, count = ops.nms3d_sorted_chunked(boxes, 0.5, return_count=True)
        torch.cuda.synchronize()
        if first is None:
            first = (keep.clone(), int(count.item()))
        else:
            assert int(count.item()) == first[1]
            assert torch.equal(keep, first[0]), "run %d differs at %d positions" % (it, int((keep != first[0]).sum()))
