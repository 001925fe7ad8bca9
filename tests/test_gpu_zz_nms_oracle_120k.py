"""Any-length greedy NMS against an independent CPU implementation at 120 000 candidates.

Kept in its own file, collected after every other GPU suite: it was added after the round's GPU budget was spent
(see DESIGN.md section 10), and the CPU oracle alone takes ~30 s."""
import pytest
import torch

from oracle import ssd3d_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from mslesions3d_b200 import ops
    return ops


def test_chunked_nms_exact_vs_cpu_oracle_at_120k():
    """120 000 candidates against an independent CPU implementation of the reference's greedy loop (spatially hashed,
    exact fp32 IoU arithmetic; tests/test_oracle_golden.py checks it against the n x n restatement)."""
    ops = _ops()
    n = 120000
    g = torch.Generator().manual_seed(n)
    c = 0.6 * torch.rand(n, 3, generator=g)
    s = (0.02 + 0.03 * torch.rand(n, 1, generator=g)).expand(n, 3)
    boxes = torch.cat([c - s / 2, c + s / 2], 1).contiguous()
    want = O.greedy_nms_grid(boxes, ops.f32(0.5))
    keep, count = ops.nms3d_sorted_chunked(boxes.cuda(), 0.5, return_count=True)
    keep = keep.cpu()
    assert torch.equal(keep, want), "keep masks differ at %d positions" % int((keep != want).sum())
    assert int(count.item()) == int(want.sum()) and 0.05 * n < int((~want).sum()) < 0.95 * n
