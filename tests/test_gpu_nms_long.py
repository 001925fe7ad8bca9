"""GPU parity of the any-length detection stages: chunked greedy NMS, the long key sort and
``detect_objects`` with a ``top_k`` that defeats the ``10*top_k`` truncation (the NMS-stress setting of
model_insight.py:146 / SURVEY.md 8d C4, C5).

Keep masks and kept prior indices must be BIT-EXACT: against the oracle's greedy loop where the CPU finishes
in seconds, against the bit-matrix kernel (itself pinned by the oracle) at sizes beyond that, and through
size-independent properties (greedy prefix property, idempotence) at the full 2.5 M candidates.
"""
import numpy as np
import pytest
import torch

from oracle import ssd3d_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    from mslesions3d_b200 import ops
    return ops


def _dense_boxes(n, g, extent=0.3, side=(0.03, 0.08), dup=0):
    """Cubes crowded into [0, extent]^3 so that a good part of them is suppressed."""
    c = extent * torch.rand(n, 3, generator=g)
    s = (side[0] + (side[1] - side[0]) * torch.rand(n, 1, generator=g)).expand(n, 3)
    b = torch.cat([c - s / 2, c + s / 2], 1).contiguous()
    if dup:
        b[n - dup:] = b[:dup]          # exact duplicates: IoU == 1
    return b


@pytest.mark.parametrize("n", [1, 2, 1000, 16384, 16385, 40000, 100001, 1 << 20, 2501400])
def test_sort_keys_u64_matches_numpy(n):
    ops = _ops()
    rs = np.random.RandomState(n % 65521)
    keys = rs.randint(0, 1 << 62, size=n, dtype=np.int64).astype(np.uint64)
    keys[rs.rand(n) < 0.1] |= np.uint64(1 << 63)              # unsigned compare: top bit set sorts last
    if n > 10:
        keys[n // 2:n // 2 + n // 8] = keys[:n // 8]          # duplicates
    got = ops.sort_keys_u64(torch.from_numpy(keys.view(np.int64)).cuda()).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, np.sort(keys, kind="stable"))


def test_sort_keys_u64_is_stable_for_packed_candidate_keys():
    """{~orderable(score) << 32 | prior}: ascending key order = descending score, ascending prior on ties."""
    ops = _ops()
    n = 50000
    g = torch.Generator().manual_seed(3)
    score = torch.rand(n, generator=g)
    score[1000:9000] = score[1000]                              # 8000 exact ties
    bits = score.view(torch.int32).to(torch.int64) | 0x80000000  # orderable() of a positive float
    keys = (((~bits) & 0xFFFFFFFF) << 32) | torch.arange(n, dtype=torch.int64)
    perm = torch.randperm(n, generator=g)
    got = ops.sort_keys_u64(keys[perm].contiguous().cuda()).cpu()
    want = torch.sort(score, descending=True, stable=True)[1]
    assert torch.equal(got & 0xFFFFFFFF, want)
    assert torch.equal(ops._key_scores(got.cuda()).cpu(), score[want])


@pytest.mark.parametrize("n,thr,dup,chunk", [(1, 0.5, 0, 64), (63, 0.5, 0, 64), (64, 0.3, 4, 64), (65, 0.5, 0, 64),
                                             (1000, 0.5, 50, 64), (1000, 0.1, 0, 128), (4097, 0.45, 7, 256),
                                             (5000, 0.5, 100, 0), (12000, 0.5, 100, 1024), (12000, 0.3, 0, 4096),
                                             (20000, 0.5, 0, 0)])
def test_chunked_nms_keep_mask_bit_exact_vs_oracle(n, thr, dup, chunk):
    ops = _ops()
    g = torch.Generator().manual_seed(n + chunk)
    boxes = _dense_boxes(n, g, extent=0.3 if n >= 1000 else 0.15, dup=dup)
    want = O.greedy_nms(boxes, ops.f32(thr))
    for use_grid in (True, False):                          # grid-pruned and dense cross test
        keep, count = ops.nms3d_sorted_chunked(boxes.cuda(), thr, chunk, return_count=True, use_grid=use_grid)
        keep = keep.cpu()
        assert torch.equal(keep, want), "grid=%s: keep masks differ at %d positions" % (use_grid, int((keep != want).sum()))
        assert int(count.item()) == int(want.sum()) and bool(keep[0])
    if n >= 1000:
        assert 0.05 * n < int(want.sum()) < 0.95 * n       # the case really exercises suppression


@pytest.mark.parametrize("thr", [0.0, -0.25, 0.5])
def test_chunked_nms_threshold_edge_and_arbitrary_coordinates(thr):
    """thr = 0: touching boxes (intersection 0) survive; thr < 0: even disjoint boxes suppress each other (the
    grid must not be used); coordinates far outside [0, 1] (voxel units, negative offsets) map onto the grid."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    boxes = _dense_boxes(6000, g, extent=0.3) * 37.0 - 11.0
    boxes[100:200] = boxes[:100]                            # exact duplicates
    lattice = torch.arange(0, 500, dtype=torch.float32)[:, None] * torch.tensor([1., 0, 0, 1., 0, 0]) + \
        torch.tensor([40., 40, 40, 41, 41, 41])             # 500 unit cubes in a row, each touching the next
    boxes = torch.cat([boxes, lattice]).contiguous()
    want = O.greedy_nms(boxes, ops.f32(thr))
    keep = ops.nms3d_sorted_chunked(boxes.cuda(), thr, 512).cpu()
    assert torch.equal(keep, want), "%d differences" % int((keep != want).sum())
    if thr == 0.0:
        assert bool(want[-500:].all())
    if thr < 0:
        assert int(want.sum()) == 1


def test_chunked_nms_matches_bit_matrix_at_100k():
    ops = _ops()
    g = torch.Generator().manual_seed(7)
    boxes = _dense_boxes(100000, g, extent=0.5, dup=500).cuda()
    want = ops.nms3d_sorted(boxes, 0.5)                     # pinned by the oracle in test_gpu_detect.py
    for chunk, use_grid in ((0, True), (16384, True), (1984, True), (4096, False)):
        keep = ops.nms3d_sorted_chunked(boxes, 0.5, chunk, use_grid=use_grid)
        assert torch.equal(keep, want), "chunk %d grid %s: %d differences" % (chunk, use_grid, int((keep != want).sum()))
    assert 0.05 < float(want.float().mean()) < 0.95


def test_chunked_nms_full_size_properties():
    """2.5 M candidates (every prior of the whole-brain config with the layer-0 head): greedy prefix property
    against the bit-matrix kernel, idempotence on the kept set, count consistency."""
    ops = _ops()
    n = 2501400
    g = torch.Generator().manual_seed(11)
    c = torch.rand(n, 3, generator=g)
    s = (0.02 + 0.08 * torch.rand(n, 1, generator=g)).expand(n, 3)
    boxes = torch.cat([c - s / 2, c + s / 2], 1).contiguous().cuda()
    keep, count = ops.nms3d_sorted_chunked(boxes, 0.5, return_count=True)
    kept = int(keep.sum().item())
    assert int(count.item()) == kept and 0 < kept < n and bool(keep[0])
    dense = ops.nms3d_sorted_chunked(boxes, 0.5, 4096, use_grid=False)
    assert torch.equal(keep, dense), "grid-pruned and dense cross tests differ at %d boxes" % int((keep != dense).sum())
    m = 50000                                               # the first m decisions do not depend on the rest
    assert torch.equal(keep[:m], ops.nms3d_sorted(boxes[:m].contiguous(), 0.5))
    again = ops.nms3d_sorted_chunked(boxes[keep].contiguous(), 0.5)
    assert bool(again.all()), "NMS of the kept set removed %d more boxes" % int((~again).sum())
    # every removed box among a sample has an earlier kept box above the threshold (exact IoU, utils.py:149)
    from mslesions3d_b200 import utils as U
    removed = (~keep).nonzero().flatten()
    sample = removed[torch.linspace(0, removed.numel() - 1, 64).long()]
    kept_idx = keep.nonzero().flatten()
    kept_boxes = boxes[kept_idx]
    iou = U.find_jaccard_overlap3d(boxes[sample].contiguous(), kept_boxes.contiguous())
    earlier = kept_idx[None, :] < sample[:, None]
    assert bool(((iou > 0.5) & earlier).any(1).all())


def _check_long_detect(locs, scores, priors, min_score, max_overlap, top_k, chunk=0):
    ops = _ops()
    assert ops.detect_needs_long_lists(priors.shape[0], top_k)
    b, l, s, idx = ops.detect_objects_long(locs.cuda(), scores.cuda(), priors.cuda(), min_score, max_overlap, top_k,
                                           return_prior=True, chunk=chunk)
    probs, boxes = ops.decode_softmax(locs.cuda(), scores.cuda(), priors.cuda())
    wb, wl, ws, widx = O.detect_from_decoded(probs.cpu(), boxes.cpu(), ops.f32(min_score), ops.f32(max_overlap), top_k,
                                             return_indices=True)
    for i in range(locs.shape[0]):
        assert torch.equal(idx[i].cpu(), widx[i]), "image %d: kept prior indices differ" % i
        assert torch.equal(l[i].cpu(), wl[i])
        assert torch.equal(s[i].cpu(), ws[i])
        assert torch.equal(b[i].cpu(), wb[i])
    return b, l, s


def _crowded_priors(P, g):
    c = 0.2 + 0.35 * torch.rand(P, 3, generator=g)
    s = (0.03 + 0.05 * torch.rand(P, 1, generator=g)).expand(P, 3)
    return torch.cat([c, s], 1).contiguous()


def test_detect_long_lists_exact_vs_oracle():
    """P > 16384 and 10*top_k > 8192: the setting the fused kernel refuses.  min_score = 0 keeps every prior;
    the second image has hundreds of exact score ties; top_k cuts the merged list."""
    g = torch.Generator().manual_seed(21)
    P = 20000
    priors = _crowded_priors(P, g)
    locs = torch.randn(2, P, 6, generator=g) * 0.3
    scores = torch.randn(2, P, 2, generator=g) * 2
    scores[1, 300:900] = scores[1, 300:301]
    _check_long_detect(locs, scores, priors, 0.0, 0.5, 1500)          # list truncated to 15000, > top_k kept
    _check_long_detect(locs, scores, priors, 0.0, 0.45, 4000, chunk=1024)   # no truncation


def test_detect_long_lists_three_classes_and_empty_image():
    g = torch.Generator().manual_seed(22)
    P = 17000
    priors = _crowded_priors(P, g)
    locs = torch.randn(2, P, 6, generator=g) * 0.3
    scores = torch.randn(2, P, 3, generator=g) * 2
    scores[1, :, 0] = 50.0                                  # image 1: background wins everywhere -> placeholder
    b, l, s = _check_long_detect(locs, scores, priors, 0.3, 0.5, 1000)
    assert l[1].tolist() == [0] and s[1].tolist() == [0.0] and b[1].tolist() == [[0., 0., 0., 1., 1., 1.]]
    assert set(l[0].tolist()) == {1, 2}


def test_model_detect_objects_routes_long_lists():
    """LSSD3D.detect_objects / predict_step with the model_insight.py:146 setting (min_score=0, top_k=50000)
    on a whole-brain-style model (layer-0 head): every candidate goes through NMS; same detections from both entry points."""
    from mslesions3d_b200.ssd3d import LSSD3D
    from mslesions3d_b200 import synthetic
    size = (32, 48, 32)                                     # non-cubic: stem stride (1,2,2), 24 576 layer-0 priors
    ar = {0: [1.], 3: [1.], 5: [1.], 7: [1.]}
    m = LSSD3D(n_classes=2, input_channels=1, input_size=size, aspect_ratios=ar, min_score=0.0, top_k=50000)
    m.load_state_dict(synthetic.random_state_dict(1, ar, seed=0))
    m = m.cuda().eval()
    P = m.priors_cxcycz.shape[0]
    assert _ops().detect_needs_long_lists(P, m.top_k)
    vols = torch.from_numpy(synthetic.make_batch(1, 1, size)).cuda()
    with torch.no_grad():
        locs, scores = m(vols)
        b, l, s, idx = m.detect_objects(locs, scores, 0.0, 0.5, 50000, return_prior=True)
        pb, pl, ps = m.predict_step({"img": vols}, 0)
    probs, boxes = _ops().decode_softmax(locs, scores, m.priors_cxcycz.cuda())
    wb, wl, ws, widx = O.detect_from_decoded(probs.cpu(), boxes.cpu(), 0.0, _ops().f32(0.5), 50000, return_indices=True)
    assert torch.equal(idx[0].cpu(), widx[0]) and torch.equal(s[0].cpu(), ws[0]) and torch.equal(b[0].cpu(), wb[0])
    assert torch.equal(pb[0], b[0]) and torch.equal(pl[0], l[0]) and torch.equal(ps[0], s[0])


def _slice_cases(thr):
    """Adversarial pairs for the pruned search range: o is a slab of a (so IoU = the slab fraction f) flush with
    one face of a, f a few ulp either side of thr.  All slabs come first (kept), the big boxes later."""
    g = torch.Generator().manual_seed(int(thr * 1000) + 1)
    n = 600
    idx = torch.arange(n)
    lo = torch.stack([(idx % 10) * 3.0, ((idx // 10) % 10) * 3.0, (idx // 100) * 3.0], 1) + torch.rand(n, 3, generator=g)
    ext = 0.5 + 1.5 * torch.rand(n, 3, generator=g)
    a = torch.cat([lo, lo + ext], 1)
    f = thr * (1.0 + (torch.randint(-3, 4, (n,), generator=g).float() * 1.5e-6))
    axis = torch.randint(0, 3, (n,), generator=g)
    right = torch.rand(n, generator=g) < 0.5
    o = a.clone()
    for i in range(n):
        k = int(axis[i])
        if bool(right[i]):
            o[i, k] = a[i, 3 + k] - f[i] * ext[i, k]
        else:
            o[i, 3 + k] = a[i, k] + f[i] * ext[i, k]
    return torch.cat([o, a]).contiguous()


@pytest.mark.parametrize("thr", [0.5, 0.25, 0.7])
def test_chunked_nms_pruned_range_is_conservative_at_the_threshold(thr):
    ops = _ops()
    boxes = _slice_cases(thr)
    want = O.greedy_nms(boxes, ops.f32(thr))
    n_removed = int((~want).sum())
    assert 100 < n_removed < 500                            # both outcomes occur
    for chunk in (64, 256):
        keep = ops.nms3d_sorted_chunked(boxes.cuda(), thr, chunk).cpu()
        assert torch.equal(keep, want), "chunk %d: %d differences" % (chunk, int((keep != want).sum()))
