"""LIVE comparison of the CPU oracle with the UNMODIFIED reference, imported through ``oracle/ref_shim.py``
(from ``/root/reference`` in the build container, else from the byte-identical files ``oracle/make_ref.py``
staged under ``oracle/_ref/``).  Skipped where neither exists.  The committed goldens
(``test_oracle_golden.py``) cover the same ground anywhere; this file re-derives them on the spot with fresh
seeds the golden set does not contain, so an oracle that merely memorised the goldens would fail here."""
import pytest
import torch

from oracle import ssd3d_oracle as O
from oracle import ref_shim
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference sources not present")

torch.set_num_threads(1)


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load_reference()


def _ref_model(ssd3d, channels, size, **kw):
    torch.manual_seed(0)
    return ssd3d.LSSD3D(n_classes=kw.pop("n_classes", 2), input_channels=channels, input_size=tuple(size), **kw).eval()


def test_staged_copy_is_byte_identical_to_the_mount():
    import os
    from oracle import make_ref
    if not os.path.isdir(make_ref.SOURCE_DIR):
        pytest.skip("no reference mount: nothing to compare the staged copy with")
    make_ref.stage()
    assert make_ref.verify()
    for name in make_ref.FILES:
        a = open(os.path.join(make_ref.SOURCE_DIR, name), "rb").read()
        b = open(os.path.join(make_ref.REF_DIR, name), "rb").read()
        assert a == b, name


@pytest.mark.parametrize("size,channels,ar", [((48, 48, 48), 1, {}), ((24, 40, 32), 2, {}),
                                              ((16, 32, 16), 1, {0: [1.], 3: [1.], 5: [1.]})])
def test_priors_live(ref, size, channels, ar):
    m = _ref_model(ref[0], channels, size, aspect_ratios=ar)
    want = m.priors_cxcycz
    got = O.prior_boxes(size, ar or None, in_channels=channels)
    assert torch.equal(got, want)
    assert torch.equal(O.prior_boxes_fast(size, ar or None, in_channels=channels), want)


@pytest.mark.parametrize("seed,size,channels", [(31, (48, 48, 48), 1), (32, (40, 64, 48), 2)])
def test_forward_live(ref, seed, size, channels):
    case = dict(channels=channels, size=size, batch=2, seed=seed)
    sd, x = GI.forward_inputs(case)
    m = _ref_model(ref[0], channels, size)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        wl, ws = m(x)
        gl, gs = O.forward(sd, x)
    torch.testing.assert_close(gl, wl, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(gs, ws, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("seed,min_score,max_overlap,top_k,n_classes", [(41, 0.5, 0.5, 100, 2), (42, 0.0, 0.4, 30, 2),
                                                                        (43, 0.35, 0.45, 80, 3)])
def test_detect_objects_live(ref, seed, min_score, max_overlap, top_k, n_classes):
    size = (64, 64, 64)
    m = _ref_model(ref[0], 1, size, n_classes=n_classes)
    case = dict(batch=2, seed=seed, n_classes=n_classes)
    locs, scores = GI.detect_inputs(case, m.priors_cxcycz.shape[0])
    with torch.no_grad():
        wb, wl, ws = m.detect_objects(locs, scores, min_score, max_overlap, top_k)
    gb, gl, gs = O.detect_objects(locs, scores, m.priors_cxcycz, min_score, max_overlap, top_k)
    for i in range(2):
        assert torch.equal(gl[i], wl[i])
        assert torch.equal(gs[i], ws[i])
        assert torch.equal(gb[i], wb[i])


@pytest.mark.parametrize("seed,threshold,n_obj", [(51, 0.5, [2, 0, 5]), (52, [0.1, 0.2], [3, 4]), (53, [0.3], [120, 2])])
def test_multibox_loss_live(ref, seed, threshold, n_obj):
    size = (64, 64, 64)
    m = _ref_model(ref[0], 1, size, threshold=threshold)
    case = dict(seed=seed, n_obj=n_obj)
    locs, scores, boxes, labels = GI.match_inputs(case, m.priors_cxcycz.shape[0])
    want_c, want_l = m.loss_fn(locs, scores, boxes, labels)
    got_c, got_l = O.multibox_loss(locs, scores, boxes, labels, m.priors_cxcycz, threshold)
    assert torch.equal(got_c, want_c) and torch.equal(got_l, want_l)


def test_box_functions_live(ref):
    utils = ref[2]
    g = torch.Generator().manual_seed(61)
    a, b = GI.random_gt_boxes(g, 37), GI.random_gt_boxes(g, 53)
    assert torch.equal(O.find_intersection3d(a, b), utils.find_intersection3d(a, b))
    assert torch.equal(O.find_jaccard_overlap3d(a, b), utils.find_jaccard_overlap3d(a, b))
    pri = utils.xyz_to_cxcycz(GI.random_gt_boxes(g, 37))
    assert torch.equal(O.xyz_to_cxcycz(a), utils.xyz_to_cxcycz(a))
    assert torch.equal(O.cxcycz_to_xyz(pri), utils.cxcycz_to_xyz(pri))
    enc = utils.cxcycz_to_gcxgcygcz(utils.xyz_to_cxcycz(a), pri)
    assert torch.equal(O.cxcycz_to_gcxgcygcz(O.xyz_to_cxcycz(a), pri), enc)
    assert torch.equal(O.gcxgcygcz_to_cxcycz(enc, pri), utils.gcxgcygcz_to_cxcycz(enc, pri))


def test_greedy_nms_grid_oracle_matches_reference_loop(ref):
    """The spatially hashed NMS oracle used for the 120 k .. 2.5 M GPU tests, against the reference's own
    suppress loop (ssd3d.py:407-426) on a list the n x n code can still hold."""
    utils = ref[2]
    g = torch.Generator().manual_seed(71)
    n = 3000
    ctr = torch.rand((n, 3), generator=g)
    side = 0.02 + 0.08 * torch.rand((n, 1), generator=g)
    boxes = torch.cat([ctr - side / 2, ctr + side / 2], 1)
    overlap = utils.find_jaccard_overlap3d(boxes, boxes)
    suppress = torch.zeros((n,), dtype=torch.uint8)
    for box in range(n):                                   # ssd3d.py:414-426, verbatim semantics
        if suppress[box] == 1:
            continue
        suppress = torch.max(suppress, (overlap[box] > 0.5).to(torch.uint8))
        suppress[box] = 0
    want = (1 - suppress).bool()
    assert torch.equal(O.greedy_nms_grid(boxes, 0.5), want)
    assert torch.equal(O.greedy_nms(boxes, 0.5), want)
