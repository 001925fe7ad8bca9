"""Prior boxes as a closed-form function of the prior index on the device (SURVEY.md 8f rank 3, ssd3d.py:286-342):
the materialised table must be BIT-identical to the reference's Python-double -> FloatTensor construction (golden
``priors.pt`` from the unmodified reference, and ``create_prior_boxes`` up to the 2 501 400-prior whole-brain
config), and decode / detect / matching fed from the table must return exactly what they return from the tensor."""
import pytest
import torch

from oracle import ssd3d_oracle as O
from tests.conftest import load_golden
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _model(case, **kw):
    from mslesions3d_b200.ssd3d import LSSD3D
    return LSSD3D(n_classes=kw.pop("n_classes", 2), input_channels=case["channels"], input_size=tuple(case["size"]),
                  aspect_ratios=case.get("aspect_ratios", {}), **kw)


@pytest.mark.parametrize("name", list(GI.PRIOR_CASES))
def test_device_priors_match_reference_golden(name):
    from mslesions3d_b200 import ops
    case, gold = GI.PRIOR_CASES[name], load_golden("priors.pt")[name]
    model = _model(case)
    tbl = model._prior_source(torch.device("cuda"))
    assert isinstance(tbl, ops.PriorTable) and tbl.count == gold["n"]
    p = tbl.materialize().cpu()
    assert torch.equal(p[:4], gold["first"]) and torch.equal(p[-4:], gold["last"])
    assert float(p.double().sum()) == gold["sum64"] and torch.equal(p.double().sum(0), gold["colsum64"])
    if gold["full"] is not None:
        assert torch.equal(p, gold["full"])
    assert torch.equal(p, model.priors_cxcycz.cpu())


@pytest.mark.parametrize("size,ar,scales", [
    ((160, 192, 160), {0: [1.], 3: [1.], 5: [1.], 7: [1.]}, {}),            # C4: 2 501 400 priors
    ((160, 192, 160), {}, {}),                                              # C4 realistic: 43 800
    ((72, 56, 88), {3: [1., 2.], 5: [1.], 7: [2., 1.]}, {}),                # several ratios per layer
    ((64, 64, 64), {}, {3: 0.07, 5: 0.4, 7: 0.8}),                          # user scales; 0.8 + 0.8 clamps to 1
])
def test_device_priors_equal_create_prior_boxes(size, ar, scales):
    from mslesions3d_b200 import ops
    model = _model(dict(channels=2, size=size, aspect_ratios=ar), scales=scales)
    tbl = model._prior_source(torch.device("cuda"))
    assert isinstance(tbl, ops.PriorTable)
    got = tbl.materialize()
    want = model.priors_cxcycz
    assert got.shape == want.shape and torch.equal(got, want.to(got.device))
    assert torch.equal(want.cpu(), O.prior_boxes_fast(size, ar or None, scales or None, in_channels=2))


def test_decode_detect_and_match_from_the_table_are_bit_identical():
    from mslesions3d_b200 import ops
    case = dict(channels=1, size=(64, 64, 64))
    model = _model(case)
    dev = torch.device("cuda")
    tbl, pri = model._prior_source(dev), model.priors_cxcycz.to(dev)
    assert isinstance(tbl, ops.PriorTable)
    locs, scores = GI.detect_inputs(dict(batch=3, seed=77), pri.shape[0])
    locs, scores = locs.cuda(), scores.cuda()
    a, b = ops.decode_softmax(locs, scores, pri), ops.decode_softmax(locs, scores, tbl)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    fa, fb = ops.decode_filter(locs, scores, pri, 0.4), ops.decode_filter(locs, scores, tbl, 0.4)
    assert torch.equal(fa[0], fb[0]) and torch.equal(fa[2], fb[2])
    da = ops.detect_lists(ops.detect_objects_padded(locs, scores, pri, 0.4, 0.5, 60), return_prior=True)
    db = ops.detect_lists(ops.detect_objects_padded(locs, scores, tbl, 0.4, 0.5, 60), return_prior=True)
    for la, lb in zip(da, db):
        for x, y in zip(la, lb):
            assert torch.equal(x, y)
    _, _, boxes, labels = GI.match_inputs(dict(seed=78, n_obj=[3, 0, 120]), pri.shape[0])
    boxes, labels = [t.cuda() for t in boxes], [t.cuda() for t in labels]
    ma, mb = ops.match_priors(boxes, labels, pri, 0.1, 0.2), ops.match_priors(boxes, labels, tbl, 0.1, 0.2)
    for k in ma:
        assert torch.equal(ma[k], mb[k]), k


def test_edited_priors_switch_back_to_the_tensor():
    """``priors_cxcycz`` is a public attribute of the reference's class: once a caller edits or replaces it, the
    kernels must read the tensor again."""
    from mslesions3d_b200 import ops
    model = _model(dict(channels=1, size=(64, 64, 64)))
    dev = torch.device("cuda")
    assert isinstance(model._prior_source(dev), ops.PriorTable)
    model.priors_cxcycz[0, 3:] *= 0.5               # edited before the first move to the device
    assert isinstance(model._prior_source(dev), torch.Tensor)
    model3 = _model(dict(channels=1, size=(64, 64, 64)))
    assert isinstance(model3._prior_source(dev), ops.PriorTable)
    model3.priors_cxcycz[0, 3:] *= 0.5              # edited in place on the device
    assert isinstance(model3._prior_source(dev), torch.Tensor)
    model2 = _model(dict(channels=1, size=(64, 64, 64)))
    model2.priors_cxcycz = model2.priors_cxcycz.clone()
    assert isinstance(model2._prior_source(dev), torch.Tensor)
