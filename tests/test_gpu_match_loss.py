"""GPU parity of prior<->object matching and the MultiBox loss (ssd3d.py:741-941).

Matched indices and labels (object per prior, prior per object, true_classes incl. the -1 ignore band) are
BIT-EXACT with the oracle; IoU values and the centre part of the encoded targets too (pure +,-,*,/).  The
log() part of the targets and the loss scalars (different summation order, CUDA logf/expf) are compared
within rtol 2e-5; gradients against torch autograd of the oracle within 1e-6 absolute.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import ssd3d_oracle as O
from tests.conftest import load_golden
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _ops():
    from mslesions3d_b200 import ops
    return ops


def _thresholds(th):
    mode, t0, t1 = O.parse_threshold(th)
    return t0, t1


@pytest.mark.parametrize("name", list(GI.MATCH_CASES))
def test_matching_bit_exact(name):
    ops = _ops()
    case = GI.MATCH_CASES[name]
    priors = O.prior_boxes(case["size"], in_channels=case["channels"])
    _, _, boxes, labels = GI.match_inputs(case, priors.shape[0])
    t0, t1 = _thresholds(case["threshold"])
    m = ops.match_priors([b.cuda() for b in boxes], [l.cuda() for l in labels], priors.cuda(), t0, t1)
    off = 0
    for i, (bx, lb) in enumerate(zip(boxes, labels)):
        n_obj = bx.shape[0]
        if n_obj == 0:
            assert int(m["true_classes"][i].abs().sum()) == 0 and float(m["true_locs"][i].abs().sum()) == 0.0
            continue
        lab, tl, ov, obj, pfo = O.match_image(bx, lb, priors, case["threshold"])
        assert torch.equal(m["prior_for_object"][off:off + n_obj].cpu().long(), pfo), "image %d prior_for_object" % i
        assert torch.equal(m["object_for_prior"][i].cpu().long(), obj), "image %d object_for_prior" % i
        assert torch.equal(m["overlap"][i].cpu(), ov), "image %d overlaps" % i
        assert torch.equal(m["true_classes"][i].cpu(), lab), "image %d true_classes" % i
        got = m["true_locs"][i].cpu()
        assert torch.equal(got[:, :3], tl[:, :3])
        torch.testing.assert_close(got[:, 3:], tl[:, 3:], rtol=2e-6, atol=2e-6)
        off += n_obj


def test_matching_duplicate_best_prior_last_object_wins():
    ops = _ops()
    priors = O.prior_boxes((64, 64, 64))
    # three identical objects + one distinct: identical ones share the best prior; ties in argmax over objects
    b = torch.tensor([[0.2, 0.2, 0.2, 0.4, 0.4, 0.4]] * 3 + [[0.6, 0.6, 0.6, 0.9, 0.9, 0.9]])
    l = torch.tensor([1, 1, 1, 1])
    m = ops.match_priors([b.cuda()], [l.cuda()], priors.cuda(), 0.5, 0.5)
    lab, tl, ov, obj, pfo = O.match_image(b, l, priors, 0.5)
    assert torch.equal(m["prior_for_object"].cpu().long(), pfo)
    assert torch.equal(m["object_for_prior"][0].cpu().long(), obj)
    assert torch.equal(m["true_classes"][0].cpu(), lab)
    assert int(obj[pfo[0]]) == 2      # last writer wins on the shared best prior


@pytest.mark.parametrize("name", list(GI.MATCH_CASES))
def test_multibox_loss_vs_oracle_and_reference_golden(name):
    from mslesions3d_b200.ssd3d import MultiBoxLoss
    case, gold = GI.MATCH_CASES[name], load_golden("match.pt")[name]
    priors = O.prior_boxes(case["size"], in_channels=case["channels"])
    locs, scores, boxes, labels = GI.match_inputs(case, priors.shape[0])
    loss_fn = MultiBoxLoss(priors.cuda(), threshold=case["threshold"])
    gl = locs.cuda().requires_grad_(True)
    gs = scores.cuda().requires_grad_(True)
    conf, loc = loss_fn(gl, gs, [b.cuda() for b in boxes], [l.cuda() for l in labels])
    torch.testing.assert_close(conf.detach().cpu(), gold["conf_loss"], rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(loc.detach().cpu(), gold["loc_loss"], rtol=2e-5, atol=1e-6)
    (conf + 1.7 * loc).backward()
    # autograd through the oracle for the gradient check
    ol = locs.clone().requires_grad_(True)
    os_ = scores.clone().requires_grad_(True)
    oc, olc = O.multibox_loss(ol, os_, boxes, labels, priors, case["threshold"])
    (oc + 1.7 * olc).backward()
    torch.testing.assert_close(gl.grad.cpu(), ol.grad, rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(gs.grad.cpu(), os_.grad, rtol=1e-4, atol=1e-7)


def test_multibox_loss_hard_negative_mining_variant():
    ops = _ops()
    case = GI.MATCH_CASES["soft"]
    priors = O.prior_boxes(case["size"], in_channels=case["channels"])
    locs, scores, boxes, labels = GI.match_inputs(case, priors.shape[0])
    t0, t1 = _thresholds(case["threshold"])
    m = ops.match_priors([b.cuda() for b in boxes], [l.cuda() for l in labels], priors.cuda(), t0, t1)
    out, n_pos, g_l, g_s = ops.multibox_loss(locs.cuda(), scores.cuda(), m["true_classes"], m["true_locs"], 1.0,
                                             hard_negative_mining=True, neg_pos_ratio=3)
    oc, olc = O.multibox_loss(locs, scores, boxes, labels, priors, case["threshold"], hard_negative_mining=True)
    torch.testing.assert_close(out[0].cpu(), oc, rtol=2e-5, atol=1e-6)
    torch.testing.assert_close(out[1].cpu(), olc, rtol=2e-5, atol=1e-6)
    # the selected negatives carry gradient, the others none
    ol = locs.clone()
    os_ = scores.clone().requires_grad_(True)
    oc2, _ = O.multibox_loss(ol, os_, boxes, labels, priors, case["threshold"], hard_negative_mining=True)
    oc2.backward()
    torch.testing.assert_close(g_s.cpu(), os_.grad, rtol=1e-4, atol=1e-7)


def test_loss_without_positives_raises():
    from mslesions3d_b200.ssd3d import MultiBoxLoss
    priors = O.prior_boxes((64, 64, 64))
    P = priors.shape[0]
    loss_fn = MultiBoxLoss(priors.cuda(), threshold=0.5)
    locs, scores = torch.zeros(1, P, 6).cuda(), torch.zeros(1, P, 2).cuda()
    with pytest.raises(Exception, match="Loss is NaN"):
        loss_fn(locs, scores, [torch.zeros(0, 6).cuda()], [torch.zeros(0, dtype=torch.long).cuda()])
