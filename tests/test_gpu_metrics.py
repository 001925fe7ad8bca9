"""GPU parity of the detection-metric kernels (csrc/metrics.cu, utils.calculate_mAP) against the CPU oracle and the
golden outputs of the unmodified reference: true/false positives, detected flags and sorted scores are exact
(integer / copy semantics on exact IoUs); AP, precision, recall, F1 within 1e-6 relative (one division / an
11-term mean whose summation order torch does not specify)."""
import pytest
import torch

from oracle import ssd3d_oracle as O
from tests.conftest import load_golden
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _close(a, b):
    a, b = float(a), float(b)
    return (a != a and b != b) or abs(a - b) <= 1e-6 * max(1.0, abs(b))


@pytest.mark.parametrize("name", list(GI.MAP_CASES))
def test_calculate_map_matches_reference_golden(name):
    from mslesions3d_b200 import utils
    case, gold = GI.MAP_CASES[name], load_golden("map.pt")[name]
    db, dl, ds, tb, tl, td = GI.map_inputs(case)
    cu = lambda xs: [t.cuda() for t in xs]
    r = utils.calculate_mAP(cu(db), cu(dl), cu(ds), cu(tb), cu(tl), cu(td), min_overlap=case["min_overlap"],
                            return_detail=True)
    d = gold["detail"]
    assert _close(r["mAP"], d["mAP"]) and r["n_true_boxes"] == d["n_true_boxes"]
    simple = utils.calculate_mAP(cu(db), cu(dl), cu(ds), cu(tb), cu(tl), cu(td), min_overlap=case["min_overlap"])
    assert _close(simple[1], gold["simple"][1]) and _close(simple[0], gold["simple"][0])
    assert torch.equal(r["TP"].cpu(), d["TP"]) and torch.equal(r["FP"].cpu(), d["FP"])
    for k in ("precision", "recall", "f1_score", "APs"):
        assert _close(r[k], d[k]), (k, r[k], d[k])
    if torch.is_tensor(d["found_boxes_volumes_per_class"]):
        assert torch.equal(r["found_boxes_volumes_per_class"].cpu(), d["found_boxes_volumes_per_class"])
        assert torch.equal(r["not_found_boxes_volumes_per_class"].cpu(), d["not_found_boxes_volumes_per_class"])
    if isinstance(d["sorted_det_scores"], dict) and 1 in d["sorted_det_scores"]:
        assert torch.equal(r["sorted_det_scores"][1].cpu(), d["sorted_det_scores"][1])


@pytest.mark.parametrize("n_img,n_det,n_obj,ties,difficult", [(6, 300, 12, False, False), (3, 2500, 40, True, False),
                                                              (8, 64, 0, False, False), (4, 500, 20, True, True)])
def test_map_class_kernel_vs_oracle(n_img, n_det, n_obj, ties, difficult):
    from mslesions3d_b200 import ops
    g = torch.Generator().manual_seed(n_det + n_obj)
    t_img = torch.randint(0, n_img, (n_obj,), generator=g)
    t_box = GI.random_gt_boxes(g, n_obj)
    t_dif = (torch.rand(n_obj, generator=g) < 0.3) if difficult else torch.zeros(n_obj, dtype=torch.bool)
    d_img = torch.randint(0, n_img, (n_det,), generator=g)
    base = t_box[torch.randint(0, max(n_obj, 1), (n_det,), generator=g)] if n_obj else GI.random_gt_boxes(g, n_det)
    d_box = (base + 0.04 * torch.randn(n_det, 6, generator=g)).float()
    d_box = torch.cat([torch.minimum(d_box[:, :3], d_box[:, 3:] - 0.01), d_box[:, 3:]], 1)
    d_sc = torch.rand(n_det, generator=g)
    if ties:
        d_sc = (d_sc * 20).round() / 20          # many exact ties: stable order by input index
    thr = torch.arange(start=0, end=1.1, step=.1)
    m = ops.map_class(d_box.cuda(), d_sc.cuda(), d_img.cuda(), t_box.cuda(), t_dif.cuda(), t_img.cuda(), 0.3, thr)
    tp, fp, det, scores, order = O.metrics_per_class(d_img, d_box, d_sc, t_img, t_box, t_dif, 0.3, stable=True)
    assert torch.equal(m["sort_index"].cpu().long(), order)
    assert torch.equal(m["sorted_scores"].cpu(), scores)
    assert torch.equal(m["tp"].cpu(), tp) and torch.equal(m["fp"].cpu(), fp)
    assert torch.equal(m["detected"].cpu(), det)
    n_easy = int(t_dif.logical_not().sum())
    ap, prec, rec, p11 = O.average_precision_11pt(tp, fp, n_easy)
    if n_easy:
        torch.testing.assert_close(m["cum_precision"].cpu(), prec, rtol=1e-6, atol=0)
        torch.testing.assert_close(m["cum_recall"].cpu(), rec, rtol=1e-6, atol=0)
    assert torch.equal(m["stats"][4:].cpu(), p11) and _close(m["stats"][0], ap)


def test_training_step_reports_metrics_like_the_reference():
    """training_step every 2*n epochs runs detect_objects + calculate_mAP at IoU 0.1 and 0.5 (ssd3d.py:497-515)."""
    from mslesions3d_b200 import synthetic
    from mslesions3d_b200.ssd3d import LSSD3D
    sd = O.random_state_dict(1, seed=5)
    model = LSSD3D(n_classes=2, input_channels=1, input_size=(64, 64, 64), threshold=[0.1, 0.2], min_score=0.3)
    model.load_state_dict(sd)
    model = model.cuda().train()
    x, b, l = synthetic.make_batch(2, 1, (64, 64, 64), with_boxes=True)
    batch = {"img": torch.from_numpy(x), "boxes": [torch.from_numpy(v) for v in b], "labels": [torch.from_numpy(v) for v in l]}
    out = model.training_step(batch)
    logs = out["log"]
    assert "metrics_10" in logs and "metrics_50" in logs
    for mtr in (logs["metrics_10"], logs["metrics_50"]):
        assert set(mtr) >= {"APs", "mAP", "precision", "recall", "f1_score", "TP", "FP", "n_true_boxes"}
        assert mtr["n_true_boxes"] == sum(int(v.shape[0]) for v in b)
    out["loss"].backward()
