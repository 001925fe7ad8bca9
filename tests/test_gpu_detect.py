"""GPU parity of box geometry, decode/softmax, 3-D NMS and ``detect_objects``.

Integer outputs (NMS keep masks, kept prior indices, labels, counts) must be BIT-EXACT with the oracle on
identical inputs.  ``exp`` differs between the reference's CPU torch (Sleef) and CUDA ``expf`` by a few ulp,
so exactness is asserted stage-wise: the device's own softmax/decode output is fed to the oracle's
``detect_from_decoded`` (SURVEY.md section 7 "bit-exactness is promised on identical inputs to each stage").
Floating-point stages are compared within 8 fp32 ulp (rtol 1e-6).
"""
import pytest
import torch

from oracle import ssd3d_oracle as O
from tests.conftest import load_golden
from tests.golden import golden_inputs as GI

pytestmark = pytest.mark.gpu


def _ops():
    from mslesions3d_b200 import ops
    return ops


def test_box_functions_vs_reference_golden():
    from mslesions3d_b200 import utils as U
    g = load_golden("boxes.pt")
    a, b, gc = g["a"].cuda(), g["b"].cuda(), g["gc"].cuda()
    pri = U.xyz_to_cxcycz(b)
    # purely +,-,*,/ chains: bit-exact with the reference's torch ops
    assert torch.equal(U.find_intersection3d(a, b).cpu(), g["inter"])
    assert torch.equal(U.find_jaccard_overlap3d(a, b).cpu(), g["iou"])
    assert torch.equal(U.xyz_to_cxcycz(a).cpu(), g["cxcycz"])
    assert torch.equal(U.cxcycz_to_xyz(pri).cpu(), g["xyz"])
    dec = U.gcxgcygcz_to_cxcycz(gc, pri).cpu()
    assert torch.equal(dec[:, :3], g["decoded"][:, :3])                     # centres: no transcendental
    torch.testing.assert_close(dec[:, 3:], g["decoded"][:, 3:], rtol=1e-6, atol=0)   # exp()
    enc = U.cxcycz_to_gcxgcygcz(U.xyz_to_cxcycz(b.flip(0)), pri).cpu()
    assert torch.equal(enc[:, :3], g["encoded"][:, :3])
    torch.testing.assert_close(enc[:, 3:], g["encoded"][:, 3:], rtol=2e-6, atol=2e-6)   # log()


def test_iou_degenerate_boxes():
    from mslesions3d_b200 import utils as U
    a = torch.tensor([[0.1, 0.1, 0.1, 0.1, 0.3, 0.3],      # zero volume
                      [0.0, 0.0, 0.0, 1.0, 1.0, 1.0],
                      [0.5, 0.5, 0.5, 0.6, 0.6, 0.6]])
    got = U.find_jaccard_overlap3d(a.cuda(), a.cuda()).cpu()
    want = O.find_jaccard_overlap3d(a, a)
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(got[~torch.isnan(got)], want[~torch.isnan(want)])


@pytest.mark.parametrize("n_classes", [2, 3, 5])
def test_decode_softmax_stage(n_classes):
    ops = _ops()
    g = torch.Generator().manual_seed(n_classes)
    priors = O.prior_boxes((64, 64, 64))
    P = priors.shape[0]
    locs = torch.randn(3, P, 6, generator=g) * 0.7
    scores = torch.randn(3, P, n_classes, generator=g) * 2
    probs, boxes = ops.decode_softmax(locs.cuda(), scores.cuda(), priors.cuda())
    want_p = torch.softmax(scores, 2)
    want_b = torch.stack([O.cxcycz_to_xyz(O.gcxgcygcz_to_cxcycz(locs[i], priors)) for i in range(3)])
    torch.testing.assert_close(probs.cpu(), want_p, rtol=2e-6, atol=1e-9)
    torch.testing.assert_close(boxes.cpu(), want_b, rtol=2e-6, atol=2e-7)


def _random_boxes(n, g, side=(0.02, 0.12), dup=0):
    c = torch.rand(n, 3, generator=g)
    s = side[0] + (side[1] - side[0]) * torch.rand(n, 3, generator=g)
    b = torch.cat([c - s / 2, c + s / 2], 1)
    if dup:
        b[n - dup:] = b[:dup]          # exact duplicates: IoU == 1
    return b


@pytest.mark.parametrize("n,thr,dup", [(1, 0.5, 0), (63, 0.5, 0), (64, 0.3, 4), (65, 0.5, 0), (1000, 0.5, 50),
                                       (1000, 0.1, 0), (4097, 0.45, 7), (12000, 0.5, 100)])
def test_nms_keep_mask_bit_exact(n, thr, dup):
    ops = _ops()
    g = torch.Generator().manual_seed(n)
    side = (0.02, 0.12) if n < 5000 else (0.02, 0.06)
    boxes = _random_boxes(n, g, side, dup)
    keep = ops.nms3d_sorted(boxes.cuda(), thr).cpu()
    want = O.greedy_nms(boxes, ops.f32(thr))
    assert torch.equal(keep, want), "keep masks differ at %d positions" % int((keep != want).sum())
    assert bool(keep[0])


def _check_detect_stagewise(locs, scores, priors, min_score, max_overlap, top_k):
    ops = _ops()
    out = ops.detect_objects_padded(locs.cuda(), scores.cuda(), priors.cuda(), min_score, max_overlap, top_k)
    b, l, s, idx = ops.detect_lists(out, return_prior=True)
    probs, boxes = ops.decode_softmax(locs.cuda(), scores.cuda(), priors.cuda())
    wb, wl, ws, widx = O.detect_from_decoded(probs.cpu(), boxes.cpu(), ops.f32(min_score), ops.f32(max_overlap), top_k,
                                             return_indices=True)
    for i in range(locs.shape[0]):
        assert torch.equal(idx[i].cpu(), widx[i]), "image %d: kept prior indices differ" % i
        assert torch.equal(l[i].cpu(), wl[i])
        assert torch.equal(s[i].cpu(), ws[i])
        assert torch.equal(b[i].cpu(), wb[i])
    return b, l, s


@pytest.mark.parametrize("name", list(GI.DETECT_CASES))
def test_detect_objects_stagewise_exact_and_golden_close(name):
    case, gold = GI.DETECT_CASES[name], load_golden("detect.pt")[name]
    priors = O.prior_boxes(case["size"], in_channels=case["channels"])
    locs, scores = GI.detect_inputs(case, priors.shape[0])
    b, l, s = _check_detect_stagewise(locs, scores, priors, case["min_score"], case["max_overlap"], case["top_k"])
    # against the unmodified reference (its softmax/exp are CPU implementations): same detections,
    # values within a few ulp
    for i in range(case["batch"]):
        assert torch.equal(l[i].cpu(), gold["labels"][i]), "image %d: labels/count differ from the reference" % i
        torch.testing.assert_close(s[i].cpu(), gold["scores"][i], rtol=2e-6, atol=1e-9)
        torch.testing.assert_close(b[i].cpu(), gold["boxes"][i], rtol=2e-6, atol=2e-7)


def test_detect_ties_use_ascending_prior_index():
    priors = O.prior_boxes((64, 64, 64))
    P = priors.shape[0]
    g = torch.Generator().manual_seed(5)
    locs = torch.randn(2, P, 6, generator=g) * 0.5
    scores = torch.randn(2, P, 2, generator=g)
    scores[:, 100:400] = scores[:, 100:101]        # 300 exactly tied priors per image
    scores[1] = 0.25                               # everything tied in image 1
    _check_detect_stagewise(locs, scores, priors, 0.3, 0.5, 100)
    _check_detect_stagewise(locs, scores, priors, 0.0, 0.5, 20)


def test_detect_large_topk_all_candidates():
    # model_insight.py:146 setting: min_score=0, top_k=50000 -> every prior is a candidate, no truncation
    priors = O.prior_boxes((96, 96, 96))
    P = priors.shape[0]
    g = torch.Generator().manual_seed(9)
    locs = torch.randn(2, P, 6, generator=g) * 0.3
    scores = torch.randn(2, P, 3, generator=g)
    _check_detect_stagewise(locs, scores, priors, 0.0, 0.5, 50000)


def test_lssd3d_predict_step_and_detect_api():
    from mslesions3d_b200.ssd3d import LSSD3D
    case = GI.FORWARD_CASES["c1_64"]
    sd, x = GI.forward_inputs(case)
    m = LSSD3D(n_classes=2, input_channels=1, input_size=(64, 64, 64), min_score=0.3, top_k=40)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    with torch.no_grad():
        boxes, labels, scores = m.predict_step({"img": x.pin_memory()}, 0)      # host input, like the PL loader
        locs, cls = m(x.cuda())
        b2, l2, s2 = m.detect_objects(locs, cls, min_score=0.3, max_overlap=0.5, top_k=40)
    assert len(boxes) == 1 and boxes[0].is_cuda and labels[0].dtype == torch.int64
    assert boxes[0].shape[1] == 6 and boxes[0].shape[0] == labels[0].shape[0] == scores[0].shape[0] <= 40
    assert torch.equal(boxes[0], b2[0]) and torch.equal(labels[0], l2[0]) and torch.equal(scores[0], s2[0])
    wb, wl, ws = O.detect_from_decoded(*[t.cpu() for t in _ops().decode_softmax(locs, cls, m.priors_cxcycz)],
                                       _ops().f32(0.3), 0.5, 40)
    assert torch.equal(labels[0].cpu(), wl[0]) and torch.equal(scores[0].cpu(), ws[0])


def test_predict_step_graph_equals_eager_and_streams():
    from mslesions3d_b200.ssd3d import LSSD3D
    from mslesions3d_b200 import synthetic
    sd = O.random_state_dict(2, seed=5)
    m = LSSD3D(n_classes=2, input_channels=2, input_size=(48, 48, 48), min_score=0.35, top_k=30)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    batches = [torch.from_numpy(synthetic.make_batch(2, 2, (48, 48, 48), first_idx=10 * k)) for k in range(5)]
    with torch.no_grad():
        m.use_cuda_graph = False
        eager = [m.predict_step({"img": b.cuda()}, 0) for b in batches]
        m.use_cuda_graph = True
        graphed = [m.predict_step({"img": b.pin_memory()}, 0) for b in batches]      # host input, replayed graph
        streamed = list(m.predict_batches({"img": b.pin_memory()} for b in batches))  # pipelined copies
        resident = list(m.predict_batches({"img": b.cuda()} for b in batches))        # device-resident inputs
        hosted = list(m.predict_batches(({"img": b.pin_memory()} for b in batches), to_host=True))
    assert len(streamed) == len(batches) == len(hosted) == len(resident)
    assert not hosted[0][0][0].is_cuda and streamed[0][0][0].is_cuda
    for e, g, s, r, h in zip(eager, graphed, streamed, resident, hosted):
        for k in range(3):
            for i in range(2):
                assert torch.equal(e[k][i], g[k][i]) and torch.equal(e[k][i], s[k][i])
                assert torch.equal(e[k][i], r[k][i]) and torch.equal(e[k][i].cpu(), h[k][i])
    # a new state_dict invalidates the captured plan
    sd2 = O.random_state_dict(2, seed=6)
    m.load_state_dict(sd2)
    with torch.no_grad():
        again = m.predict_step({"img": batches[0].cuda()}, 0)
        m.use_cuda_graph = False
        again_eager = m.predict_step({"img": batches[0].cuda()}, 0)
    assert all(torch.equal(a, b) for a, b in zip(again[2], again_eager[2]))
    assert not all(torch.equal(a, b) if a.shape == b.shape else False for a, b in zip(again[2], eager[0][2]))


@pytest.mark.parametrize("P,n_classes,min_score,top_k", [(50000, 2, 0.5, 100), (50000, 3, 0.0, 300),
                                                         (16385, 2, 0.2, 100), (300000, 2, 0.3, 100)])
def test_detect_hierarchical_topk_many_candidates(P, n_classes, min_score, top_k):
    """More candidates than one sort block holds (the whole-brain / layer-0 prior sets): the hierarchical
    top-(10*top_k) reduction must give exactly the head of the full sort."""
    g = torch.Generator().manual_seed(P + top_k)
    c = torch.rand(P, 3, generator=g)
    s = 0.02 + 0.05 * torch.rand(P, 1, generator=g)
    priors = torch.cat([c, s.expand(P, 3)], 1)
    locs = torch.randn(2, P, 6, generator=g) * 0.3
    scores = torch.randn(2, P, n_classes, generator=g) * 2
    _check_detect_stagewise(locs, scores, priors, min_score, 0.5, top_k)


def test_whole_brain_layer0_config_end_to_end():
    """BASELINE config 4 shape: 2ch 160x192x160, prediction layers 0/3/5/7 -> 2 501 400 priors, default
    min_score / top_k.  Network vs the bf16-emulating oracle; detections exact on the device's decode."""
    from mslesions3d_b200.ssd3d import LSSD3D
    from mslesions3d_b200 import synthetic
    ar = {0: [1.], 3: [1.], 5: [1.], 7: [1.]}
    size = (160, 192, 160)
    sd = O.random_state_dict(2, ar, seed=11)
    m = LSSD3D(n_classes=2, input_channels=2, input_size=size, aspect_ratios=ar)
    assert m.priors_cxcycz.shape[0] == 2501400
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = torch.from_numpy(synthetic.make_batch(1, 2, size))
    with torch.no_grad():
        locs, scores = m(x.cuda())
        torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
        el, es = O.forward(sd, x, ar, emulate_bf16=True)
    dl, ds = (locs.cpu() - el).abs(), (scores.cpu() - es).abs()
    assert float(dl.max()) <= 0.06 and float(ds.max()) <= 0.06, (float(dl.max()), float(ds.max()))
    assert float(dl.mean()) <= 6e-3 and float(ds.mean()) <= 6e-3
    with torch.no_grad():
        boxes, labels, sc, idx = m.detect_objects(locs, scores, 0.5, 0.5, 100, return_prior=True)
    probs, dec = _ops().decode_softmax(locs, scores, m.priors_cxcycz)
    wb, wl, ws, widx = O.detect_from_decoded(probs.cpu(), dec.cpu(), 0.5, 0.5, 100, return_indices=True)
    assert torch.equal(idx[0].cpu(), widx[0]) and torch.equal(sc[0].cpu(), ws[0]) and torch.equal(boxes[0].cpu(), wb[0])


def test_predict_entry_point_writes_reference_style_outputs(tmp_path):
    """predict.py:235-281 -- checkpoint in, per-subject csv/json/volume + per-subject mAP json out."""
    import json
    import os
    from mslesions3d_b200 import predict
    from mslesions3d_b200.ssd3d import LSSD3D
    sd = O.random_state_dict(1, seed=9)
    ckpt = tmp_path / "model.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": dict(n_classes=2, input_channels=1, input_size=(64, 64, 64))}, ckpt)
    res = predict.predict_example(str(ckpt), str(tmp_path / "out"), dataset_path="", dataset_name="synthetic", n_classes=1,
                                  min_score=0.3, top_k=20, n_subjects=5, batch_size=2)
    out_dir = tmp_path / "out" / "synthetic" / "train_set" / "min_score_0.3"
    assert len(res) == 5
    for i in range(5):
        for ext in ("csv", "json", "nii.gz"):
            assert os.path.isfile(out_dir / ("sub-%d_preds.%s" % (i, ext)))
        info = json.load(open(out_dir / ("sub-%d_preds.json" % i)))
        n_det = len(res[str(i)]["boxes"])
        assert 1 <= n_det <= 20 and len(info) <= n_det
        for frac, box, label, score in info.values():
            assert label == 1 and score >= 0.3 and len(frac) == 6 and all(0 <= v <= 64 for v in box)
    for iou in (0.5, 0.1):
        m = json.load(open(out_dir / ("aa_metrics_per_subject_(min_IoU=%s).json" % iou)))
        assert set(m) == {str(i) for i in range(5)} and all("mAP" in v and "n_true_boxes" in v for v in m.values())
    # the same detections as a direct predict_step on the same volumes
    model = LSSD3D.load_from_checkpoint(str(ckpt), min_score=0.3).cuda().eval()
    model.top_k = 20
    from mslesions3d_b200 import synthetic
    vols = synthetic.make_batch(5, 1, (64, 64, 64))
    with torch.no_grad():
        b, l, s = model.predict_step({"img": torch.from_numpy(vols[:2])}, 0)
    assert torch.equal(torch.tensor(res["0"]["scores"]), s[0].cpu()) and torch.equal(torch.tensor(res["1"]["boxes"]), b[1].cpu())


def test_predict_entry_point_reads_generator_dataset_directory(tmp_path):
    """The generator's on-disk layout (images/*.nii.gz + labels/*.nii.gz) in, NIfTI outline volumes and per-subject
    mAP out; same detections as the in-memory synthetic path (same seeds -> same volumes)."""
    import json
    from mslesions3d_b200 import nifti, predict, synthetic
    sd = O.random_state_dict(1, seed=9)
    ckpt = tmp_path / "model.ckpt"
    torch.save({"state_dict": sd, "hyper_parameters": dict(n_classes=2, input_channels=1, input_size=(64, 64, 64))}, ckpt)
    data_dir = synthetic.write_dataset(str(tmp_path / "data"), 3, (64, 64, 64))
    res = predict.predict_example(str(ckpt), str(tmp_path / "out"), dataset_path=data_dir, dataset_name=None,
                                  min_score=0.3, top_k=20, batch_size=2)
    mem = predict.predict_example(str(ckpt), str(tmp_path / "out_mem"), dataset_path="", dataset_name=None,
                                  min_score=0.3, top_k=20, n_subjects=3, batch_size=2)
    assert list(res) == ["0000", "0001", "0002"]
    for i, sid in enumerate(res):
        assert res[sid] == mem[str(i)]
    out_dir = tmp_path / "out" / "train_set" / "min_score_0.3"
    seg, affine = nifti.load_nifti(str(out_dir / "sub-0000_preds.nii.gz"))
    info = json.load(open(out_dir / "sub-0000_preds.json"))
    assert seg.shape == (64, 64, 64) and (affine == torch.eye(4).numpy()).all()
    assert set(int(v) for v in set(seg.flatten().tolist())) == {0} | {int(k) for k in info}
    m = json.load(open(out_dir / "aa_metrics_per_subject_(min_IoU=0.5).json"))
    assert set(m) == {"0000", "0001", "0002"}


@pytest.mark.parametrize("shape", [(2, 1, 32, 32, 32), (3, 2, 17, 19, 23), (1, 1, 64, 64, 64)])
def test_normalize_intensity_nonzero_matches_data_module(shape):
    """datasets.py:403 (MONAI NormalizeIntensity(nonzero=True)) restated in synthetic.normalize_nonzero (numpy,
    fp64 statistics): the device version must agree to fp32 rounding, keep zeros at zero, and emit the stem's bf16."""
    import numpy as np
    from mslesions3d_b200 import ops, synthetic
    rs = np.random.RandomState(sum(shape))
    x = rs.rand(*shape).astype(np.float32)
    x[x < 0.3] = 0.0                                   # plenty of exact zeros (background)
    x[0, 0, :2] = 0.0
    want = np.stack([np.stack([synthetic.normalize_nonzero(x[n, c]) for c in range(shape[1])]) for n in range(shape[0])])
    got = ops.normalize_intensity_nonzero(torch.from_numpy(x).cuda(), torch.float32).cpu().numpy()
    assert np.array_equal(got == 0, x == 0) or np.all(got[x == 0] == 0)
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-6)
    got16 = ops.normalize_intensity_nonzero(torch.from_numpy(x).cuda()).float().cpu()
    assert torch.equal(got16, torch.from_numpy(got).to(torch.bfloat16).float())
    z = torch.zeros(1, 1, 8, 8, 8, device="cuda")
    assert float(ops.normalize_intensity_nonzero(z, torch.float32).abs().max()) == 0.0
    c = torch.full((1, 1, 8, 8, 8), 3.0, device="cuda")           # std == 0 -> divide by 1
    assert float(ops.normalize_intensity_nonzero(c, torch.float32).abs().max()) == 0.0


def test_predict_batches_ragged_last_batch_and_early_stop():
    """A loader whose last batch is smaller gets its own plan; abandoning the generator early leaves the model usable."""
    from mslesions3d_b200 import synthetic
    from mslesions3d_b200.ssd3d import LSSD3D
    sd = O.random_state_dict(1, seed=3)
    m = LSSD3D(n_classes=2, input_channels=1, input_size=(64, 64, 64), min_score=0.4, top_k=30)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    vols = torch.from_numpy(synthetic.make_batch(7, 1, (64, 64, 64)))
    batches = [vols[0:3], vols[3:6], vols[6:7]]
    with torch.no_grad():
        ref = [m.predict_step({"img": b.cuda()}, 0) for b in batches]
        got = list(m.predict_batches({"img": b.pin_memory()} for b in batches))
        assert [len(g[0]) for g in got] == [3, 3, 1]
        for r, g in zip(ref, got):
            for i in range(len(r[0])):
                assert torch.equal(r[0][i], g[0][i]) and torch.equal(r[1][i], g[1][i]) and torch.equal(r[2][i], g[2][i])
        gen = m.predict_batches({"img": b.cuda()} for b in batches * 3)
        first = next(gen)
        gen.close()                                   # consumer walks away with batches still in flight
        again = list(m.predict_batches({"img": b.cuda()} for b in batches))
        for r, g in zip(ref, again):
            assert torch.equal(r[0][0], g[0][0]) and torch.equal(r[2][0], g[2][0])
        assert torch.equal(first[0][0], ref[0][0][0])
