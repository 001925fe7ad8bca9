"""Generate the committed golden vectors by running the UNMODIFIED reference.

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

It imports the reference through ``oracle/ref_shim.py`` (stub modules for the
absent pytorch_lightning/monai/matplotlib, nothing else changed), feeds it
seeded inputs that the tests can regenerate bit-for-bit (``oracle.random_state_dict``,
``mslesions3d_b200.synthetic``, ``golden_inputs`` below) and stores ONLY the
reference's outputs (plus input checksums to catch RNG drift) in
``tests/golden/*.pt``.  The GPU box has no /root/reference; tests there replay
these files.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ssd3d_oracle as O  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402
from tests.golden.golden_inputs import (FORWARD_CASES, DETECT_CASES, MATCH_CASES, PRIOR_CASES, MAP_CASES,  # noqa: E402
                                        forward_inputs, detect_inputs, match_inputs, map_inputs, checksum)


def build_reference_model(ssd3d, case, **over):
    kw = dict(n_classes=case.get("n_classes", 2), input_channels=case["channels"],
              input_size=tuple(case["size"]), aspect_ratios=case.get("aspect_ratios", {}))
    kw.update(over)
    torch.manual_seed(0)
    return ssd3d.LSSD3D(**kw).eval()


def gtbox_goldens(utils):
    """BoundingBoxesGeneratord.converter of the reference on seeded segmentations (utils.py:438-482)."""
    from tests.golden.golden_inputs import GTBOX_CASES, gtbox_inputs
    out = {}
    for name, case in GTBOX_CASES.items():
        segs = gtbox_inputs(case)
        kw = dict(segmentation_mode=case["mode"])
        if case["mode"] == "classes":
            kw["n_classes"] = case["n_classes"]
        gen = utils.BoundingBoxesGeneratord(keys=["seg"], **kw)
        res = []
        for v in range(segs.shape[0]):
            try:
                res.append(gen.converter(segs[v][None].copy()))
            except RuntimeError as e:      # a volume without objects: FloatTensor([]) / FloatTensor(6) raises
                assert segs[v].sum() == 0, e
                res.append((None, None))
        out[name] = dict(boxes=[None if b is None else b.clone() for b, _ in res],
                         labels=[None if l is None else l.clone() for _, l in res],
                         in_sum=float(segs.astype(np.float64).sum()))
        print("gtbox", name, [None if b is None else int(b.shape[0]) for b, _ in res])
    torch.save(out, os.path.join(HERE, "gtbox.pt"))


def gtbox_instance_goldens(utils):
    """BoundingBoxesGeneratord.converter in "instances" mode (utils.py:439-441,483-513) -> gtbox_instances.pt."""
    from tests.golden.golden_inputs import GTBOX_INSTANCE_CASES, gtbox_instance_inputs
    out = {}
    for name, case in GTBOX_INSTANCE_CASES.items():
        segs = gtbox_instance_inputs(case)
        gen = utils.BoundingBoxesGeneratord(keys=["seg"], segmentation_mode="instances", thresholds=case["thresholds"])
        res = []
        for v in range(segs.shape[0]):
            try:
                res.append(gen.converter(segs[v][None].copy()))
            except RuntimeError as e:      # a volume without objects: FloatTensor([]) / FloatTensor(6) raises
                assert segs[v].sum() == 0, e
                res.append((None, None))
        out[name] = dict(boxes=[None if b is None else b.clone() for b, _ in res],
                         labels=[None if l is None else l.clone() for _, l in res],
                         in_sum=float(segs.astype(np.float64).sum()))
        print("gtbox instances", name, [None if b is None else (int(b.shape[0]), l.tolist()) for b, l in res])
    torch.save(out, os.path.join(HERE, "gtbox_instances.pt"))


def detect_long_goldens(ssd3d):
    """LSSD3D.detect_objects of the reference on candidate lists beyond the fused kernel's limits -> detect_long.pt."""
    from tests.golden.golden_inputs import DETECT_LONG_CASES, detect_long_inputs
    out = {}
    for name, case in DETECT_LONG_CASES.items():
        m = build_reference_model(ssd3d, case)
        locs, scores = detect_long_inputs(case, m.priors_cxcycz.shape[0])
        with torch.no_grad():
            b, l, s = m.detect_objects(locs, scores, case["min_score"], case["max_overlap"], case["top_k"])
        out[name] = dict(boxes=[t.clone() for t in b], labels=[t.clone() for t in l], scores=[t.clone() for t in s],
                         n_priors=int(m.priors_cxcycz.shape[0]),
                         in_sum=checksum(torch.cat([locs.flatten(), scores.flatten()])))
        print("detect_long", name, m.priors_cxcycz.shape[0], [t.shape[0] for t in b])
    torch.save(out, os.path.join(HERE, "detect_long.pt"))


def main():
    ssd3d, mobilenet, utils = load_reference()
    torch.set_num_threads(1)
    if len(sys.argv) > 1 and sys.argv[1] == "detect_long":
        torch.set_num_threads(8)
        detect_long_goldens(ssd3d)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gtbox":      # only this section (the others are unchanged)
        gtbox_goldens(utils)
        return
    if len(sys.argv) > 1 and sys.argv[1] == "gtbox_instances":
        gtbox_instance_goldens(utils)
        return
    gtbox_instance_goldens(utils)
    gtbox_goldens(utils)

    # ---- priors --------------------------------------------------------------
    out = {}
    for name, case in PRIOR_CASES.items():
        m = build_reference_model(ssd3d, case)
        p = m.priors_cxcycz
        out[name] = dict(n=p.shape[0], first=p[:4].clone(), last=p[-4:].clone(),
                         sum64=float(p.double().sum()), colsum64=p.double().sum(0),
                         full=p.clone() if p.shape[0] <= 5000 else None)
        print("priors", name, p.shape)
    torch.save(out, os.path.join(HERE, "priors.pt"))

    # ---- forward -------------------------------------------------------------
    out = {}
    for name, case in FORWARD_CASES.items():
        sd, x = forward_inputs(case)
        m = build_reference_model(ssd3d, case)
        missing = m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            locs, scores = m(x)
        out[name] = dict(locs=locs.clone(), scores=scores.clone(), x_sum=checksum(x),
                         w_sum=checksum(torch.cat([v.flatten().float() for v in sd.values()])),
                         keys=list(m.state_dict().keys()))
        print("forward", name, tuple(locs.shape), float(locs.abs().mean()), float(scores.abs().mean()))
    torch.save(out, os.path.join(HERE, "forward.pt"))

    # ---- detect_objects --------------------------------------------------------
    out = {}
    for name, case in DETECT_CASES.items():
        m = build_reference_model(ssd3d, case)
        locs, scores = detect_inputs(case, m.priors_cxcycz.shape[0])
        with torch.no_grad():
            b, l, s = m.detect_objects(locs, scores, case["min_score"], case["max_overlap"], case["top_k"])
        out[name] = dict(boxes=[t.clone() for t in b], labels=[t.clone() for t in l],
                         scores=[t.clone() for t in s], in_sum=checksum(torch.cat([locs.flatten(), scores.flatten()])))
        print("detect", name, [t.shape[0] for t in b])
    torch.save(out, os.path.join(HERE, "detect.pt"))

    # ---- MultiBoxLoss ----------------------------------------------------------
    out = {}
    for name, case in MATCH_CASES.items():
        m = build_reference_model(ssd3d, case, threshold=case["threshold"])
        P = m.priors_cxcycz.shape[0]
        locs, scores, boxes, labels = match_inputs(case, P)
        rec = {}
        real_ce, real_l1 = m.loss_fn.cross_entropy, m.loss_fn.smooth_l1

        class Spy(torch.nn.Module):           # observe only; the reference source is untouched
            def __init__(self, real, hook):
                super().__init__()
                self.real, self.hook = real, hook

            def forward(self, a, b):
                self.hook(a, b)
                return self.real(a, b)

        def ce_hook(inp, tgt):
            rec["tc"] = tgt.clone()

        def l1_hook(a, b):
            rec["true_locs_pos"] = b.clone()
            rec["n_pos"] = a.shape[0]

        m.loss_fn.cross_entropy = Spy(real_ce, ce_hook)
        m.loss_fn.smooth_l1 = Spy(real_l1, l1_hook)
        conf, loc = m.loss_fn(locs, scores, boxes, labels)
        out[name] = dict(conf_loss=conf.clone(), loc_loss=loc.clone(), tc=rec["tc"].view(len(boxes), P),
                         true_locs_pos=rec["true_locs_pos"], n_pos=rec["n_pos"])
        print("match", name, float(conf), float(loc), rec["n_pos"])
    torch.save(out, os.path.join(HERE, "match.pt"))

    # ---- calculate_mAP ------------------------------------------------------------------
    out = {}
    for name, case in MAP_CASES.items():
        db, dl, ds, tb, tl, td = map_inputs(case)
        r = utils.calculate_mAP(db, dl, ds, tb, tl, td, min_overlap=case["min_overlap"], return_detail=True)
        simple = utils.calculate_mAP(db, dl, ds, tb, tl, td, min_overlap=case["min_overlap"])
        out[name] = dict(detail={k: (v.clone() if torch.is_tensor(v) else v) for k, v in r.items()}, simple=simple,
                         in_sum=checksum(torch.cat([torch.cat(db).flatten(), torch.cat(ds)])))
        print("mAP", name, r["mAP"], r["n_true_boxes"])
    torch.save(out, os.path.join(HERE, "map.pt"))

    # ---- training step (train-mode forward, MultiBox loss, backward) ---------------------------
    from tests.golden.golden_inputs import TRAIN_CASES, train_inputs
    out = {}
    for name, case in TRAIN_CASES.items():
        sd, x, boxes, labels = train_inputs(case)
        m = build_reference_model(ssd3d, case, threshold=case["threshold"])
        m.load_state_dict(sd, strict=True)
        m.train()
        locs, scores = m(x)
        conf, loc = m.loss_fn(locs, scores, boxes, labels)
        (conf + m.loss_fn.alpha * loc).backward()
        grads = {}
        for k, p in m.named_parameters():
            if p.grad is None:
                grads[k] = None
            else:
                gflat = p.grad.flatten()
                grads[k] = dict(norm=float(gflat.double().norm()), sum=float(gflat.double().sum()),
                                head=gflat[:8].clone(), tail=gflat[-8:].clone())
        buffers = {k: v.clone() for k, v in m.named_buffers() if v.numel() <= 64 or "features.0." in k}
        out[name] = dict(conf=conf.detach().clone(), loc=loc.detach().clone(), grads=grads, buffers=buffers,
                         locs_sum=checksum(locs.detach()), x_sum=checksum(x))
        print("train", name, float(conf), float(loc))
    torch.save(out, os.path.join(HERE, "train.pt"))

    # ---- box utility functions ----------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    a = torch.rand(64, 3, generator=g) * 0.8
    a = torch.cat([a, a + 0.02 + 0.3 * torch.rand(64, 3, generator=g)], 1)
    b = torch.rand(96, 3, generator=g) * 0.8
    b = torch.cat([b, b + 0.02 + 0.3 * torch.rand(96, 3, generator=g)], 1)
    pri = utils.xyz_to_cxcycz(b)
    gc = torch.randn(96, 6, generator=g)
    out = dict(a=a, b=b, gc=gc,
               iou=utils.find_jaccard_overlap3d(a, b), inter=utils.find_intersection3d(a, b),
               cxcycz=utils.xyz_to_cxcycz(a), xyz=utils.cxcycz_to_xyz(pri),
               decoded=utils.gcxgcygcz_to_cxcycz(gc, pri),
               encoded=utils.cxcycz_to_gcxgcygcz(utils.xyz_to_cxcycz(b.flip(0)), pri))
    torch.save(out, os.path.join(HERE, "boxes.pt"))
    print("done")


if __name__ == "__main__":
    main()
