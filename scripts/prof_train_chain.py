#!/usr/bin/env python
"""Where does the captured training step spend its time?  Captures truncated variants of the C3 step
(1ch 96^3, batch 16) into their own CUDA graphs and times the replays with CUDA events:

    forward                      : train-mode forward only
    forward+loss                 : + matching + MultiBox loss/gradient
    no_leaves                    : + backward with every weight-gradient launch replaced by a no-op
    no_side_leaves               : + the stem weight gradient (main stream) kept
    only_{head,pw,dw}_wgrad      : no_side_leaves + that one family of side-stream weight gradients
    full                         : the real step (weight gradients on the side stream) + optimizer

The differences give the length of the forward chain, the loss, the backward data-gradient chain and what the
weight gradients add on top (contention with / tail after the critical chain).  Profiling only: the truncated
variants compute no usable gradients.  Prints one JSON object.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    from mslesions3d_b200 import _lib, ops, synthetic, training
    from mslesions3d_b200.ssd3d import LSSD3D

    _lib.load()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    size = (96,) * 3
    batch = 16
    x, b, l = synthetic.make_batch(batch, 1, size, first_idx=0, with_boxes=True)
    data = {"img": torch.from_numpy(x).to(dev), "boxes": [torch.from_numpy(v).to(dev) for v in b],
            "labels": [torch.from_numpy(v).to(dev) for v in l]}

    real = {k: getattr(ops, k) for k in ("pwconv_wgrad", "dwconv3d_wgrad", "head_wgrad", "stem_wgrad")}
    orig_run = training._TrainPlan._run

    def make_run(stop):
        def _run(self, model, with_optimizer):
            eng = model.train_engine()
            lf = model.loss_fn
            t0, t1 = (lf.threshold, lf.threshold) if lf.thresholding_mode == "hard" else lf.threshold
            locs, scores = eng.forward(self.image, eng.packed)
            self.loss = locs[0, 0, :2]
            if stop == "forward":
                eng.tape = None
                return
            m = ops.match_priors_packed(self.gt_boxes, self.gt_labels, self.offsets, self.n, self.tmax,
                                        model._prior_source(model.device), t0, t1)
            out, n_pos, g_locs, g_scores = ops.multibox_loss(locs, scores, m["true_classes"], m["true_locs"],
                                                             alpha=float(lf.alpha),
                                                             hard_negative_mining=lf.hard_negative_mining,
                                                             neg_pos_ratio=lf.neg_pos_ratio, want_grads=True)
            self.loss = out
            if stop == "forward+loss":
                eng.tape = None
                return
            grads = training._Grads(((n, p) for n, p in model.named_parameters()
                                     if p.requires_grad and n != "rescale_factors"), self.flat)
            eng.backward(g_locs, g_scores, grads)
        return _run

    res = {}
    only = {"only_head_wgrad": "head_wgrad", "only_pw_wgrad": "pwconv_wgrad", "only_dw_wgrad": "dwconv3d_wgrad"}
    names = ("forward", "forward+loss", "no_leaves", "no_side_leaves") + tuple(only) + ("full",)
    if len(sys.argv) > 1:
        names = tuple(sys.argv[1].split(","))
    for name in names:
        for k, v in real.items():
            setattr(ops, k, v)
        if name in ("no_leaves", "no_side_leaves") or name in only:
            for k in real:
                if name != "no_leaves" and k == "stem_wgrad":
                    continue
                if only.get(name) == k:
                    continue
                setattr(ops, k, lambda *a, **kw: None)
        training._TrainPlan._run = orig_run if name == "full" else make_run(name)
        sd = synthetic.random_state_dict(1, seed=0)
        model = LSSD3D(n_classes=2, input_channels=1, input_size=size, threshold=[0.1, 0.2], lr=1e-4)
        model.load_state_dict(sd)
        model = model.to(dev).train()
        model.use_cuda_graph = True
        for _ in range(3):
            model.fit_step(data)
        torch.cuda.synchronize()
        plan = next(iter(model.train_engine().plans.values()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = None
        for rep in range(3):
            e0.record()
            for _ in range(30):
                plan.graph.replay()
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 30
            best = t if best is None or t < best else best
        res[name] = {"ms": round(best, 4), "launches": plan.n_kernels}
        del model, plan
    for k, v in real.items():
        setattr(ops, k, v)
    training._TrainPlan._run = orig_run
    print(json.dumps(res))


if __name__ == "__main__":
    main()
