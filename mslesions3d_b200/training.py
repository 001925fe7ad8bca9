"""Training step of the SSD3D network on the sm_100a kernels (``LSSD3D.training_step`` /
``configure_optimizers``, ssd3d.py:467-531,704-722).

The reference gets the backward pass from torch autograd over ``nn.Conv3d`` / ``nn.BatchNorm3d`` /
``nn.ReLU`` and steps ``torch.optim.Adam``.  Here the whole network is ONE autograd node: ``TrainEngine``
runs the train-mode forward (raw conv -> batch-statistic BN -> ReLU per unit, fused SSD heads) while
recording a tape of saved activations, and walks it backwards with hand-written kernels (BN/ReLU backward,
pointwise / depthwise / stem / head weight and data gradients).  Two ways in:

* ``LSSD3D.training_step(batch)`` returns a loss tensor whose ``backward()`` fills ``p.grad`` of every
  parameter -- the reference's contract, any torch optimizer can follow;
* ``LSSD3D.fit_step(batch)`` is the fused loop body: forward, MultiBox loss with its analytic gradient,
  backward straight into one flat fp32 gradient buffer, one NCCL all-reduce of that buffer when
  ``torch.distributed`` is initialised (data parallel over volumes), one fused Adam
  launch over the flat parameter buffer (weight decay 5e-4, biases at 2x lr, cosine schedule).

PyTorch is plumbing (memory, streams, autograd hand-off, NCCL); every arithmetic step is a library kernel.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib, ops
from .mobilenet import Block, ConvBN, _stride3


class _Grads:
    """Destination tensors of the parameter gradients, keyed by parameter name."""

    def __init__(self, named_params, flat: Optional["FlatParams"] = None):
        self.t: Dict[str, torch.Tensor] = {}
        for name, p in named_params:
            if not p.requires_grad:
                continue
            if flat is not None and name in flat.views_grad:
                self.t[name] = flat.views_grad[name]
            else:
                self.t[name] = torch.empty_like(p, dtype=torch.float32, memory_format=torch.contiguous_format)

    def __getitem__(self, name):
        return self.t[name]


class FlatParams:
    """All trainable parameters re-homed as views of ONE fp32 buffer, gradients and Adam moments likewise.
    Layout: [weights ...][biases ...] (``name.endswith('.bias')``, the reference's 2x-lr group,
    ssd3d.py:706-715), every tensor padded to a multiple of 8 elements so that kernels may use 16/32-byte
    accesses on any view.  Parameters that never receive a gradient (``rescale_factors``) stay
    outside, exactly as torch.optim.Adam skips ``grad is None``."""

    def __init__(self, model: nn.Module, skip=("rescale_factors",)):
        named = [(n, p) for n, p in model.named_parameters() if p.requires_grad and n not in skip]
        weights = [(n, p) for n, p in named if not n.endswith(".bias")]
        biases = [(n, p) for n, p in named if n.endswith(".bias")]
        dev = named[0][1].device
        off = 0
        self.offsets: Dict[str, int] = {}
        for n, p in weights:
            self.offsets[n] = off
            off += (p.numel() + 7) // 8 * 8
        self.bias_start = off
        for n, p in biases:
            self.offsets[n] = off
            off += (p.numel() + 7) // 8 * 8
        self.numel = off
        self.param = torch.zeros((off,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((off,), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros((off,), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros((off,), dtype=torch.float32, device=dev)
        self.views_grad: Dict[str, torch.Tensor] = {}
        self.step = 0
        with torch.no_grad():
            for n, p in weights + biases:
                o = self.offsets[n]
                view = self.param[o:o + p.numel()].view(p.shape)
                view.copy_(p.detach().float())
                p.data = view
                self.views_grad[n] = self.grad[o:o + p.numel()].view(p.shape)


class TrainEngine:
    """Train-mode forward with a tape + hand-written backward for ``LSSD3D`` (MobileNet base + SSD heads)."""

    def __init__(self, model):
        self.model = model
        self.flat: Optional[FlatParams] = None
        self.tape = None

    # ------------------------------------------------------------------------------------------
    def flatten(self) -> FlatParams:
        if self.flat is None or self.flat.param.device != self.model.device:
            self.flat = FlatParams(self.model)
            self.model.invalidate_packed()
        return self.flat

    # ------------------------------------------------------------------------------------------
    def forward(self, image: torch.Tensor):
        """image (N, Cin, D, H, W) -> locs (N,P,6), scores (N,P,n_classes); records the tape."""
        m = self.model
        dev = m.device
        if dev.type != "cuda":
            raise RuntimeError("training needs the model on a CUDA device; there is no CPU path")
        if m.n_classes * m.boxes_per_location + 6 * m.boxes_per_location > 16:
            raise NotImplementedError("the head backward kernels are built for bpl*(6+n_classes) <= 16 "
                                      "(the reference's binary lesion/background setting)")
        if image.device != dev:
            image = image.to(dev, non_blocking=True)
        if image.dtype not in (torch.float32, torch.bfloat16):
            image = image.float()
        image = image.contiguous()
        flag = m.base.nan_flag(dev)
        tape = {"image": image, "units": [], "heads": []}
        x = image
        feats = {}
        keys = list(m.aspect_ratios.keys())
        for i, feat in enumerate(m.base.features):
            if isinstance(feat, ConvBN):
                conv, bn = feat[0], feat[1]
                sd = _stride3(conv.stride)[0]
                w = ops.pack_stem_weight(conv.weight)
                z = ops.stem_conv_raw(x, w, sd)
                a, st = ops.bn_train_relu(z, bn, flag)
                tape["units"].append(dict(kind="stem", idx=i, x=x, z=z, st=st, stride=sd))
                x = a
            elif isinstance(feat, Block):
                s = _stride3(feat.conv1.stride)[0]
                wd = ops.pack_dw_weight(feat.conv1.weight)
                z1 = ops.dwconv3d_raw(x, wd, s)
                a1, st1 = ops.bn_train_relu(z1, feat.bn1, None)
                wp = ops.pack_pw_weight(feat.conv2.weight)
                z2 = ops.pwconv_raw(a1, wp)
                a2, st2 = ops.bn_train_relu(z2, feat.bn2, flag)
                tape["units"].append(dict(kind="block", idx=i, x=x, z1=z1, st1=st1, a1=a1, z2=z2, st2=st2, wd=wd, wp=wp,
                                          stride=s))
                x = a2
            else:
                raise NotImplementedError("unexpected backbone layer %r" % type(feat))
            if i in keys:
                feats[i] = x
        pc = m.pred_convs
        n = image.shape[0]
        counts = []
        for j, k in enumerate(keys):
            _, _, d, h, w_ = feats[k].shape
            counts.append(d * h * w_ * pc.n_boxes[k])
        total = int(sum(counts))
        locs = torch.empty((n, total, 6), dtype=torch.float32, device=dev)
        scores = torch.empty((n, total, pc.n_classes), dtype=torch.float32, device=dev)
        off = 0
        packed = pc._pack()
        for j, k in enumerate(keys):
            pc.run_head(j, feats[k], locs, scores, off, flag)
            tape["heads"].append(dict(j=j, layer=k, feat=feats[k], w=packed[j][0], off=off, bpl=pc.n_boxes[k]))
            off += counts[j]
        self.tape = tape
        return locs, scores

    # ------------------------------------------------------------------------------------------
    def backward(self, dlocs: torch.Tensor, dscores: torch.Tensor, grads: _Grads) -> None:
        """d(loss)/d(locs), d(loss)/d(scores) -> every parameter gradient (written into ``grads``)."""
        tape = self.tape
        if tape is None:
            raise RuntimeError("TrainEngine.backward without a recorded forward")
        self.tape = None
        m = self.model
        pc = m.pred_convs
        dlocs = dlocs.float().contiguous()
        dscores = dscores.float().contiguous()
        n = tape["image"].shape[0]
        head_at = {h["layer"]: h for h in tape["heads"]}
        dO = {}
        for h in tape["heads"]:
            j = h["j"]
            _, c, d, hh, w = h["feat"].shape
            bpl = h["bpl"]
            dO[h["layer"]] = ops.head_grad_pack(dlocs, dscores, n, d, hh, w, bpl, pc.n_classes, h["off"],
                                                grads["pred_convs.loc_convs.%d.bias" % j],
                                                grads["pred_convs.cl_convs.%d.bias" % j])
            ops.head_wgrad(dO[h["layer"]], h["feat"], bpl * 6, bpl * pc.n_classes,
                           grads["pred_convs.loc_convs.%d.weight" % j], grads["pred_convs.cl_convs.%d.weight" % j])
        g = None    # gradient w.r.t. the output of the unit being processed (channels-last bf16)
        for u in reversed(tape["units"]):
            i = u["idx"]
            p = "base.features.%d" % i
            if i in head_at:
                h = head_at[i]
                g = ops.head_dgrad(dO[i], h["w"], h["feat"], addend=g)
            if g is None:
                raise RuntimeError("no gradient reaches backbone layer %d" % i)
            if u["kind"] == "block":
                dz2 = ops.bn_relu_backward(u["z2"], g, u["st2"], grads[p + ".bn2.weight"], grads[p + ".bn2.bias"])
                ops.pwconv_wgrad(dz2, u["a1"], grads[p + ".conv2.weight"])
                g1 = torch.empty_like(u["a1"])
                nn_, c1, d1, h1, w1 = u["a1"].shape
                wt = u["wp"].t().contiguous()           # (Cin, Cout): data gradient = dz . W
                ops.pw_gemm_raw(nn_ * d1 * h1 * w1, dz2, wt, g1)
                dz1 = ops.bn_relu_backward(u["z1"], g1, u["st1"], grads[p + ".bn1.weight"], grads[p + ".bn1.bias"])
                ops.dwconv3d_wgrad(dz1, u["x"], u["stride"], grads[p + ".conv1.weight"])
                g = ops.dwconv3d_dgrad(dz1, u["wd"], u["x"], u["stride"])
            else:
                dz = ops.bn_relu_backward(u["z"], g, u["st"], grads[p + ".1.weight"], grads[p + ".1.bias"])
                ops.stem_wgrad(dz, u["x"], u["stride"], grads[p + ".0.weight"])
                g = None


class _NetFn(torch.autograd.Function):
    """The whole network as one autograd node: forward = TrainEngine.forward, backward = TrainEngine.backward
    returning one gradient per parameter (autograd accumulates them into ``p.grad``)."""

    @staticmethod
    def forward(ctx, engine, image, names, *params):
        ctx.engine, ctx.names, ctx.params = engine, names, params
        locs, scores = engine.forward(image)
        return locs, scores

    @staticmethod
    def backward(ctx, dlocs, dscores):
        eng = ctx.engine
        grads = _Grads(zip(ctx.names, ctx.params))
        eng.backward(dlocs, dscores, grads)
        out = []
        for name, p in zip(ctx.names, ctx.params):
            out.append(grads.t.get(name) if name != "rescale_factors" else None)
        return (None, None, None, *out)


def forward_train(model, image):
    """Autograd-visible train-mode forward of ``model`` (an ``LSSD3D``)."""
    eng = model.train_engine()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    names = tuple(n for n, _ in named)
    return _NetFn.apply(eng, image, names, *[p for _, p in named])


def cosine_lr(base_lr: float, step: int, t_max: int = 40, eta_min: float = 0.0) -> float:
    """CosineAnnealingLR(T_max) in closed form (ssd3d.py:718-720; stepped once per batch, ssd3d.py:525-527)."""
    return eta_min + (base_lr - eta_min) * (1.0 + math.cos(math.pi * step / t_max)) / 2.0


def fit_step(model, batch, world_size: int = 1, allreduce=None):
    """One fused optimisation step: forward, MultiBox loss + its gradient, backward into the flat gradient
    buffer, (all-reduce), fused Adam.  Returns a (2,) device tensor [conf_loss, loc_loss] -- no host sync."""
    eng = model.train_engine()
    flat = eng.flatten()
    dev = model.device
    images, gt_boxes, gt_labels = batch["img"], batch["boxes"], batch["labels"]
    gt_boxes = [b.to(dev) for b in gt_boxes]
    gt_labels = [l.to(dev) for l in gt_labels]
    with torch.no_grad():
        locs, scores = eng.forward(images)
        lf = model.loss_fn
        lf_m = lf.match(gt_boxes, gt_labels)
        out, n_pos, g_locs, g_scores = ops.multibox_loss(locs, scores, lf_m["true_classes"], lf_m["true_locs"],
                                                         alpha=float(lf.alpha), hard_negative_mining=lf.hard_negative_mining,
                                                         neg_pos_ratio=lf.neg_pos_ratio, want_grads=True)
        grads = _Grads(((n, p) for n, p in model.named_parameters() if p.requires_grad and n != "rescale_factors"),
                       flat)
        eng.backward(g_locs, g_scores, grads)
        if allreduce is not None:
            allreduce(flat.grad)
        flat.step += 1
        lr = float(model.lr)
        if model.scheduler != "none":
            # the reference steps the scheduler inside training_step, i.e. before the optimizer step of the
            # same batch (ssd3d.py:525-527): optimizer step k (1-based) runs at the k-th scheduled rate
            lr = cosine_lr(lr, flat.step)
        ops.adam_step(flat.param, flat.grad, flat.exp_avg, flat.exp_avg_sq, flat.bias_start, lr, 2.0 * lr, flat.step,
                      weight_decay=0.0005, grad_scale=1.0 / float(world_size))
        model.invalidate_packed()
    return out
