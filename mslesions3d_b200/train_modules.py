"""Training-mode forward + backward of the building blocks called STAND-ALONE (``conv_bn`` / ``Block`` /
``PredictionConvolutions`` outside ``LSSD3D``), as the reference allows: there they are plain ``nn.Module``s
whose train-mode forward uses batch statistics and whose backward comes from autograd
(mobilenet.py:26-49, ssd3d.py:113-169).

Each module is ONE ``torch.autograd.Function`` over the same kernels ``training.TrainEngine`` drives for the
whole network (raw conv -> batch-statistic BatchNorm + ReLU; BN/ReLU backward, weight and data gradients):
the saved activations live on the autograd context, so any number of modules / forwards may be in flight.
Activations and activation gradients are channels-last bf16, parameter gradients fp32.  ``LSSD3D`` itself keeps
using the fused engine (one node for the whole network, CUDA-graph capturable).
"""
from __future__ import annotations

import torch

from . import ops


def _grad_like(g: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    """Incoming gradient as a private channels-last bf16 buffer (the BN backward overwrites it in place)."""
    return ops.to_channels_last_bf16(g).clone(memory_format=torch.preserve_format)


class ConvBNTrainFn(torch.autograd.Function):
    """conv_bn in training mode: dense 3x3x3 conv -> BatchNorm3d (batch statistics) -> ReLU (mobilenet.py:26-31).
    The input is the image: there is no data gradient (``x.requires_grad`` is refused by the caller)."""

    @staticmethod
    def forward(ctx, mod, x, weight, gamma, beta):
        conv, bn = mod[0], mod[1]
        sd = ops_stride(conv.stride)[0]
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        z = ops.stem_conv_raw(x, ops.pack_stem_weight(weight), sd)
        a, st = ops.bn_train_relu(z, bn, None)
        ctx.mod, ctx.sd, ctx.st = mod, sd, st
        ctx.save_for_backward(x, z)
        ctx.wshape = tuple(weight.shape)
        return a

    @staticmethod
    def backward(ctx, g):
        x, z = ctx.saved_tensors
        dev = z.device
        c = z.shape[1]
        dgamma = torch.empty((c,), dtype=torch.float32, device=dev)
        dbeta = torch.empty((c,), dtype=torch.float32, device=dev)
        dz = ops.bn_relu_backward(z, _grad_like(g, z), ctx.st, dgamma, dbeta)
        dw = torch.empty(ctx.wshape, dtype=torch.float32, device=dev)
        ops.stem_wgrad(dz, x, ctx.sd, dw)
        return None, None, dw, dgamma, dbeta


class BlockTrainFn(torch.autograd.Function):
    """Block in training mode: depthwise 3x3x3 -> BN -> ReLU -> pointwise -> BN -> ReLU (mobilenet.py:34-49)."""

    @staticmethod
    def forward(ctx, mod, x, w1, g1, b1, w2, g2, b2):
        s = ops_stride(mod.conv1.stride)[0]
        x_cl = ops.to_channels_last_bf16(x)
        wd = ops.pack_dw_weight(w1)
        wp = ops.pack_pw_weight(w2)
        z1 = ops.dwconv3d_raw(x_cl, wd, s)
        a1, st1 = ops.bn_train_relu(z1, mod.bn1, None)
        z2 = ops.pwconv_raw(a1, wp)
        flag = mod.nan_flag
        own = flag is None
        if own:
            flag = torch.zeros((1,), dtype=torch.int32, device=x.device)
        a2, st2 = ops.bn_train_relu(z2, mod.bn2, flag)
        if own and int(flag.item()) != 0:       # stand-alone use keeps the reference's check (mobilenet.py:46-48)
            raise Exception("NaN Loss in MobileNet Block")
        ctx.s, ctx.st1, ctx.st2 = s, st1, st2
        ctx.x_dtype = x.dtype
        ctx.w1shape, ctx.w2shape = tuple(w1.shape), tuple(w2.shape)
        ctx.save_for_backward(x_cl, z1, a1, z2, wd, wp)
        return a2

    @staticmethod
    def backward(ctx, g):
        x, z1, a1, z2, wd, wp = ctx.saved_tensors
        dev = x.device
        cin, cout = z1.shape[1], z2.shape[1]
        f32 = dict(dtype=torch.float32, device=dev)
        dg2, db2 = torch.empty((cout,), **f32), torch.empty((cout,), **f32)
        dz2 = ops.bn_relu_backward(z2, _grad_like(g, z2), ctx.st2, dg2, db2)
        dw2 = torch.empty(ctx.w2shape, **f32)
        ops.pwconv_wgrad(dz2, a1, dw2)
        ga1 = torch.empty_like(a1)
        n, _, d, h, w = a1.shape
        ops.pw_gemm_raw(n * d * h * w, dz2, wp.t().contiguous(), ga1)          # dx = dz . W
        dg1, db1 = torch.empty((cin,), **f32), torch.empty((cin,), **f32)
        dz1 = ops.bn_relu_backward(z1, ga1, ctx.st1, dg1, db1)
        dw1 = torch.empty(ctx.w1shape, **f32)
        ops.dwconv3d_wgrad(dz1, x, ctx.s, dw1)
        dx = None
        if ctx.needs_input_grad[1]:
            dx = ops.dwconv3d_dgrad(dz1, wd, x, ctx.s).to(ctx.x_dtype)
        return None, dx, dw1, dg1, db1, dw2, dg2, db2


class HeadsTrainFn(torch.autograd.Function):
    """PredictionConvolutions with gradients (ssd3d.py:134-169): inputs = the feature maps followed by
    (loc weight, loc bias, class weight, class bias) per head; outputs locs (N,P,6), scores (N,P,n_classes)."""

    @staticmethod
    def forward(ctx, mod, nan_flag, n_feats, *tensors):
        feats = [ops.to_channels_last_bf16(t) for t in tensors[:n_feats]]
        params = tensors[n_feats:]
        keys = list(mod.aspect_ratios.keys())
        n = feats[0].shape[0]
        dev = feats[0].device
        counts = [f.shape[2] * f.shape[3] * f.shape[4] * mod.n_boxes[keys[i]] for i, f in enumerate(feats)]
        total = int(sum(counts))
        locs = torch.empty((n, total, 6), dtype=torch.float32, device=dev)
        scores = torch.empty((n, total, mod.n_classes), dtype=torch.float32, device=dev)
        packed, off = [], 0
        for i, f in enumerate(feats):
            lw, lb, cw, cb = params[4 * i:4 * i + 4]
            w, b = ops.pack_head_weight(lw, lb, cw, cb)
            ops.head_conv(f, w, b, locs, scores, mod.n_boxes[keys[i]], mod.n_classes, off, nan_flag)
            packed.append(w)
            off += counts[i]
        ctx.mod, ctx.n_feats, ctx.counts = mod, n_feats, counts
        ctx.in_dtypes = [t.dtype for t in tensors[:n_feats]]
        ctx.shapes = [tuple(p.shape) for p in params]
        ctx.save_for_backward(*feats, *packed)
        return locs, scores

    @staticmethod
    def backward(ctx, dlocs, dscores):
        mod, nf = ctx.mod, ctx.n_feats
        feats, packed = ctx.saved_tensors[:nf], ctx.saved_tensors[nf:]
        keys = list(mod.aspect_ratios.keys())
        dev = feats[0].device
        dlocs, dscores = dlocs.float().contiguous(), dscores.float().contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        d_feats, d_params, off = [], [], 0
        for i, f in enumerate(feats):
            n, c, d, h, w = f.shape
            bpl = mod.n_boxes[keys[i]]
            shp = ctx.shapes[4 * i:4 * i + 4]
            dlw, dlb, dcw, dcb = (torch.empty(s, **f32) for s in shp)
            dO = ops.head_grad_pack(dlocs, dscores, n, d, h, w, bpl, mod.n_classes, off, dlb, dcb)
            ops.head_wgrad(dO, f, bpl * 6, bpl * mod.n_classes, dlw, dcw)
            if ctx.needs_input_grad[3 + i]:
                d_feats.append(ops.head_dgrad(dO, packed[i], f, n_cols=bpl * (6 + mod.n_classes)).to(ctx.in_dtypes[i]))
            else:
                d_feats.append(None)
            d_params += [dlw, dlb, dcw, dcb]
            off += ctx.counts[i]
        return (None, None, None, *d_feats, *d_params)


def ops_stride(stride):
    if isinstance(stride, int):
        return (stride, stride, stride)
    return tuple(int(s) for s in stride)


def conv_bn_train(mod, x):
    if x.requires_grad:
        raise NotImplementedError("conv_bn (the network stem) has no data-gradient kernel: its input is the image")
    conv, bn = mod[0], mod[1]
    return ConvBNTrainFn.apply(mod, x, conv.weight, bn.weight, bn.bias)


def block_train(mod, x):
    return BlockTrainFn.apply(mod, x, mod.conv1.weight, mod.bn1.weight, mod.bn1.bias, mod.conv2.weight,
                              mod.bn2.weight, mod.bn2.bias)


def heads_train(mod, feats, nan_flag=None):
    keys = list(feats.keys())
    params = []
    for lc, cc in zip(mod.loc_convs, mod.cl_convs):
        params += [lc.weight, lc.bias, cc.weight, cc.bias]
    return HeadsTrainFn.apply(mod, nan_flag, len(keys), *[feats[k] for k in keys], *params)
