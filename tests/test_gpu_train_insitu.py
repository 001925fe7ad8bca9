"""Stage-wise IN-SITU parity of the training step.

A real ``LSSD3D.training_step`` + ``loss.backward()`` runs on the device with ``TrainEngine.record`` switched on:
the engine then keeps its tape (every unit's own input activation, raw conv output and BatchNorm state) and a clone
of every gradient that flows through a unit.  Each of the 15 conv -> BN -> ReLU units (stem + 7 x {depthwise,
pointwise}) and each of the 3 heads is then replayed ALONE through torch-CPU autograd (the reference's mechanism,
mobilenet.py:26-49, ssd3d.py:131-167) from the product path's own bf16 tensors:

    conv forward      z   = conv(x_own, w_bf16)                       vs the saved raw output        <= 1 bf16 ulp
    BN + ReLU forward a   = relu(batch_norm(z_own))                   vs the next unit's input       <= 1 bf16 ulp
    BN + ReLU backward (dz, dgamma, dbeta) from the incoming g_own    dz <= 1.5 ulp, dgamma/dbeta <= 2e-3 rel L2
    conv backward     (dW, dx) from dz_own                            dW <= 2e-3 rel L2, dx <= 1.5 ulp

so no stage inherits another stage's rounding, nothing is amplified by the batch-statistic BatchNorms of the deep
maps, and a wrong tap / border / stride in any single kernel shows up as an O(1) error instead of hiding inside a
whole-network tolerance."""
import pytest
import torch
import torch.nn.functional as F

from oracle import ssd3d_oracle as O
from mslesions3d_b200 import synthetic

pytestmark = pytest.mark.gpu

EPS = 1e-5


def bf16r(t):
    return t.to(torch.bfloat16).float()


def rel_l2(got, want):
    got, want = got.float().cpu().flatten(), want.float().cpu().flatten()
    return float((got - want).norm() / want.norm().clamp(min=1e-20))


def bf16_mismatch(got, want, ulps):
    """Fraction of elements further than ``ulps`` bf16 ulp apart (relative to max(|want|, scale/64))."""
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape, (tuple(got.shape), tuple(want.shape))
    scale = float(want.abs().max())
    tol = ulps * (2.0 ** -7) * torch.clamp(want.abs(), min=scale * 2.0 ** -6)
    return float(((got - want).abs() > tol).float().mean())


def check_unit(name, conv, x_own, w_param, z_own, st, a_own, g_own, dz_own, bn_mod, dw_got, dx_got, report):
    """One conv -> BN -> ReLU unit from the product path's own tensors; ``conv(x, w)`` is the torch-CPU conv."""
    x_own, z_own, a_own, g_own, dz_own = (t.float().cpu() for t in (x_own, z_own, a_own, g_own, dz_own))
    w = bf16r(w_param.detach().float().cpu())
    # conv forward
    z_ref = conv(bf16r(x_own), w)
    bad = bf16_mismatch(z_own, bf16r(z_ref), 1.0)
    assert bad == 0.0, "%s conv forward: %.2e of elements beyond 1 bf16 ulp" % (name, bad)
    # BN (batch statistics) + ReLU forward and backward on the product's own z and incoming gradient
    zl = z_own.clone().requires_grad_(True)
    gamma = bn_mod.weight.detach().float().cpu().clone().requires_grad_(True)
    beta = bn_mod.bias.detach().float().cpu().clone().requires_grad_(True)
    a_ref = F.relu(F.batch_norm(zl, None, None, gamma, beta, True, 0.0, EPS))
    bad = bf16_mismatch(a_own, bf16r(a_ref.detach()), 1.0)
    assert bad <= 1e-5, "%s BN+ReLU forward: %.2e of elements beyond 1 bf16 ulp" % (name, bad)
    mean, var = z_own.transpose(0, 1).flatten(1).mean(1), z_own.transpose(0, 1).flatten(1).var(1, unbiased=False)
    torch.testing.assert_close(st.mean.cpu(), mean, rtol=1e-4, atol=1e-5, msg=name + " batch mean")
    torch.testing.assert_close(st.invstd.cpu(), 1.0 / torch.sqrt(var + EPS), rtol=1e-4, atol=1e-6, msg=name + " invstd")
    a_ref.backward(g_own)
    # ReLU ties: z is bf16, so several elements of a channel share one value; where that value's pre-activation
    # z*scale+shift cancels to within a few fp32 ulps of its terms, its sign (the ReLU mask) depends on how the
    # implementation rounds scale / shift -- not an error of either side.  Such elements take the reference value.
    with torch.no_grad():
        shape = (1, -1, 1, 1, 1)
        sc64 = (gamma.double() / torch.sqrt(var.double() + EPS)).reshape(shape)
        zs = z_own.double() * sc64
        sh64 = (beta.double() - mean.double() * sc64.flatten()).reshape(shape)
        tie = (zs + sh64).abs() <= 4e-6 * (zs.abs() + sh64.abs())
        assert float(tie.float().mean()) <= 1e-3, "%s: implausibly many ReLU ties" % name
        dz_cmp = torch.where(tie, bf16r(zl.grad), dz_own)
    bad = bf16_mismatch(dz_cmp, bf16r(zl.grad), 1.5)
    r_g, r_b = rel_l2(bn_mod.weight.grad, gamma.grad), rel_l2(bn_mod.bias.grad, beta.grad)
    # a ReLU whose pre-activation rounds to the other side of 0 flips one element: allow a vanishing fraction
    assert bad <= 2e-5, "%s BN+ReLU backward dz: %.2e of elements beyond 1.5 bf16 ulp" % (name, bad)
    assert r_g <= 2e-3 and r_b <= 2e-3, "%s dgamma %.2e dbeta %.2e" % (name, r_g, r_b)
    # conv backward on the product's own dz
    xl = bf16r(x_own).requires_grad_(dx_got is not None)
    wl = w.clone().requires_grad_(True)
    conv(xl, wl).backward(dz_own)
    r_w = rel_l2(dw_got, wl.grad)
    assert r_w <= 2e-3, "%s dW rel L2 %.2e" % (name, r_w)
    bad_x = 0.0
    if dx_got is not None:
        bad_x = bf16_mismatch(dx_got, bf16r(xl.grad), 1.5)
        assert bad_x == 0.0, "%s dx: %.2e of elements beyond 1.5 bf16 ulp" % (name, bad_x)
    report.append("%-28s dgamma %.1e dbeta %.1e dW %.1e" % (name, r_g, r_b, r_w))


@pytest.mark.parametrize("n_classes,channels,size,batch", [(2, 1, (64, 64, 64), 4), (3, 2, (32, 64, 48), 3)])
def test_training_step_stagewise_in_situ(n_classes, channels, size, batch):
    from mslesions3d_b200.ssd3d import LSSD3D
    torch.set_num_threads(8)
    sd = O.random_state_dict(channels, n_classes=n_classes, seed=21)
    x, b, l = synthetic.make_batch(batch, channels, size, first_idx=5, with_boxes=True)
    g = torch.Generator().manual_seed(3)
    boxes = [torch.from_numpy(v) for v in b]
    labels = [torch.randint(1, n_classes, (v.shape[0],), generator=g) for v in b]
    model = LSSD3D(n_classes=n_classes, input_channels=channels, input_size=size, threshold=[0.1, 0.2])
    model.load_state_dict(sd)
    model = model.cuda().train()
    eng = model.train_engine()
    eng.record = {}
    out = model.training_step({"img": torch.from_numpy(x), "boxes": boxes, "labels": labels})
    out["loss"].backward()
    torch.cuda.synchronize()
    rec, eng.record = eng.record, None
    tape, units = rec["tape"], rec["tape"]["units"]
    assert len(units) == 8 and len(tape["heads"]) == 3
    report = []
    feats_after = {}                                     # output activation of every backbone layer
    for k, u in enumerate(units):
        nxt = units[k + 1]["x"] if k + 1 < len(units) else tape["heads"][-1]["feat"]
        feats_after[u["idx"]] = nxt
    for u in units:
        i = u["idx"]
        r = rec["units"][i]
        feat = model.base.features[i]
        if u["kind"] == "stem":
            sd_ = u["stride"]
            check_unit("f0 stem", lambda a, w: F.conv3d(a, w, None, (sd_, 2, 2), 1), u["x"], feat[0].weight, u["z"],
                       u["st"], feats_after[i], r["g"], r["dz"], feat[1], feat[0].weight.grad, None, report)
            continue
        s, c = u["stride"], u["x"].shape[1]
        check_unit("f%d depthwise" % i, lambda a, w: F.conv3d(a, w, None, s, 1, 1, c), u["x"], feat.conv1.weight,
                   u["z1"], u["st1"], u["a1"], r["g1"], r["dz1"], feat.bn1, feat.conv1.weight.grad, r["dx"], report)
        check_unit("f%d pointwise" % i, lambda a, w: F.conv3d(a, w), u["a1"], feat.conv2.weight, u["z2"], u["st2"],
                   feats_after[i], r["g2"], r["dz2"], feat.bn2, feat.conv2.weight.grad, r["g1"], report)
    # ---- heads: weight / bias / data gradients from the loss gradient rows the kernels really used ----
    pc = model.pred_convs
    dlocs, dscores = rec["dlocs"].cpu(), rec["dscores"].cpu()
    loc_norms = []
    for h in tape["heads"]:
        j, layer, bpl = h["j"], h["layer"], h["bpl"]
        hr = rec["heads"][layer]
        feat = h["feat"].float().cpu()
        n, c, d, hh, w = feat.shape
        nl, nc = bpl * 6, bpl * n_classes
        rows = hr["dO"].float().cpu()                       # (G, M, 16): bf16-rounded gradient rows
        groups = rows.shape[0]
        rows = rows.permute(1, 0, 2).reshape(n * d * hh * w, groups * 16)
        cnt = d * hh * w * bpl
        want_rows = torch.cat([dlocs[:, h["off"]:h["off"] + cnt].reshape(-1, nl),
                               dscores[:, h["off"]:h["off"] + cnt].reshape(-1, nc)], 1)
        assert torch.equal(rows[:, :nl + nc], bf16r(want_rows)) and not bool(rows[:, nl + nc:].any())
        xl = feat.clone().requires_grad_(True)
        lw = bf16r(pc.loc_convs[j].weight.detach().float().cpu()).requires_grad_(True)
        cw = bf16r(pc.cl_convs[j].weight.detach().float().cpu()).requires_grad_(True)
        ol, oc = F.conv3d(xl, lw, None, 1, 1), F.conv3d(xl, cw, None, 1, 1)
        g_l = rows[:, :nl].reshape(n, d, hh, w, nl).permute(0, 4, 1, 2, 3)
        g_c = rows[:, nl:nl + nc].reshape(n, d, hh, w, nc).permute(0, 4, 1, 2, 3)
        torch.autograd.backward([ol, oc], [g_l, g_c])
        r_l, r_c = rel_l2(pc.loc_convs[j].weight.grad, lw.grad), rel_l2(pc.cl_convs[j].weight.grad, cw.grad)
        assert float(cw.grad.norm()) > 0
        loc_norms.append(float(lw.grad.norm()))         # a layer without a positive prior has no loc gradient
        assert r_l <= 2e-3 and r_c <= 2e-3, "head %d dW: loc %.2e class %.2e" % (j, r_l, r_c)
        torch.testing.assert_close(pc.loc_convs[j].bias.grad.cpu(), want_rows[:, :nl].sum(0), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(pc.cl_convs[j].bias.grad.cpu(), want_rows[:, nl:].sum(0), rtol=1e-4, atol=1e-5)
        want_dx = xl.grad + (hr["addend"].float().cpu() if hr["addend"] is not None else 0.0)
        bad = bf16_mismatch(hr["out"], bf16r(want_dx), 1.5)
        assert bad == 0.0, "head %d data gradient: %.2e of elements off" % (j, bad)
        report.append("head %d (layer %d, %d cols)     dW loc %.1e class %.1e" % (j, layer, nl + nc, r_l, r_c))
    assert max(loc_norms) > 0, "no localisation gradient reached any head"
    print("\n".join(report))
