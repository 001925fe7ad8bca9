"""CPU oracle for the SSD3D hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A restatement, in plain torch-CPU fp32 ops, of the reference's algorithm for the
path BASELINE.json names.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import it; the
product package ``mslesions3d_b200`` never does.

Parity pin: the reference ships no tests or golden vectors of its own (SURVEY.md
section 8c), so this oracle is pinned against outputs of the UNMODIFIED reference
run in the build container (``tests/golden/make_golden.py`` -> committed
``tests/golden/*.pt``; ``tests/test_oracle_golden.py`` replays them anywhere, and
``tests/test_oracle_vs_reference.py`` re-runs the live comparison when
``/root/reference`` is mounted).

Every function cites the reference lines (relative to ``lesions3d/``) it follows.
Each arithmetic step is one separately-rounded fp32 torch op in the reference's
order, so on the same torch build the results are bit-identical to the reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# channel / repeat / stride plan of the 3-D MobileNet (mobilenet.py:13-20)
MOBILENET_PLAN = ((64, 1, 2), (128, 2, 2), (256, 2, 2), (512, 6, 2), (1024, 2, 1))
STEM_CHANNELS = 32
DEFAULT_ASPECT_RATIOS = {3: [1.0], 5: [1.0], 7: [1]}  # ssd3d.py:25
BOXES_PER_LOCATION = 2  # hard-coded, ssd3d.py:213
BN_EPS = 1e-5


# --------------------------------------------------------------------------- #
# architecture bookkeeping
# --------------------------------------------------------------------------- #
def backbone_layers(in_channels: int, cube: bool, last_layer: int, width_mult: float = 1.0):
    """Layer list of the truncated backbone (ssd3d.py:56-75, mobilenet.py:26-41).

    Returns a list of dicts: index 0 is the stem (dense 3^3 conv), the others are
    depthwise-separable blocks.  ``stride`` is a 3-tuple.
    """
    c_in = int(STEM_CHANNELS * width_mult)
    first = (2, 2, 2) if cube else (1, 2, 2)
    layers = [dict(kind="stem", cin=in_channels, cout=c_in, stride=first)]
    for c, n, s in MOBILENET_PLAN:
        if len(layers) - 1 == last_layer:
            break
        c_out = int(c * width_mult)
        for i in range(n):
            if len(layers) - 1 == last_layer:
                break
            st = s if i == 0 else 1
            layers.append(dict(kind="block", cin=c_in, cout=c_out, stride=(st, st, st)))
            c_in = c_out
    return layers


def conv_out_dim(d: int, stride: int) -> int:
    # 3-wide kernel, padding 1: floor((d + 2 - 3) / s) + 1
    return (d + 2 - 3) // stride + 1


def feature_map_dims(input_size: Sequence[int], layers) -> List[Tuple[int, int, int]]:
    """Spatial size after every layer (what get_feature_map_infos measures, ssd3d.py:102-110)."""
    dims = []
    cur = tuple(int(v) for v in input_size)
    for L in layers:
        cur = tuple(conv_out_dim(cur[a], L["stride"][a]) for a in range(3))
        dims.append(cur)
    return dims


def default_scales(aspect_ratios: Dict[int, list], input_size, min_object_size=6, max_object_size=14):
    """ssd3d.py:228-232 -- float64 linspace over input_size[0] only."""
    keys = list(aspect_ratios.keys())
    vals = np.linspace(min_object_size / input_size[0], max_object_size / input_size[0], len(keys))
    return {k: v for k, v in zip(keys, vals)}


def prior_boxes(input_size, aspect_ratios=None, scales=None, in_channels=1,
                min_object_size=6, max_object_size=14) -> torch.Tensor:
    """Default boxes in centre-size form, (P, 6) fp32 (ssd3d.py:286-342).

    Note the axis quirk kept from ssd3d.py:304-309: coordinate 0 follows array
    axis 1, coordinate 1 follows array axis 0.
    """
    if not aspect_ratios:
        aspect_ratios = DEFAULT_ASPECT_RATIOS
    if not scales:
        scales = default_scales(aspect_ratios, input_size, min_object_size, max_object_size)
    cube = input_size[0] == input_size[1] == input_size[2]
    layers = backbone_layers(in_channels, cube, max(aspect_ratios.keys()))
    dims = feature_map_dims(input_size, layers)
    rows = []
    for f in aspect_ratios.keys():
        d0, d1, d2 = dims[f]
        s = scales[f]
        for i in range(d0):
            cy = (i + 0.5) / d0
            for j in range(d1):
                cx = (j + 0.5) / d1
                for k in range(d2):
                    cz = (k + 0.5) / d2
                    for ratio in aspect_ratios[f]:
                        rows.append([cx, cy, cz, s, s, s])
                        if ratio == 1.0:
                            for div in range(1, BOXES_PER_LOCATION):
                                extra = s + s / div
                                rows.append([cx, cy, cz, extra, extra, extra])
    out = torch.tensor(np.asarray(rows, dtype=np.float64), dtype=torch.float32)
    return out.clamp_(0, 1)


def prior_boxes_fast(input_size, aspect_ratios=None, scales=None, in_channels=1,
                     min_object_size=6, max_object_size=14) -> torch.Tensor:
    """Vectorised float64 equivalent of :func:`prior_boxes` for multi-million-prior configs."""
    if not aspect_ratios:
        aspect_ratios = DEFAULT_ASPECT_RATIOS
    if not scales:
        scales = default_scales(aspect_ratios, input_size, min_object_size, max_object_size)
    cube = input_size[0] == input_size[1] == input_size[2]
    layers = backbone_layers(in_channels, cube, max(aspect_ratios.keys()))
    dims = feature_map_dims(input_size, layers)
    chunks = []
    for f in aspect_ratios.keys():
        d0, d1, d2 = dims[f]
        s = float(scales[f])
        cy = (np.arange(d0, dtype=np.float64) + 0.5) / d0
        cx = (np.arange(d1, dtype=np.float64) + 0.5) / d1
        cz = (np.arange(d2, dtype=np.float64) + 0.5) / d2
        CY, CX, CZ = np.meshgrid(cy, cx, cz, indexing="ij")
        sizes = []
        for ratio in aspect_ratios[f]:
            sizes.append(s)
            if ratio == 1.0:
                for div in range(1, BOXES_PER_LOCATION):
                    sizes.append(s + s / div)
        per = np.empty((d0, d1, d2, len(sizes), 6), dtype=np.float64)
        per[..., 0] = CX[..., None]
        per[..., 1] = CY[..., None]
        per[..., 2] = CZ[..., None]
        per[..., 3:] = np.asarray(sizes)[None, None, None, :, None]
        chunks.append(per.reshape(-1, 6))
    out = torch.tensor(np.concatenate(chunks, 0), dtype=torch.float32)
    return out.clamp_(0, 1)


# --------------------------------------------------------------------------- #
# network forward (mobilenet.py:26-49, ssd3d.py:86-100,143-169,248-263)
# --------------------------------------------------------------------------- #
def _bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _bn_relu(x, sd, prefix, emulate_bf16):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    if not emulate_bf16:
        return F.relu(F.batch_norm(x, rm, rv, w, b, False, 0.0, BN_EPS))
    # product-path numerics: fp32 scale/shift epilogue, ReLU, round to bf16 storage
    scale = w.float() / torch.sqrt(rv.float() + BN_EPS)
    shift = b.float() - rm.float() * scale
    y = x * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    return _bf16(F.relu(y))


def forward(sd: Dict[str, torch.Tensor], image: torch.Tensor, aspect_ratios=None,
            n_classes: int = 2, emulate_bf16: bool = False, return_features: bool = False):
    """``LSSD3D.forward`` in eval mode -> (locs (N,P,6), scores (N,P,n_classes)).

    ``emulate_bf16=True`` additionally rounds the input, the conv weights and every
    stored activation to bf16 (the product path's storage precision) while
    accumulating in fp32 -- used for the tight GPU-vs-oracle comparison.
    """
    if not aspect_ratios:
        aspect_ratios = DEFAULT_ASPECT_RATIOS
    q = _bf16 if emulate_bf16 else (lambda t: t)
    cube = image.shape[2] == image.shape[3] == image.shape[4]
    layers = backbone_layers(image.shape[1], cube, max(aspect_ratios.keys()))
    x = q(image.float())
    feats = {}
    for i, L in enumerate(layers):
        p = "base.features.%d" % i
        if L["kind"] == "stem":
            x = F.conv3d(x, q(sd[p + ".0.weight"].float()), None, L["stride"], 1)
            x = _bn_relu(x, sd, p + ".1", emulate_bf16)
        else:
            x = F.conv3d(x, q(sd[p + ".conv1.weight"].float()), None, L["stride"], 1, 1, L["cin"])
            x = _bn_relu(x, sd, p + ".bn1", emulate_bf16)
            x = F.conv3d(x, q(sd[p + ".conv2.weight"].float()), None, 1, 0)
            x = _bn_relu(x, sd, p + ".bn2", emulate_bf16)
        if i in aspect_ratios:
            feats[i] = x
    n = image.shape[0]
    locs, scores = [], []
    for hi, f in enumerate(aspect_ratios.keys()):
        lw, lb = sd["pred_convs.loc_convs.%d.weight" % hi], sd["pred_convs.loc_convs.%d.bias" % hi]
        cw, cb = sd["pred_convs.cl_convs.%d.weight" % hi], sd["pred_convs.cl_convs.%d.bias" % hi]
        l = F.conv3d(feats[f], q(lw.float()), lb.float(), 1, 1)
        c = F.conv3d(feats[f], q(cw.float()), cb.float(), 1, 1)
        locs.append(l.permute(0, 2, 3, 4, 1).reshape(n, -1, 6))
        scores.append(c.permute(0, 2, 3, 4, 1).reshape(n, -1, n_classes))
    out = (torch.cat(locs, 1), torch.cat(scores, 1))
    if return_features:
        return out + (feats,)
    return out


# --------------------------------------------------------------------------- #
# box geometry (utils.py:42-149)
# --------------------------------------------------------------------------- #
def cxcycz_to_xyz(c: torch.Tensor) -> torch.Tensor:
    """utils.py:50-51."""
    half = c[:, 3:] / 2
    return torch.cat([c[:, :3] - half, c[:, :3] + half], 1)


def xyz_to_cxcycz(b: torch.Tensor) -> torch.Tensor:
    """utils.py:101-102."""
    return torch.cat([(b[:, 3:] + b[:, :3]) / 2, b[:, 3:] - b[:, :3]], 1)


def gcxgcygcz_to_cxcycz(g: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """utils.py:67-68 -- (g*p_wh)/10 + p_c ; exp(g/5)*p_wh."""
    return torch.cat([g[:, :3] * priors[:, 3:] / 10 + priors[:, :3],
                      torch.exp(g[:, 3:] / 5) * priors[:, 3:]], 1)


def cxcycz_to_gcxgcygcz(c: torch.Tensor, priors: torch.Tensor) -> torch.Tensor:
    """utils.py:88-89."""
    return torch.cat([(c[:, :3] - priors[:, :3]) / (priors[:, 3:] / 10),
                      torch.log(c[:, 3:] / priors[:, 3:]) * 5], 1)


def find_intersection3d(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils.py:119-122."""
    lo = torch.max(a[:, None, :3], b[None, :, :3])
    hi = torch.min(a[:, None, 3:], b[None, :, 3:])
    d = torch.clamp(hi - lo, min=0)
    return d[:, :, 0] * d[:, :, 1] * d[:, :, 2]


def find_jaccard_overlap3d(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """utils.py:135-149 -- inter / ((vol_a + vol_b) - inter), every op rounded separately."""
    inter = find_intersection3d(a, b)
    va = (a[:, 3] - a[:, 0]) * (a[:, 4] - a[:, 1]) * (a[:, 5] - a[:, 2])
    vb = (b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1]) * (b[:, 5] - b[:, 2])
    union = va[:, None] + vb[None, :] - inter
    return inter / union


def iou_rows_chunked(boxes: torch.Tensor, i0: int, i1: int) -> torch.Tensor:
    return find_jaccard_overlap3d(boxes[i0:i1], boxes)


# --------------------------------------------------------------------------- #
# detection (ssd3d.py:344-460)
# --------------------------------------------------------------------------- #
def greedy_nms(boxes_sorted: torch.Tensor, max_overlap: float, row_chunk: int = 2048) -> torch.Tensor:
    """ssd3d.py:407-426 -- returns the boolean keep mask over score-sorted boxes.

    The reference materialises the full n x n IoU; here rows are produced in
    chunks (identical values) so that n ~ 1e5 stays within memory.
    """
    n = boxes_sorted.shape[0]
    suppress = np.zeros(n, dtype=bool)
    for c0 in range(0, n, row_chunk):
        c1 = min(n, c0 + row_chunk)
        over = (iou_rows_chunked(boxes_sorted, c0, c1) > max_overlap).numpy()
        for r in range(c0, c1):
            if suppress[r]:
                continue
            suppress |= over[r - c0]
            suppress[r] = False
    return torch.from_numpy(~suppress)


def detect_objects(predicted_locs, predicted_scores, priors_cxcycz, min_score, max_overlap, top_k,
                   stable: bool = True, return_indices: bool = False):
    """``LSSD3D.detect_objects`` (ssd3d.py:344-460).

    ``stable=True`` fixes the tie rule the reference leaves unspecified (SURVEY.md
    M8): equal scores keep ascending prior index.  With tie-free scores this is
    exactly the reference.  With ``return_indices`` a fourth list carries, per
    image, the prior index of every returned detection (-1 for the placeholder).
    """
    n_img, n_priors, n_classes = predicted_scores.shape
    assert n_priors == priors_cxcycz.shape[0] == predicted_locs.shape[1]
    probs = F.softmax(predicted_scores, dim=2)                                   # ssd3d.py:363
    decoded = torch.stack([cxcycz_to_xyz(gcxgcygcz_to_cxcycz(predicted_locs[i], priors_cxcycz))
                           for i in range(n_img)])                               # ssd3d.py:373-374
    return detect_from_decoded(probs, decoded, min_score, max_overlap, top_k, stable, return_indices)


def detect_from_decoded(probs, decoded, min_score, max_overlap, top_k, stable: bool = True,
                        return_indices: bool = False):
    """The integer-exact part of ``detect_objects`` (ssd3d.py:376-453): filter, sort, truncate,
    greedy NMS, per-image concatenation and top-k, on GIVEN class probabilities (N,P,C) and decoded
    boundary boxes (N,P,6).  The stage-wise parity tests feed it the device's own softmax/decode
    output, so that exp() implementation differences cannot leak into the index comparison."""
    n_img, n_priors, n_classes = probs.shape
    out_b, out_l, out_s, out_i = [], [], [], []
    for i in range(n_img):
        ib, il, isc, ii = [], [], [], []
        for c in range(1, n_classes):
            cs = probs[i][:, c]
            above = cs > min_score
            n_above = int(above.sum())
            if n_above == 0:
                continue
            idx = torch.nonzero(above).flatten()
            cs = cs[above]
            cb = decoded[i][above]
            cs, order = cs.sort(dim=0, descending=True, stable=stable)
            cb = cb[order]
            idx = idx[order]
            n_keep = min(10 * top_k, n_above)
            cs, cb, idx = cs[:n_keep], cb[:n_keep], idx[:n_keep]
            keep = greedy_nms(cb, max_overlap)
            ib.append(cb[keep])
            il.append(torch.full((int(keep.sum()),), c, dtype=torch.long))
            isc.append(cs[keep])
            ii.append(idx[keep])
        if not ib:
            ib.append(torch.tensor([[0., 0., 0., 1., 1., 1.]]))
            il.append(torch.zeros(1, dtype=torch.long))
            isc.append(torch.zeros(1))
            ii.append(torch.full((1,), -1, dtype=torch.long))
        ib, il, isc, ii = torch.cat(ib, 0), torch.cat(il, 0), torch.cat(isc, 0), torch.cat(ii, 0)
        if isc.shape[0] > top_k:
            isc, order = isc.sort(dim=0, descending=True, stable=stable)
            isc = isc[:top_k]
            ib = ib[order][:top_k]
            il = il[order][:top_k]
            ii = ii[order][:top_k]
        out_b.append(ib)
        out_l.append(il)
        out_s.append(isc)
        out_i.append(ii)
    if return_indices:
        return out_b, out_l, out_s, out_i
    return out_b, out_l, out_s


# --------------------------------------------------------------------------- #
# prior <-> ground-truth matching and MultiBox loss (ssd3d.py:741-941)
# --------------------------------------------------------------------------- #
def parse_threshold(threshold):
    """ssd3d.py:762-773 -> (mode, t0, t1)."""
    if isinstance(threshold, list):
        if len(threshold) == 1:
            return "hard", float(threshold[0]), float(threshold[0])
        assert len(threshold) == 2
        return "soft", float(threshold[0]), float(threshold[1])
    if isinstance(threshold, float):
        return "hard", threshold, threshold
    raise Exception("Type error. Expected float or list of floats for threshold")


def match_image(boxes: torch.Tensor, labels: torch.Tensor, priors_cxcycz: torch.Tensor, threshold,
                chunk: int = 100):
    """One image of ssd3d.py:786-888 -> (true_classes (P,), true_locs (P,6), overlap (P,), obj (P,), prior_for_obj)."""
    mode, t0, t1 = parse_threshold(threshold)
    priors_xyz = cxcycz_to_xyz(priors_cxcycz)
    n_obj = boxes.shape[0]
    ov_parts, obj_parts, pfo_parts = [], [], []
    for s in range(0, n_obj, chunk):
        part = find_jaccard_overlap3d(boxes[s:s + chunk], priors_xyz)
        ov, ob = part.max(dim=0)
        _, pf = part.max(dim=1)
        ov_parts.append(ov.view(1, -1))
        obj_parts.append((ob + s).view(1, -1))
        pfo_parts.append(pf)
    prior_for_obj = torch.cat(pfo_parts)
    ov_all = torch.cat(ov_parts)
    obj_all = torch.cat(obj_parts)
    overlap, which = ov_all.max(dim=0)
    obj = obj_all.gather(0, which.view(1, -1)).view(-1)
    obj[prior_for_obj] = torch.arange(n_obj)          # ssd3d.py:865 (duplicates: last writer wins)
    overlap[prior_for_obj] = 1.0                      # ssd3d.py:868
    lab = labels[obj]
    if mode == "hard":
        lab[overlap < t0] = 0
    else:
        lab[overlap < t0] = 0
        lab[(overlap >= t0) & (overlap < t1)] = -1
    true_locs = cxcycz_to_gcxgcygcz(xyz_to_cxcycz(boxes[obj]), priors_cxcycz)
    return lab, true_locs, overlap, obj, prior_for_obj


def multibox_loss(predicted_locs, predicted_scores, boxes: List[torch.Tensor], labels: List[torch.Tensor],
                  priors_cxcycz, threshold=0.5, neg_pos_ratio=3, hard_negative_mining=False,
                  return_targets=False):
    """``MultiBoxLoss.forward`` (ssd3d.py:775-941) -> (conf_loss, loc_loss).

    Default follows the shipped code: confidence loss summed over ALL non-ignored
    priors / n_positives (ssd3d.py:933).  ``hard_negative_mining=True`` is the
    variant of the commented-out lines ssd3d.py:926-932.
    """
    n_img, n_priors, n_classes = predicted_scores.shape
    true_locs = torch.zeros((n_img, n_priors, 6), dtype=torch.float)
    true_classes = torch.zeros((n_img, n_priors), dtype=torch.long)
    for i in range(n_img):
        if boxes[i].shape[0] == 0:
            continue
        lab, tl, _, _, _ = match_image(boxes[i], labels[i], priors_cxcycz, threshold)
        true_classes[i] = lab
        true_locs[i] = tl
    pos = true_classes > 0
    loc_loss = (predicted_locs[pos] - true_locs[pos]).abs().mean()
    n_pos = pos.sum(dim=1)
    tc = true_classes.view(-1).clone()
    tc[tc == -1] = 0
    ce = F.cross_entropy(predicted_scores.view(-1, n_classes), tc, reduction="none").view(n_img, n_priors)
    ce = ce.clone()
    ce[true_classes < 0] = 0
    ce_pos = ce[pos]
    ce_neg = ce.clone()
    ce_neg[pos] = 0.0
    if hard_negative_mining:
        ce_neg, _ = ce_neg.sort(dim=1, descending=True)
        ranks = torch.arange(n_priors).unsqueeze(0).expand_as(ce_neg)
        hard = ranks < (neg_pos_ratio * n_pos).unsqueeze(1)
        conf_loss = (ce_neg[hard].sum() + ce_pos.sum()) / n_pos.sum().float()
    else:
        conf_loss = (ce_neg.sum() + ce_pos.sum()) / n_pos.sum().float()
    if torch.isnan(loc_loss):
        raise Exception("Loss is NaN")
    if return_targets:
        return conf_loss, loc_loss, true_classes, true_locs
    return conf_loss, loc_loss


# --------------------------------------------------------------------------- #
# weights
# --------------------------------------------------------------------------- #
def random_state_dict(in_channels=1, aspect_ratios=None, n_classes=2, seed=0, randomize_bn=True):
    """Random-init weights with the reference's 103 state-dict keys/shapes: input
    generation, shared with the benchmark, lives in ``mslesions3d_b200.synthetic`` (the oracle is the checker, not
    the source of inputs)."""
    from mslesions3d_b200 import synthetic
    return synthetic.random_state_dict(in_channels, aspect_ratios, n_classes, seed, randomize_bn)


# --------------------------------------------------------------------------- #
# training step (ssd3d.py:467-531, 704-722): train-mode forward, loss, autograd, Adam
# --------------------------------------------------------------------------- #
BN_MOMENTUM = 0.1  # nn.BatchNorm3d default, mobilenet.py:29,39,41


def _ste_bf16(t: torch.Tensor) -> torch.Tensor:
    """bf16 storage rounding that is transparent to autograd (straight-through)."""
    return t + (_bf16(t.detach()) - t.detach())


class _RoundGradBF16(torch.autograd.Function):
    """Identity whose backward rounds the gradient to bf16: marks the places where the product path stores an
    activation gradient (after each BN/ReLU backward and after each data-gradient kernel)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf16(g)


def _rg(t: torch.Tensor) -> torch.Tensor:
    return _RoundGradBF16.apply(t)


def _bn_relu_train(x, params, running, prefix, emulate_bf16):
    """nn.BatchNorm3d in training mode + ReLU (mobilenet.py:29-30,39,41,44-45): batch statistics, running
    statistics updated in ``running`` (momentum 0.1, unbiased variance)."""
    if emulate_bf16:
        x = _rg(_ste_bf16(x))  # the raw conv output (and later its gradient) is stored in bf16
    y = F.batch_norm(x, running[prefix + ".running_mean"], running[prefix + ".running_var"],
                     params[prefix + ".weight"], params[prefix + ".bias"], True, BN_MOMENTUM, BN_EPS)
    running[prefix + ".num_batches_tracked"] = running[prefix + ".num_batches_tracked"] + 1
    y = F.relu(y)
    return _rg(_ste_bf16(y)) if emulate_bf16 else y


def forward_train(params: Dict[str, torch.Tensor], running: Dict[str, torch.Tensor], image: torch.Tensor,
                  aspect_ratios=None, n_classes: int = 2, emulate_bf16: bool = False):
    """``LSSD3D.forward`` in train mode.  ``params`` holds the learnable tensors (they may require grad),
    ``running`` the BN buffers, UPDATED IN PLACE like the module's.  -> (locs, scores)."""
    if not aspect_ratios:
        aspect_ratios = DEFAULT_ASPECT_RATIOS
    q = _ste_bf16 if emulate_bf16 else (lambda t: t)
    cube = image.shape[2] == image.shape[3] == image.shape[4]
    layers = backbone_layers(image.shape[1], cube, max(aspect_ratios.keys()))
    x = _bf16(image.float()) if emulate_bf16 else image.float()
    feats = {}
    for i, L in enumerate(layers):
        p = "base.features.%d" % i
        rgi = _rg if emulate_bf16 else (lambda t: t)   # a data-gradient kernel's output is bf16
        if L["kind"] == "stem":
            x = F.conv3d(x, q(params[p + ".0.weight"]), None, L["stride"], 1)
            x = _bn_relu_train(x, params, running, p + ".1", emulate_bf16)
        else:
            x = F.conv3d(rgi(x), q(params[p + ".conv1.weight"]), None, L["stride"], 1, 1, L["cin"])
            x = _bn_relu_train(x, params, running, p + ".bn1", emulate_bf16)
            x = F.conv3d(rgi(x), q(params[p + ".conv2.weight"]), None, 1, 0)
            x = _bn_relu_train(x, params, running, p + ".bn2", emulate_bf16)
        if i in aspect_ratios:
            feats[i] = x
    rgo = _rg if emulate_bf16 else (lambda t: t)       # head gradient rows are packed to bf16
    n = image.shape[0]
    locs, scores = [], []
    for hi, f in enumerate(aspect_ratios.keys()):
        l = rgo(F.conv3d(feats[f], q(params["pred_convs.loc_convs.%d.weight" % hi]), None, 1, 1)) \
            + params["pred_convs.loc_convs.%d.bias" % hi].view(1, -1, 1, 1, 1)
        c = rgo(F.conv3d(feats[f], q(params["pred_convs.cl_convs.%d.weight" % hi]), None, 1, 1)) \
            + params["pred_convs.cl_convs.%d.bias" % hi].view(1, -1, 1, 1, 1)
        locs.append(l.permute(0, 2, 3, 4, 1).reshape(n, -1, 6))
        scores.append(c.permute(0, 2, 3, 4, 1).reshape(n, -1, n_classes))
    return torch.cat(locs, 1), torch.cat(scores, 1)


def split_state_dict(sd: Dict[str, torch.Tensor]):
    """-> (params requiring grad, BN buffers), both fresh fp32 clones.  ``rescale_factors`` is a parameter of
    the reference that never receives a gradient."""
    params, running = {}, {}
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var") or k.endswith("num_batches_tracked"):
            running[k] = v.clone()
        else:
            params[k] = v.clone().float().requires_grad_(True)
    return params, running


def train_step_grads(sd, image, boxes, labels, priors_cxcycz, threshold=0.5, alpha=1.0, aspect_ratios=None,
                     n_classes=2, emulate_bf16=False):
    """One ``training_step`` + backward of the reference (ssd3d.py:467-494): train-mode forward, MultiBox
    loss, ``(conf + alpha*loc).backward()``.
    -> dict(conf, loc, locs, scores, grads {name: tensor | None}, running {BN buffers after the step})."""
    params, running = split_state_dict(sd)
    locs, scores = forward_train(params, running, image, aspect_ratios, n_classes, emulate_bf16)
    conf, loc = multibox_loss(locs, scores, boxes, labels, priors_cxcycz, threshold)
    (conf + alpha * loc).backward()
    return dict(conf=conf.detach(), loc=loc.detach(), locs=locs.detach(), scores=scores.detach(),
                grads={k: (v.grad.clone() if v.grad is not None else None) for k, v in params.items()},
                running=running, params=params)


def make_optimizer(params: Dict[str, torch.Tensor], lr: float):
    """``configure_optimizers`` (ssd3d.py:704-722): Adam, biases at 2x lr, weight decay 5e-4, cosine schedule
    (T_max 40) stepped once per batch."""
    biases = [p for n, p in params.items() if n.endswith(".bias")]
    others = [p for n, p in params.items() if not n.endswith(".bias")]
    opt = torch.optim.Adam([{"params": biases, "lr": 2 * lr}, {"params": others}], lr=lr, weight_decay=0.0005)
    sch = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=40)
    return opt, sch


def fit_steps(sd, batches, priors_cxcycz, threshold, lr, alpha=1.0, aspect_ratios=None, n_classes=2):
    """A few optimisation steps the way pytorch-lightning drives the reference: per batch zero_grad,
    training_step (which steps the scheduler, ssd3d.py:525-527), backward, optimizer.step.
    -> (state dict after the steps, list of (conf, loc))."""
    params, running = split_state_dict(sd)
    opt, sch = make_optimizer(params, lr)
    losses = []
    for image, boxes, labels in batches:
        opt.zero_grad(set_to_none=True)
        locs, scores = forward_train(params, running, image, aspect_ratios, n_classes)
        conf, loc = multibox_loss(locs, scores, boxes, labels, priors_cxcycz, threshold)
        sch.step()
        (conf + alpha * loc).backward()
        opt.step()
        losses.append((float(conf), float(loc)))
    out = {k: v.detach().clone() for k, v in params.items()}
    out.update({k: v.clone() for k, v in running.items()})
    return out, losses


# --------------------------------------------------------------------------- #
# detection metrics (utils.py:155-396)
# --------------------------------------------------------------------------- #
def metrics_per_class(det_images, det_boxes, det_scores, true_images, true_boxes, true_difficulties, min_overlap,
                      stable: bool = True):
    """``compute_metrics_per_class`` (utils.py:155-230): walk the detections of one class in descending score
    order; a detection is a true positive when its best-overlapping object of the same image (first maximum)
    exceeds ``min_overlap`` (strict), is not difficult and has not been claimed by an earlier detection.
    -> (tp, fp, detected, sorted_scores, sort_index)."""
    n_det = det_boxes.shape[0]
    detected = torch.zeros(true_boxes.shape[0], dtype=torch.uint8)
    det_scores, order = torch.sort(det_scores, dim=0, descending=True, stable=stable)
    det_images, det_boxes = det_images[order], det_boxes[order]
    tp, fp = torch.zeros(n_det), torch.zeros(n_det)
    all_idx = torch.arange(true_boxes.shape[0])
    for d in range(n_det):
        same = true_images == det_images[d]
        cand = true_boxes[same]
        if cand.shape[0] == 0:
            fp[d] = 1
            continue
        ov = find_jaccard_overlap3d(det_boxes[d:d + 1], cand).squeeze(0)
        best, ind = torch.max(ov, dim=0)
        orig = all_idx[same][ind]
        if best.item() > min_overlap:
            if true_difficulties[same][ind] == 0:
                if detected[orig] == 0:
                    tp[d] = 1
                    detected[orig] = 1
                else:
                    fp[d] = 1
        else:
            fp[d] = 1
    return tp, fp, detected, det_scores, order


def average_precision_11pt(tp, fp, n_easy):
    """utils.py:296-318 -> (AP, cumulative precision, cumulative recall, precision at the 11 thresholds)."""
    ctp, cfp = torch.cumsum(tp, 0), torch.cumsum(fp, 0)
    prec = ctp / (ctp + cfp + 1e-10)
    rec = ctp / n_easy
    thr = torch.arange(start=0, end=1.1, step=.1).tolist()
    p11 = torch.zeros(len(thr))
    for i, t in enumerate(thr):
        above = rec >= t
        p11[i] = prec[above].max() if above.any() else 0.
    return p11.mean(), prec, rec, p11


def calculate_map(det_boxes, det_labels, det_scores, true_boxes, true_labels, true_difficulties, min_overlap=0.5,
                  n_classes=2, stable: bool = True):
    """``calculate_mAP`` (utils.py:233-396) for ``n_classes`` (the reference hard-codes 2 via its label map).
    -> dict(mAP, per class c: AP, tp, fp, detected, sorted_scores, precision, recall, f1, volumes)."""
    t_img = torch.cat([torch.full((l.shape[0],), i, dtype=torch.long) for i, l in enumerate(true_labels)])
    t_box, t_lab, t_dif = torch.cat(true_boxes), torch.cat(true_labels), torch.cat(true_difficulties)
    d_img = torch.cat([torch.full((l.shape[0],), i, dtype=torch.long) for i, l in enumerate(det_labels)])
    d_box, d_lab, d_sc = torch.cat(det_boxes), torch.cat(det_labels), torch.cat(det_scores)
    aps = torch.zeros(n_classes - 1)
    out = {}
    for c in range(1, n_classes):
        st, sd_ = t_lab == c, d_lab == c
        if int(sd_.sum()) == 0:
            continue
        tp, fp, detected, scores, order = metrics_per_class(d_img[sd_], d_box[sd_], d_sc[sd_], t_img[st], t_box[st],
                                                            t_dif[st], min_overlap, stable)
        n_easy = int(t_dif[st].logical_not().sum())
        ap, prec, rec, p11 = average_precision_11pt(tp, fp, n_easy)
        aps[c - 1] = ap
        fn = 1 - detected
        r = tp.sum() / (tp.sum() + fn.sum())
        p = tp.sum() / (tp.sum() + fp.sum())
        vols = torch.tensor([float((b[3] - b[0]) * (b[4] - b[1]) * (b[5] - b[2])) for b in t_box[st]])
        out[c] = dict(AP=ap, tp=tp, fp=fp, detected=detected, sorted_scores=scores, order=order, recall=r, precision=p,
                      f1=(2 * p * r) / (p + r), cum_precision=prec, cum_recall=rec, p11=p11, volumes=vols)
    out["mAP"] = aps.mean().item()
    return out


# ---------------------------------------------------------------------------------------------------
# Ground-truth boxes from a segmentation (utils.py:438-513, BoundingBoxesGeneratord "binary" / "classes")
# ---------------------------------------------------------------------------------------------------
def _label_face_connected(mask: np.ndarray):
    """Face-connected components of a 3-D boolean mask, numbered 1.. in C order of each component's first voxel
    (what ``scipy.ndimage.label`` with its default structure returns, utils.py:13,447,462).  Plain flood fill."""
    lab = np.zeros(mask.shape, dtype=np.int64)
    D, H, W = mask.shape
    n = 0
    for start in np.argwhere(mask):          # argwhere is in C order
        if lab[tuple(start)]:
            continue
        n += 1
        lab[tuple(start)] = n
        stack = [tuple(start)]
        while stack:
            d, h, w = stack.pop()
            for dd, hh, ww in ((d - 1, h, w), (d + 1, h, w), (d, h - 1, w), (d, h + 1, w), (d, h, w - 1), (d, h, w + 1)):
                if 0 <= dd < D and 0 <= hh < H and 0 <= ww < W and mask[dd, hh, ww] and not lab[dd, hh, ww]:
                    lab[dd, hh, ww] = n
                    stack.append((dd, hh, ww))
    return lab, n


def gt_boxes_from_segmentation(seg, n_classes: int = 0):
    """One volume (D, H, W) -> (boxes (n, 6) fp32, labels (n,) int64).  ``n_classes == 0``: "binary" mode (non-zero
    voxels, label 1, utils.py:446-449); otherwise "classes" mode with classes 1..n_classes (utils.py:451-468).
    Boxes are [min index, max index] / image size (utils.py:500,472); zero-volume boxes are dropped
    (utils.py:475-480).  A volume without objects gives empty tensors (the reference raises there)."""
    seg = np.squeeze(np.asarray(seg))
    boxes, labels = [], []
    classes = [1] if n_classes == 0 else list(range(1, n_classes + 1))
    for ci, c in enumerate(classes):
        mask = (seg != 0) if n_classes == 0 else (seg == c)
        lab, n = _label_face_connected(mask)
        for k in range(1, n + 1):
            idx = np.argwhere(lab == k)
            boxes.append(list(idx.min(0)) + list(idx.max(0)))
            labels.append(ci + 1)
    if not boxes:
        return torch.zeros((0, 6)), torch.zeros((0,), dtype=torch.long)
    b = torch.tensor(boxes, dtype=torch.float32) / torch.tensor(list(seg.shape) * 2, dtype=torch.float32)
    l = torch.tensor(labels, dtype=torch.long)
    vol = (b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1]) * (b[:, 5] - b[:, 2])
    keep = vol != 0.0
    return b[keep], l[keep]


def gt_boxes_from_instances(seg, thresholds):
    """``BoundingBoxesGeneratord`` "instances" mode (utils.py:439-441,483-513) on one volume (D, H, W) whose voxels
    hold one id per object: per class c (id range [min_c, max_c)), one box [min index, max index] / image size per
    id in ascending id order (np.unique), label c+1; zero-volume boxes dropped (utils.py:475-480)."""
    seg = np.squeeze(np.asarray(seg))
    ids = [v for v in np.unique(seg) if v != 0]                                 # utils.py:491-492
    boxes, labels = [], []
    for c, (lo, hi) in enumerate(thresholds):                                   # utils.py:493-511
        for v in ids:
            if lo <= v < hi:
                idx = np.argwhere(seg == v)
                boxes.append(list(idx.min(0)) + list(idx.max(0)))
                labels.append(c + 1)
    if not boxes:
        return torch.zeros((0, 6)), torch.zeros((0,), dtype=torch.long)
    b = torch.tensor(boxes, dtype=torch.float32) / torch.tensor(list(seg.shape) * 2, dtype=torch.float32)
    l = torch.tensor(labels, dtype=torch.long)
    vol = (b[:, 3] - b[:, 0]) * (b[:, 4] - b[:, 1]) * (b[:, 5] - b[:, 2])
    keep = vol != 0.0
    return b[keep], l[keep]


def greedy_nms_grid(boxes_sorted: torch.Tensor, max_overlap: float) -> torch.Tensor:
    """The same keep mask as :func:`greedy_nms` (ssd3d.py:407-426) for lists far beyond what the n x n formulation
    can hold on the CPU (hundreds of thousands of boxes), for ``max_overlap >= 0`` and finite, well-formed boxes.

    Boxes are visited in score order; the kept ones are hashed by the cell of their minimum corner (cell edge = the
    largest extent of any box, so two intersecting boxes are at most one cell apart on every axis), and a box is
    tested, with the reference's exact fp32 IoU arithmetic (utils.py:105-149, one rounded op at a time), only
    against the kept boxes of the 27 surrounding cells.  IoU > max_overlap >= 0 needs a positive intersection, so
    no other kept box can suppress it.  Checked against :func:`greedy_nms` in tests/test_oracle_golden.py."""
    if not max_overlap >= 0:
        raise ValueError("greedy_nms_grid needs max_overlap >= 0")
    b = boxes_sorted.numpy().astype(np.float32)
    n = b.shape[0]
    keep = np.zeros(n, dtype=bool)
    if n == 0:
        return torch.from_numpy(keep)
    thr = np.float32(max_overlap)
    ext = b[:, 3:] - b[:, :3]
    cell = float(max(ext.max(), 1e-6))
    lo = b[:, :3].min(0)
    cells = np.floor((b[:, :3] - lo) / cell).astype(np.int64)
    vol = (ext[:, 0] * ext[:, 1]) * ext[:, 2]                                  # utils.py:142-147
    buckets: Dict[Tuple[int, int, int], List[int]] = {}
    offsets = [(dx, dy, dz) for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]
    for i in range(n):
        cx, cy, cz = (int(v) for v in cells[i])
        cand: List[int] = []
        for dx, dy, dz in offsets:
            lst = buckets.get((cx + dx, cy + dy, cz + dz))
            if lst:
                cand.extend(lst)
        suppressed = False
        if cand:
            o = b[cand]
            d = np.minimum(b[i, 3:], o[:, 3:]) - np.maximum(b[i, :3], o[:, :3])  # utils.py:119-121
            d = np.where(d < 0, np.float32(0), d)
            inter = (d[:, 0] * d[:, 1]) * d[:, 2]
            union = (vol[i] + vol[cand]) - inter                                  # utils.py:148
            with np.errstate(divide="ignore", invalid="ignore"):
                suppressed = bool(((inter / union) > thr).any())                  # NaN > thr is False
        if not suppressed:
            keep[i] = True
            buckets.setdefault((cx, cy, cz), []).append(i)
    return torch.from_numpy(keep)
