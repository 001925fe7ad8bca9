"""In-memory synthetic lesion volumes (the benchmark's input spec).

Restates the distribution of the reference's dataset generator without the
NIfTI round trip (generate_artificial_dataset.py:63-105): uniform noise, then
``randint(lo, hi) + 1`` axis-aligned cubes of side ``randint(smin, smax)`` at
+0.4 intensity, clipped to [0, 1]; followed by the non-zero z-score
normalisation the data module applies (datasets.py:403).  Ground-truth boxes use
the reference's convention ``[min_idx, max_idx] / dims`` in array-axis order with
an inclusive max index (utils.py:472,500).

``make_batch`` & co. are host-side numpy and reproduce the reference's MT19937 stream (the golden vectors and the
benchmark inputs come from them).  ``make_batch_device`` is the same construction generated where it is consumed
(SURVEY.md 8f rank 2): Philox4x32-10 keyed by (seed, volume, channel, voxel) -- the numpy stream is NOT reproduced
there, only the distribution -- followed by the device-side NormalizeIntensity and ground-truth box extraction.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

try:  # connected components are only needed when cubes touch
    from scipy.ndimage import label as _cc_label
except Exception:  # pragma: no cover
    _cc_label = None


def default_object_size(image_size: Sequence[int]) -> Tuple[int, int]:
    """Object-size range scaled with the volume."""
    s = int(image_size[0])
    if tuple(image_size) == (160, 192, 160):
        return 6, 14
    return max(2, round(6 * s / 64)), max(3, round(14 * s / 64))


def generate_volume(idx: int, image_size: Sequence[int] = (64, 64, 64), num_objects=(1, 5),
                    object_size=None, random_seed: int = 0, noise: bool = True):
    """One raw volume + mask + cube list, cf. generate_artificial_dataset.py:63-87.

    Returns (data float64 (D,H,W) in [0,1], mask uint8, cubes [(corner3, side)]).
    """
    if object_size is None:
        object_size = default_object_size(image_size)
    smin, smax = sorted(int(v) for v in object_size)
    rng_seed = random_seed + idx
    np.random.seed(rng_seed)
    data = np.random.rand(*image_size) if noise else np.zeros(tuple(image_size))
    mask = np.zeros(tuple(image_size), dtype=np.uint8)
    n_objects = np.random.randint(*num_objects)
    cubes = []
    for _ in range(n_objects + 1):
        side = np.random.randint(smin, smax)
        _cls = np.random.randint(0, 1)  # keeps the RNG stream aligned with the generator (n_classes=1)
        corner = [np.random.randint(0, image_size[a] - side) for a in range(3)]
        sl = tuple(slice(c, c + side) for c in corner)
        data[sl] = data[sl] + 0.4 if noise else 1.0
        data = data.clip(0, 1)
        mask[sl] = 1
        cubes.append((tuple(corner), side))
    return data, mask, cubes


def normalize_nonzero(x: np.ndarray) -> np.ndarray:
    """MONAI ``NormalizeIntensity(nonzero=True)`` semantics (datasets.py:403)."""
    nz = x != 0
    if not nz.any():
        return x.astype(np.float32)
    sel = x[nz]
    mean, std = sel.mean(), sel.std()
    out = x.astype(np.float64).copy()
    out[nz] = (sel - mean) / (std if std != 0 else 1.0)
    return out.astype(np.float32)


def boxes_from_mask(mask: np.ndarray) -> np.ndarray:
    """GT boxes [min_idx, max_idx]/dims per connected component (utils.py:446-472,500)."""
    dims = np.asarray(mask.shape, dtype=np.float64)
    if _cc_label is None:
        raise RuntimeError("scipy is required for connected-component GT extraction")
    lab, n = _cc_label(mask)
    out = []
    for k in range(1, n + 1):
        pos = np.argwhere(lab == k)
        lo, hi = pos.min(0), pos.max(0)
        if np.any(hi == lo):
            continue  # zero-extent boxes are filtered upstream (utils.py:475-481)
        out.append(np.concatenate([lo / dims, hi / dims]))
    return np.asarray(out, dtype=np.float32).reshape(-1, 6)


def make_batch(batch_size: int, channels: int = 1, image_size: Sequence[int] = (64, 64, 64),
               first_idx: int = 0, random_seed: int = 0, num_objects=(1, 5), object_size=None,
               with_boxes: bool = False):
    """Batch of normalised volumes (N, C, D, H, W) fp32 (+ GT boxes / labels lists).

    Extra channels reuse the same cubes with fresh noise from ``seed + idx + 10**6 * c``
    (the reference's data modules are single-sequence).
    """
    vols = np.empty((batch_size, channels) + tuple(image_size), dtype=np.float32)
    boxes: List[np.ndarray] = []
    labels: List[np.ndarray] = []
    for b in range(batch_size):
        idx = first_idx + b
        data, mask, cubes = generate_volume(idx, image_size, num_objects, object_size, random_seed)
        vols[b, 0] = normalize_nonzero(data)
        for c in range(1, channels):
            rs = np.random.RandomState(random_seed + idx + 10 ** 6 * c)
            extra = rs.rand(*image_size)
            for corner, side in cubes:
                sl = tuple(slice(k, k + side) for k in corner)
                extra[sl] = extra[sl] + 0.4
            vols[b, c] = normalize_nonzero(extra.clip(0, 1))
        if with_boxes:
            bx = boxes_from_mask(mask)
            boxes.append(bx)
            labels.append(np.ones((bx.shape[0],), dtype=np.int64))
    if with_boxes:
        return vols, boxes, labels
    return vols


def make_batch_device(batch_size: int, channels: int = 1, image_size: Sequence[int] = (64, 64, 64), first_idx: int = 0,
                      random_seed: int = 0, num_objects=(1, 5), object_size=None, with_boxes: bool = False,
                      device=None, dtype=None):
    """``make_batch`` without the host: volumes are generated (``ssd3d_generate_volumes``), normalised
    (``ssd3d_normalize_intensity_nonzero``, datasets.py:403) and -- with ``with_boxes`` -- turned into ground-truth
    boxes (``ssd3d_gt_boxes_from_segmentation``, utils.py:438-513: connected components of the mask, so touching
    cubes merge exactly as in the reference's data module) on the device.  Returns CUDA tensors: volumes
    (N, C, D, H, W) in ``dtype`` (bf16 by default, the stem's input format) [+ boxes / labels lists].
    A different random stream than ``make_batch`` (see the module docstring)."""
    import torch
    from . import ops
    if object_size is None:
        object_size = default_object_size(image_size)
    raw, mask, _, _ = ops.generate_volumes(batch_size, channels, image_size, first_idx, random_seed, num_objects,
                                           object_size, device)
    vols = ops.normalize_intensity_nonzero(raw, torch.bfloat16 if dtype is None else dtype)
    if not with_boxes:
        return vols
    boxes, labels = ops.gt_boxes_from_segmentation(mask)
    return vols, boxes, labels


def write_dataset(output_dir: str, num_images: int, image_size: Sequence[int] = (64, 64, 64), num_objects=(1, 5),
                  object_size=None, random_seed: int = 0) -> str:
    """The on-disk data set of generate_artificial_dataset.py:54-111: ``images/sub-XXXX_image.nii.gz`` (raw
    float64 volume) and ``labels/sub-XXXX_seg.nii.gz`` (mask), identity affine."""
    import os
    from .nifti import save_nifti
    image_dir, seg_dir = os.path.join(output_dir, "images"), os.path.join(output_dir, "labels")
    os.makedirs(image_dir, exist_ok=True)
    os.makedirs(seg_dir, exist_ok=True)
    for idx in range(num_images):
        data, mask, _ = generate_volume(idx, image_size, num_objects, object_size, random_seed)
        save_nifti(os.path.join(image_dir, "sub-%s_image.nii.gz" % str(idx).zfill(4)), data)
        save_nifti(os.path.join(seg_dir, "sub-%s_seg.nii.gz" % str(idx).zfill(4)), mask.astype(np.float64))
    return output_dir


def load_dataset_dir(dataset_dir: str, with_boxes: bool = False):
    """Read a data set written by ``write_dataset`` / the reference's generator: -> subject ids, normalised
    volumes (N, 1, D, H, W) fp32 (datasets.py:403), and with ``with_boxes`` the GT boxes / labels extracted from
    the masks like ``BoundingBoxesGeneratord`` does (utils.py:446-500)."""
    import os
    from .nifti import load_nifti
    image_dir, seg_dir = os.path.join(dataset_dir, "images"), os.path.join(dataset_dir, "labels")
    names = sorted(f for f in os.listdir(image_dir) if f.endswith("_image.nii.gz") or f.endswith("_image.nii"))
    subjects, vols, boxes, labels = [], [], [], []
    for f in names:
        sid = f.split("_image")[0].replace("sub-", "")
        data, _ = load_nifti(os.path.join(image_dir, f))
        subjects.append(sid)
        vols.append(normalize_nonzero(data)[None])
        if with_boxes:
            seg_name = f.replace("_image", "_seg")
            mask, _ = load_nifti(os.path.join(seg_dir, seg_name))
            bx = boxes_from_mask((mask != 0).astype(np.uint8))
            boxes.append(bx)
            labels.append(np.ones((bx.shape[0],), dtype=np.int64))
    vols = np.stack(vols).astype(np.float32)
    if with_boxes:
        return subjects, vols, boxes, labels
    return subjects, vols


# --------------------------------------------------------------------------------------------------
# random-init weights of the SSD3D-MobileNet architecture (benchmark / test input generation)
# --------------------------------------------------------------------------------------------------
_MOBILENET_PLAN = ((64, 1, 2), (128, 2, 2), (256, 2, 2), (512, 6, 2), (1024, 2, 1))   # mobilenet.py:13-20
_DEFAULT_ASPECT_RATIOS = {3: [1.0], 5: [1.0], 7: [1]}                                  # ssd3d.py:25


def _backbone_plan(in_channels: int, last_layer: int):
    """(kind, cin, cout) of every backbone layer up to ``last_layer`` (ssd3d.py:56-75)."""
    layers = [("stem", in_channels, 32)]
    c_in = 32
    for c, n, _ in _MOBILENET_PLAN:
        for _i in range(n):
            if len(layers) - 1 == last_layer:
                return layers
            layers.append(("block", c_in, c))
            c_in = c
    return layers


def random_state_dict(in_channels=1, aspect_ratios=None, n_classes=2, seed=0, randomize_bn=True):
    """Random-init weights with the reference's 103 state-dict keys / shapes: conv weights
    ~ U(-b, b), b = sqrt(6 / fan_in); BN affine parameters and running statistics randomised so that folding is
    exercised.  Deterministic in ``seed`` (torch CPU generator)."""
    import math
    import torch
    if not aspect_ratios:
        aspect_ratios = _DEFAULT_ASPECT_RATIOS
    g = torch.Generator().manual_seed(seed)
    layers = _backbone_plan(in_channels, max(aspect_ratios.keys()))
    sd = {}

    def conv_w(co, ci, k):
        bound = math.sqrt(6.0 / (ci * k ** 3))
        return (torch.rand((co, ci, k, k, k), generator=g) * 2 - 1) * bound

    def bn(prefix, c):
        sd[prefix + ".weight"] = 1.0 + 0.2 * (torch.rand(c, generator=g) - 0.5) if randomize_bn else torch.ones(c)
        sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=g) if randomize_bn else torch.zeros(c)
        sd[prefix + ".running_mean"] = 0.1 * torch.randn(c, generator=g) if randomize_bn else torch.zeros(c)
        sd[prefix + ".running_var"] = 0.5 + torch.rand(c, generator=g) if randomize_bn else torch.ones(c)
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0)

    first_pred = min(aspect_ratios.keys())
    sd["rescale_factors"] = torch.full((1, layers[first_pred][2], 1, 1, 1), 20.0)
    for i, (kind, cin, cout) in enumerate(layers):
        p = "base.features.%d" % i
        if kind == "stem":
            sd[p + ".0.weight"] = conv_w(cout, cin, 3)
            bn(p + ".1", cout)
        else:
            sd[p + ".conv1.weight"] = conv_w(cin, 1, 3)
            bn(p + ".bn1", cin)
            sd[p + ".conv2.weight"] = conv_w(cout, cin, 1)
            bn(p + ".bn2", cout)
    for hi, f in enumerate(aspect_ratios.keys()):
        c = layers[f][2]
        nb = len(aspect_ratios[f]) + 1      # boxes_per_location is hard-coded to 2 (ssd3d.py:213)
        sd["pred_convs.loc_convs.%d.weight" % hi] = conv_w(nb * 6, c, 3)
        sd["pred_convs.loc_convs.%d.bias" % hi] = 0.05 * torch.randn(nb * 6, generator=g)
        sd["pred_convs.cl_convs.%d.weight" % hi] = conv_w(nb * n_classes, c, 3)
        sd["pred_convs.cl_convs.%d.bias" % hi] = 0.05 * torch.randn(nb * n_classes, generator=g)
    return sd
