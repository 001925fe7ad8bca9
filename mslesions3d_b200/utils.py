"""Box geometry with the reference's function names and argument meaning (``lesions3d/utils.py:42-149``),
computed by the CUDA kernels of ``csrc/box_ops.cu``.

Inputs are CUDA float tensors of shape (n, 6); results are bit-identical with the reference's chain of
fp32 torch ops (each step separately rounded).  CPU tensors raise: the accelerated path has no CPU twin.
"""
from __future__ import annotations

import torch

from . import _lib, ops


def cxcycz_to_xyz(cxcycz: torch.Tensor) -> torch.Tensor:
    """Centre-size (cx, cy, cz, w, h, d) -> boundary (x_min, y_min, z_min, x_max, y_max, z_max).  utils.py:42-51."""
    return ops.box_transform(_lib.BOX_CXCYCZ_TO_XYZ, cxcycz)


def xyz_to_cxcycz(xy: torch.Tensor) -> torch.Tensor:
    """Boundary -> centre-size coordinates.  utils.py:92-102."""
    return ops.box_transform(_lib.BOX_XYZ_TO_CXCYCZ, xy)


def gcxgcygcz_to_cxcycz(gcxgcygcz: torch.Tensor, priors_cxcycz: torch.Tensor) -> torch.Tensor:
    """Decode predicted offsets w.r.t. the priors into centre-size boxes.  utils.py:54-68."""
    return ops.box_transform(_lib.BOX_GCXGCYGCZ_TO_CXCYCZ, gcxgcygcz, priors_cxcycz)


def cxcycz_to_gcxgcygcz(cxcycz: torch.Tensor, priors_cxcycz: torch.Tensor) -> torch.Tensor:
    """Encode centre-size boxes w.r.t. the priors (variances 10 and 5).  utils.py:71-89."""
    return ops.box_transform(_lib.BOX_CXCYCZ_TO_GCXGCYGCZ, cxcycz, priors_cxcycz)


def find_intersection3d(set_1: torch.Tensor, set_2: torch.Tensor) -> torch.Tensor:
    """Intersection volume of every box pair, (n1, n2).  utils.py:105-122."""
    return ops.iou3d_pairwise(set_1, set_2, want_iou=False)


def find_jaccard_overlap3d(set_1: torch.Tensor, set_2: torch.Tensor) -> torch.Tensor:
    """Jaccard overlap (IoU) of every box pair, (n1, n2).  utils.py:125-149."""
    return ops.iou3d_pairwise(set_1, set_2, want_iou=True)
