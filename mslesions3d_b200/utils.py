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


# ----------------------------------------------------------------------------------------------
# detection metrics (utils.py:155-396), on the device
# ----------------------------------------------------------------------------------------------
voc_labels = tuple(["lesion"])                                  # utils.py:27
label_map = {k: v + 1 for v, k in enumerate(voc_labels)}        # utils.py:28-30
label_map['background'] = 0
rev_label_map = {v: k for k, v in label_map.items()}


def volume(box):
    """utils.py:150-152."""
    return (box[3] - box[0]) * (box[4] - box[1]) * (box[5] - box[2])


def _recall_thresholds():
    return torch.arange(start=0, end=1.1, step=.1)              # utils.py:303


def compute_metrics_per_class(det_class_images, det_class_boxes, det_class_scores, true_class_images,
                              true_class_boxes, true_class_difficulties, min_overlap):
    """True / false positives of one class's detections in descending score order (utils.py:155-230), computed by
    ``csrc/metrics.cu`` without the reference's per-detection Python loop and host syncs.  Same six return values:
    true_positives, false_positives, true_class_boxes_detected, sorted scores, volumes of the found and of the
    not-found objects.  Equal scores are ordered by ascending input index."""
    m = ops.map_class(det_class_boxes, det_class_scores, det_class_images, true_class_boxes, true_class_difficulties,
                      true_class_images, min_overlap, _recall_thresholds())
    return _metrics_tuple(m, true_class_difficulties)


def _metrics_tuple(m, true_class_difficulties):
    easy = true_class_difficulties.to(m["detected"].device).logical_not()
    vols = m["volumes"][easy]                                   # utils.py:219 (objects that are not difficult)
    return (m["tp"], m["fp"], m["detected"], m["sorted_scores"], vols[m["detected"] == 1], vols[m["detected"] == 0])


def calculate_mAP(det_boxes, det_labels, det_scores, true_boxes, true_labels, true_difficulties, min_overlap=0.5,
                  return_detail=False):
    """Mean average precision of detected objects (utils.py:233-396): same arguments (lists of per-image tensors)
    and the same return value -- ``(APs, mAP)`` or, with ``return_detail``, the reference's dictionary."""
    assert len(det_boxes) == len(det_labels) == len(det_scores) == len(true_boxes) == len(true_labels) == len(
        true_difficulties)
    n_classes = len(label_map)
    dev = det_boxes[0].device if len(det_boxes) else torch.device("cuda")
    true_images = torch.cat([torch.full((int(l.size(0)),), i, dtype=torch.int32) for i, l in enumerate(true_labels)]
                            or [torch.zeros(0, dtype=torch.int32)]).to(dev)
    true_boxes = torch.cat([b.to(dev).reshape(-1, 6) for b in true_boxes], dim=0)
    true_labels = torch.cat([l.to(dev) for l in true_labels], dim=0)
    true_difficulties = torch.cat([d.to(dev) for d in true_difficulties], dim=0)
    assert true_images.size(0) == true_boxes.size(0) == true_labels.size(0)
    det_images = torch.cat([torch.full((int(l.size(0)),), i, dtype=torch.int32) for i, l in enumerate(det_labels)]
                           or [torch.zeros(0, dtype=torch.int32)]).to(dev)
    det_boxes = torch.cat([b.reshape(-1, 6) for b in det_boxes], dim=0)
    det_labels = torch.cat(list(det_labels), dim=0)
    det_scores = torch.cat(list(det_scores), dim=0)
    assert det_images.size(0) == det_boxes.size(0) == det_labels.size(0) == det_scores.size(0)

    average_precisions = torch.zeros((n_classes - 1), dtype=torch.float)
    true_positives_per_class, false_positives_per_class, true_boxes_detected_per_class = {}, {}, {}
    found_boxes_volumes_per_class, not_found_boxes_volumes_per_class, sorted_scores_per_class = {}, {}, {}
    recalls_per_class, precisions_per_class, f1_scores_per_class = {}, {}, {}
    n_easy_class_objects = 0
    thresholds = _recall_thresholds()
    stats = {}
    for c in range(1, n_classes):
        sel_t = true_labels == c
        true_class_images, true_class_boxes = true_images[sel_t], true_boxes[sel_t]
        true_class_difficulties = true_difficulties[sel_t]
        sel_d = det_labels == c
        det_class_images, det_class_boxes, det_class_scores = det_images[sel_d], det_boxes[sel_d], det_scores[sel_d]
        if det_class_boxes.size(0) == 0:
            continue
        m = ops.map_class(det_class_boxes, det_class_scores, det_class_images, true_class_boxes,
                          true_class_difficulties, true_class_images, min_overlap, thresholds)
        tp, fp, detected, sorted_scores, found, not_found = _metrics_tuple(m, true_class_difficulties)
        true_positives_per_class[c], false_positives_per_class[c] = tp, fp
        true_boxes_detected_per_class[c] = detected
        found_boxes_volumes_per_class[c], not_found_boxes_volumes_per_class[c] = found, not_found
        sorted_scores_per_class[c] = sorted_scores
        stats[c] = m["stats"]
    host = {c: v.cpu() for c, v in stats.items()}                        # the one device->host read
    for c, v in host.items():
        average_precisions[c - 1] = v[0]
        recalls_per_class[c], precisions_per_class[c], f1_scores_per_class[c] = v[1], v[2], v[3]
    mean_average_precision = average_precisions.mean().item()
    average_precisions = {rev_label_map[c + 1]: v for c, v in enumerate(average_precisions.tolist())}
    if n_classes == 2:
        try:
            recalls_per_class = recalls_per_class[1]
            precisions_per_class = precisions_per_class[1]
            f1_scores_per_class = f1_scores_per_class[1]
            average_precisions = average_precisions[list(average_precisions.keys())[0]]
            true_boxes_detected_per_class = true_boxes_detected_per_class[1]
            found_boxes_volumes_per_class = found_boxes_volumes_per_class[1]
            not_found_boxes_volumes_per_class = not_found_boxes_volumes_per_class[1]
            true_positives_per_class = true_positives_per_class[1]
            false_positives_per_class = false_positives_per_class[1]
        except KeyError:   # no detected objects (utils.py:371-381)
            recalls_per_class = 0.
            precisions_per_class = 0.
            f1_scores_per_class = 0.
            average_precisions = 0.
            n_easy_class_objects = int(true_difficulties[true_labels == n_classes - 1].logical_not().sum())
            true_boxes_detected_per_class = torch.zeros(n_easy_class_objects, dtype=torch.uint8).to(dev)
            true_positives_per_class = torch.Tensor([]).to(dev)
            false_positives_per_class = torch.Tensor([]).to(dev)
            found_boxes_volumes_per_class = torch.Tensor([]).to(dev)
            not_found_boxes_volumes_per_class = torch.FloatTensor([float(volume(b)) for b in true_boxes.cpu()]).to(dev)
    if not return_detail:
        return average_precisions, mean_average_precision
    return {"APs": average_precisions,
            "mAP": mean_average_precision,
            "precision": precisions_per_class,
            "recall": recalls_per_class,
            "f1_score": f1_scores_per_class,
            "sorted_det_scores": sorted_scores_per_class,
            "TP": true_positives_per_class,
            "FP": false_positives_per_class,
            "n_true_boxes": true_boxes_detected_per_class.size(0),
            "found_boxes_volumes_per_class": found_boxes_volumes_per_class,
            "not_found_boxes_volumes_per_class": not_found_boxes_volumes_per_class,
            }


class BoundingBoxesGeneratord:
    """Device-side mirror of the reference's dictionary transform (utils.py:398-513): ``d["boxes"]``,
    ``d["labels"]`` from the segmentation under each key.  ``segmentation_mode`` "binary" (every non-zero voxel,
    label 1) and "classes" (voxels equal to c, for ``classes = [1..n_classes]`` as the data modules pass them,
    datasets.py:405) and "instances" (pre-labelled volumes, ``thresholds`` = per-class id ranges [min, max)) all
    run on the GPU.

    The segmentation may be one volume ((D,H,W) or (1,D,H,W): boxes (n,6), labels (n,), like the reference) or a
    batch (N,1,D,H,W): lists of per-volume tensors, the format ``training_step`` consumes (ssd3d.py:470-472)."""

    def __init__(self, keys, allow_missing_keys: bool = False, segmentation_mode: str = "instances", thresholds=None,
                 classes=None, n_classes: int = None, max_boxes: int = 1024):
        self.keys = [keys] if isinstance(keys, str) else list(keys)
        self.allow_missing_keys = allow_missing_keys
        if n_classes is not None and not classes:
            classes = list(range(1, n_classes + 1))
        if not n_classes and classes:
            n_classes = len(classes)
        assert segmentation_mode in ["instances", "binary", "classes"]
        assert segmentation_mode != "binary" or (not classes and not n_classes) or n_classes == 1
        assert segmentation_mode != "classes" or type(classes) in [list, dict]
        assert segmentation_mode != "instances" or thresholds          # utils.py:417
        if segmentation_mode == "classes" and list(classes) != list(range(1, n_classes + 1)):
            raise NotImplementedError("classes must be [1..n_classes] (datasets.py:405)")
        self.segmentation_mode = segmentation_mode
        self.thresholds = thresholds
        self.classes = classes
        self.n_classes = n_classes
        self.max_boxes = max_boxes

    def converter(self, seg: torch.Tensor):
        from . import ops
        batched = seg.dim() == 5
        s = seg if batched else seg.reshape((1,) + tuple(seg.shape[-3:]))
        if self.segmentation_mode == "instances":
            boxes, labels = ops.gt_boxes_from_instances(s, self.thresholds, self.max_boxes)
        else:
            boxes, labels = ops.gt_boxes_from_segmentation(
                s, 0 if self.segmentation_mode == "binary" else int(self.n_classes), self.max_boxes)
        return (boxes, labels) if batched else (boxes[0], labels[0])

    def __call__(self, data):
        d = dict(data)
        for key in self.keys:
            if key not in d:
                if self.allow_missing_keys:
                    continue
                raise KeyError(key)
            d["boxes"], d["labels"] = self.converter(d[key])
        return d

    def inverse(self, data):
        return dict(data)
