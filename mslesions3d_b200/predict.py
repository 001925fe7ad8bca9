"""``predict.py`` entry point of the reference (``lesions3d/predict.py:29-44,235-323``) on the B200 path.

Same flags and the same ``predict_example(...)`` signature.  The reference reads NIfTI volumes through
MONAI/nibabel (absent here, and data loading is outside the accelerated path): this entry point accepts
either a ``.npy``/``.pt`` stack of volumes at ``--dataset_path`` or, by default, generates the synthetic
cube volumes of ``generate_artificial_dataset.py`` in memory.  Checkpoints are the reference's
PyTorch-Lightning ``.ckpt`` files (``state_dict`` + ``hyper_parameters``).  Detections are written as
``predictions.json`` (per subject: boxes in fractional boundary coordinates, labels, scores).
"""
from __future__ import annotations

import argparse
import json
import os
from os.path import exists as pexists
from os.path import join as pjoin

import numpy as np
import torch

from . import synthetic
from .ssd3d import LSSD3D, device


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('-d', '--dataset_path', type=str, default=r'../data/artificial_dataset',
                        help="path to a .npy/.pt stack of volumes (N,C,D,H,W); synthetic volumes if it does not exist")
    parser.add_argument('-dn', '--dataset_name', type=str, help="name of dataset to use", default=None)
    parser.add_argument('-m', '--model_path', type=str, help="path to model", default=r'model_final.onnx')
    parser.add_argument('-mn', '--model_name', type=str, help="wandb model name", default=None)
    parser.add_argument('-p', '--percentage', type=float, help="percentage of the dataset to predict on", default=1.)
    parser.add_argument('-su', '--subject', type=str, default=None,
                        help="if prediction has to be done on 1 subject only, specify its id")
    parser.add_argument('-c', '--n_classes', type=int, help="number of classes in dataset", default=1)
    parser.add_argument('-nw', '--num_workers', type=int, default=8, help="number of workers for the dataset")
    parser.add_argument('-ps', '--predict_subset', type=str, help="subset to predict on",
                        choices=['train', 'validation', 'test', 'all'], default=r'train')
    parser.add_argument('-sc', '--min_score', type=float, default=0.5,
                        help="minimum score for a candidate box to be considered as positive")
    parser.add_argument('-k', '--top_k', type=int, default=100,
                        help="if there are a lot of resulting detection across all classes, keep only the top 'k'")
    parser.add_argument('-o', '--output_dir', type=str, help="path to output", default=r"../data/predictions/")
    parser.add_argument('-si', '--save_images', type=int, default=1,
                        help="whether to save the predictions (JSON here; NIfTI in the reference)")
    parser.add_argument('-bs', '--batch_size', type=int, default=8, help="volumes per forward (reference: 1)")
    parser.add_argument('-n', '--n_subjects', type=int, default=16, help="synthetic subjects when no dataset file")
    return parser


def load_volumes(dataset_path, n_subjects, channels, size, percentage=1.0):
    if dataset_path and os.path.isfile(dataset_path):
        vols = np.load(dataset_path) if dataset_path.endswith(".npy") else torch.load(dataset_path).numpy()
        vols = vols.astype(np.float32)
    else:
        vols = synthetic.make_batch(n_subjects, channels, size)
    n = max(1, int(round(vols.shape[0] * percentage)))
    return vols[:n]


def predict_example(model_path, output_dir, dataset_path, dataset_name, n_classes=1, subject=None, percentage=1.,
                    predict_subset="train", min_score=0.5, top_k=10, num_workers=8, save_images=True, model_name=None,
                    batch_size=8, n_subjects=16, model=None):
    torch.manual_seed(970205)
    output_dir = output_dir if dataset_name is None else pjoin(output_dir, dataset_name)
    output_dir = output_dir if model_name is None else pjoin(output_dir, model_name)
    output_dir = pjoin(output_dir, f"{predict_subset}_set", f"min_score_{min_score}")
    if not pexists(output_dir):
        os.makedirs(output_dir)

    if model is None:
        model = LSSD3D.load_from_checkpoint(model_path, min_score=min_score)
    model = model.to(device).eval()
    model.top_k = top_k
    model.min_score = min_score

    vols = load_volumes(dataset_path, n_subjects, model.input_channels, tuple(model.input_size), percentage)
    if subject is not None:
        vols = vols[int(subject):int(subject) + 1]
    det_locs, det_labels, det_scores = [], [], []
    with torch.no_grad():
        for s in range(0, vols.shape[0], batch_size):
            batch = {"img": torch.from_numpy(vols[s:s + batch_size]).pin_memory()}
            locs, labels, scores = model.predict_step(batch, s // batch_size)
            det_locs += locs
            det_labels += labels
            det_scores += scores
    results = {str(i): {"boxes": det_locs[i].cpu().tolist(), "labels": det_labels[i].cpu().tolist(),
                        "scores": det_scores[i].cpu().tolist()} for i in range(len(det_locs))}
    if save_images and output_dir is not None:
        with open(pjoin(output_dir, "predictions.json"), "w") as f:
            json.dump(results, f)
    return results


def main(argv=None):
    args = build_parser().parse_args(argv)
    subsets = ["train", "validation", "test"] if args.predict_subset == 'all' else [args.predict_subset]
    for psubset in subsets:
        predict_example(model_path=args.model_path, output_dir=args.output_dir, dataset_path=args.dataset_path,
                        dataset_name=args.dataset_name, n_classes=args.n_classes, subject=args.subject,
                        percentage=args.percentage, predict_subset=psubset, min_score=args.min_score,
                        top_k=args.top_k, num_workers=args.num_workers, save_images=args.save_images,
                        model_name=args.model_name, batch_size=args.batch_size, n_subjects=args.n_subjects)


if __name__ == "__main__":
    main()
