#!/usr/bin/env python
"""The other BASELINE.json configs through the public API (LSSD3D.predict_batches, device-resident inputs):
  C1  1ch 64^3 batch 1 (the reference's CPU-runnable case)
  C4  whole-brain 2ch 160x192x160 batch 1: default prediction layers (43 800 priors) and with a layer-0 head
      (2 501 400 priors), realistic (min_score .5, top_k 100) and NMS-stress (min_score 0, top_k 800; and top_k = P/10: no truncation at all) settings
One JSON line per case: volumes/s, ms/step (CUDA events), priors; the CPU oracle timed on the same shape where it
finishes in seconds (C1).  `python scripts/bench_configs.py [--json out.json]`"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mslesions3d_b200 import synthetic  # noqa: E402
from mslesions3d_b200.ssd3d import LSSD3D  # noqa: E402
from oracle import ssd3d_oracle as O  # noqa: E402  (CPU baseline leg only)

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--only", default="", help="substring of the case name")
args = ap.parse_args()
dev = torch.device("cuda")
CASES = [
    dict(name="C1 1ch 64^3 b1", ch=1, size=(64, 64, 64), batch=1, ar=None, min_score=0.5, top_k=100, cpu=True),
    dict(name="C4 whole-brain 2ch 160x192x160 b1, layers 3/5/7", ch=2, size=(160, 192, 160), batch=1, ar=None,
         min_score=0.5, top_k=100, cpu=False),
    dict(name="C4 whole-brain + layer-0 head (2.5M priors), realistic", ch=2, size=(160, 192, 160), batch=1,
         ar={0: [1.], 3: [1.], 5: [1.], 7: [1.]}, min_score=0.5, top_k=100, cpu=False),
    dict(name="C4 whole-brain + layer-0 head (2.5M priors), NMS stress", ch=2, size=(160, 192, 160), batch=1,
         ar={0: [1.], 3: [1.], 5: [1.], 7: [1.]}, min_score=0.0, top_k=800, cpu=False),
    # top_k = P/10 defeats the 10*top_k truncation: ALL 2.5 M candidates are sorted and go through
    # the chunked greedy NMS (ops.detect_objects_long); the reference would need a 25 TB IoU matrix here
    dict(name="C4 whole-brain + layer-0 head (2.5M priors), NMS stress untruncated", ch=2, size=(160, 192, 160),
         batch=1, ar={0: [1.], 3: [1.], 5: [1.], 7: [1.]}, min_score=0.0, top_k=250140, cpu=False, steps=2),
]
rows = []
for c in CASES:
    if args.only and args.only not in c["name"]:
        continue
    steps = c.get("steps", args.steps)
    kw = dict(aspect_ratios=c["ar"]) if c["ar"] else {}
    sd = synthetic.random_state_dict(c["ch"], c["ar"], seed=0)
    model = LSSD3D(n_classes=2, input_channels=c["ch"], input_size=c["size"], min_score=c["min_score"],
                   top_k=c["top_k"], **kw)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    xs = [torch.from_numpy(synthetic.make_batch(c["batch"], c["ch"], c["size"], first_idx=i)).to(dev).to(torch.bfloat16)
          for i in range(3)]
    depth = model.pipeline_depth
    with torch.no_grad():
        for _ in model.predict_batches({"img": xs[i % 3]} for i in range(min(depth + 3, steps))):
            pass
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n_det = 0
        for boxes, labels, scores in model.predict_batches({"img": xs[i % 3]} for i in range(steps)):
            n_det = int(boxes[0].shape[0])
        b.record()
        torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    row = dict(config=c["name"], priors=int(model.priors_cxcycz.shape[0]), ms_per_step=ms,
               volumes_per_s=c["batch"] / (ms / 1e3), min_score=c["min_score"], top_k=c["top_k"], detections=n_det,
               pipeline_depth=depth)
    if c["cpu"]:
        torch.set_num_threads(min(8, os.cpu_count() or 1))
        x = xs[0].float().cpu()
        pri = O.prior_boxes_fast(c["size"], c["ar"], in_channels=c["ch"])
        with torch.no_grad():
            O.forward(sd, x, c["ar"])
            t0 = time.perf_counter()
            for _ in range(3):
                l, s = O.forward(sd, x, c["ar"])
                O.detect_objects(l, s, pri, c["min_score"], 0.5, c["top_k"])
            row["cpu_oracle_volumes_per_s"] = 3 * c["batch"] / (time.perf_counter() - t0)
            row["cpu_threads"] = torch.get_num_threads()
    rows.append(row)
    print(json.dumps(row), flush=True)
    del model, xs
    torch.cuda.empty_cache()
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
