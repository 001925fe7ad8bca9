#!/usr/bin/env python
"""Second baseline of SURVEY.md 8c/8d: the reference's own torch ops run by STOCK PyTorch (cuDNN / ATen) on the
same B200 -- what a user gets today by moving the reference model to the GPU -- beside this library, on
BASELINE.json configs[1] (2ch 128^3, batch 8).  Not a product path: it times torch, through the oracle port's
forward (same F.conv3d / F.batch_norm / F.relu calls as mobilenet.py:26-49, ssd3d.py:131-167) and a restatement
of the reference's detect loop (ssd3d.py:376-453: softmax, filter, sort, n x n IoU, Python greedy loop with a
host sync per kept box) on CUDA tensors.

    python scripts/bench_torch_gpu.py [--json out.json]
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import synthetic  # noqa: E402
from mslesions3d_b200.ssd3d import LSSD3D  # noqa: E402
from oracle import ssd3d_oracle as O  # noqa: E402  (baseline leg only: the thing timed here is stock torch)

ap = argparse.ArgumentParser()
ap.add_argument("--json", default="")
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
CH, SIZE, BATCH = 2, (128, 128, 128), 8
MIN_SCORE, MAX_OVERLAP, TOP_K = 0.5, 0.5, 100


def torch_detect(locs, scores, priors):
    """ssd3d.py:344-460 with torch ops on the tensors' device."""
    out = []
    probs = F.softmax(scores, dim=2)
    for i in range(locs.shape[0]):
        boxes = O.cxcycz_to_xyz(O.gcxgcygcz_to_cxcycz(locs[i], priors))
        ib, il, isc = [], [], []
        for c in range(1, probs.shape[2]):
            s = probs[i][:, c]
            above = s > MIN_SCORE
            if int(above.sum()) == 0:
                continue
            s, b = s[above], boxes[above]
            s, order = s.sort(dim=0, descending=True)
            b = b[order][:10 * TOP_K]
            s = s[:10 * TOP_K]
            overlap = O.find_jaccard_overlap3d(b, b)
            suppress = torch.zeros(b.shape[0], dtype=torch.uint8, device=b.device)
            for k in range(b.shape[0]):
                if suppress[k] == 1:                       # host sync, as in the reference
                    continue
                suppress = torch.max(suppress, (overlap[k] > MAX_OVERLAP).to(torch.uint8))
                suppress[k] = 0
            keep = suppress == 0
            ib.append(b[keep]); il.append(torch.full((int(keep.sum()),), c, device=b.device)); isc.append(s[keep])
        if not ib:
            out.append(None)
            continue
        ib, il, isc = torch.cat(ib), torch.cat(il), torch.cat(isc)
        if isc.shape[0] > TOP_K:
            isc, order = isc.sort(dim=0, descending=True)
            ib, il, isc = ib[order][:TOP_K], il[order][:TOP_K], isc[:TOP_K]
        out.append((ib, il, isc))
    return out


def timed(fn, steps):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


sd_cpu = synthetic.random_state_dict(CH, seed=0)
sd = {k: v.to(dev) for k, v in sd_cpu.items()}
x = torch.from_numpy(synthetic.make_batch(BATCH, CH, SIZE)).to(dev)
priors = O.prior_boxes_fast(SIZE, None, in_channels=CH).to(dev)
rows = {}
with torch.no_grad():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rows["torch_fp32_forward_ms"] = timed(lambda: O.forward(sd, x), args.steps)
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    rows["torch_tf32_forward_ms"] = timed(lambda: O.forward(sd, x), args.steps)
    torch.backends.cudnn.benchmark = True

    def fwd_bf16():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.forward(sd, x)
    rows["torch_bf16_autocast_forward_ms"] = timed(fwd_bf16, args.steps)
    xcl = x.contiguous(memory_format=torch.channels_last_3d)

    def fwd_bf16_cl():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.forward(sd, xcl)
    rows["torch_bf16_channels_last_forward_ms"] = timed(fwd_bf16_cl, args.steps)
    locs, scores = O.forward(sd, x)
    rows["torch_detect_ms"] = timed(lambda: torch_detect(locs, scores, priors), 2)
    best_fwd = min(v for k, v in rows.items() if k.endswith("forward_ms"))
    rows["torch_gpu_volumes_per_s"] = BATCH / ((best_fwd + rows["torch_detect_ms"]) / 1e3)
    rows["torch_gpu_forward_only_volumes_per_s"] = BATCH / (best_fwd / 1e3)

    # this library, same inputs, one lone predict_step at a time (no pipelining) and the streaming API
    model = LSSD3D(n_classes=2, input_channels=CH, input_size=SIZE, min_score=MIN_SCORE, max_overlap=MAX_OVERLAP,
                   top_k=TOP_K)
    model.load_state_dict(sd_cpu)
    model = model.to(dev).eval()
    xb = x.to(torch.bfloat16)
    rows["ours_lone_predict_step_ms"] = timed(lambda: model.predict_step({"img": xb}, 0), 50)
    rows["ours_forward_only_ms"] = timed(lambda: model(xb), 50)

    def stream():
        for _ in model.predict_batches({"img": xb} for _ in range(60)):
            pass
    rows["ours_pipelined_ms_per_batch"] = timed(stream, 2) / 60
    rows["ours_volumes_per_s"] = BATCH / (rows["ours_pipelined_ms_per_batch"] / 1e3)
    # same detections? (labels / counts; torch-GPU exp differs from expf by ulps, so values are close, not equal)
    ref = torch_detect(locs, scores, priors)
    b, l, s = model.predict_step({"img": xb}, 0)
    rows["detections_ours"] = [int(t.shape[0]) for t in b]
    rows["detections_torch_fp32"] = [0 if r is None else int(r[0].shape[0]) for r in ref]
rows["config"] = "2ch 128^3 batch 8 (BASELINE configs[1]); torch %s, cuDNN %s" % (torch.__version__,
                                                                                 torch.backends.cudnn.version())
print(json.dumps(rows, indent=1))
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
