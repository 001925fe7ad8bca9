"""In-graph timing of the SSD3D inference step at the benchmark shape (2ch 128^3, batch 8): segments of the step
are captured as separate CUDA graphs and replayed back to back, so launch gaps are what they are inside the real
plan (no per-launch host overhead, no event overhead between kernels).  Prints us per replay of each segment."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import ops  # noqa: E402
from mslesions3d_b200.ssd3d import LSSD3D  # noqa: E402
from mslesions3d_b200 import synthetic as O  # noqa: E402  (random weights)

dev = torch.device("cuda")
N, S = 8, 128
model = LSSD3D(n_classes=2, input_channels=2, input_size=(S, S, S))
model.load_state_dict(O.random_state_dict(2, seed=0))
model = model.to(dev).eval()
flag = torch.zeros(1, dtype=torch.int32, device=dev)
xs = [torch.randn(N, 2, S, S, S, device=dev).to(torch.bfloat16) for _ in range(4)]
feats = model.base.features
pri = model.priors_cxcycz.to(dev)
model.base.nan_flag(dev)      # blocks share the network's NaN word (no per-block host check)
P = pri.shape[0]
REPS = 50


def graph_time(fn, reps=REPS):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        fn(); fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


with torch.no_grad():
    acts = [None] * len(feats)
    acts[0] = feats[0](xs[0])
    mids = [None] * len(feats)
    for i in range(1, len(feats)):
        wd, s1, b1, wp, s2, b2 = feats[i]._pack()
        mids[i] = ops.dwconv3d_bn_relu(acts[i - 1], wd, s1, b1, feats[i].conv1.stride[0])
        acts[i] = ops.pwconv_bn_relu(mids[i], wp, s2, b2, flag)
    locs = torch.empty(N, P, 6, device=dev)
    scores = torch.empty(N, P, 2, device=dev)
    packed = model.pred_convs._pack()

    res = {}
    k = [0]

    def stem():
        k[0] += 1
        feats[0](xs[k[0] % 4], out=acts[0])
    # eager stem (as the plan runs it), rotating inputs
    for _ in range(3):
        stem()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        stem()
    b.record()
    torch.cuda.synchronize()
    res["stem (eager, back to back)"] = a.elapsed_time(b) * 1e3 / REPS

    def seg(lo, hi):
        def fn():
            cur = acts[lo - 1]
            for i in range(lo, hi + 1):
                cur = feats[i](cur)
        return fn
    for i in range(1, len(feats)):
        wd, s1, b1, wp, s2, b2 = feats[i]._pack()
        st = feats[i].conv1.stride[0]
        res["f%d dw" % i] = graph_time(lambda i=i, wd=wd, s1=s1, b1=b1, st=st: ops.dwconv3d_bn_relu(acts[i - 1], wd, s1, b1, st))
        res["f%d pw" % i] = graph_time(lambda i=i, wp=wp, s2=s2, b2=b2: ops.pwconv_bn_relu(mids[i], wp, s2, b2, flag))
    res["f1..f3 (6 kernels)"] = graph_time(seg(1, 3))
    res["f4..f7 (8 kernels)"] = graph_time(seg(4, 7))
    res["f1..f7 (14 kernels)"] = graph_time(seg(1, 7))
    off = 0
    offs = []
    for hi, kk in enumerate((3, 5, 7)):
        offs.append(off)
        off += acts[kk].numel() // acts[kk].shape[1] // N * 2
    for hi, kk in enumerate((3, 5, 7)):
        res["head f%d" % kk] = graph_time(lambda hi=hi, kk=kk: model.pred_convs.run_head(hi, acts[kk], locs, scores, offs[hi], flag))
    res["heads x3 serial"] = graph_time(lambda: [model.pred_convs.run_head(hi, acts[kk], locs, scores, offs[hi], flag)
                                                  for hi, kk in enumerate((3, 5, 7))])
    res["detect_objects"] = graph_time(lambda: ops.detect_objects_padded(locs, scores, pri, 0.5, 0.5, 100))
    # the real plan
    plan = model._plan_for(xs[0])
    plan.launch(xs[0])
    torch.cuda.synchronize()
    rl, rs = plan.locs.clone(), plan.scores.clone()
    res["detect_objects on the plan's real outputs"] = graph_time(lambda: ops.detect_objects_padded(rl, rs, pri, 0.5, 0.5, 100))
    probs = torch.softmax(rs, 2)[:, :, 1]
    print("candidates above 0.5 per image:", (probs > 0.5).sum(1).tolist())

    head_streams = [torch.cuda.Stream() for _ in range(3)]

    def net_forked():
        main = torch.cuda.current_stream()
        cur = acts[0]
        for i in range(1, len(feats)):
            cur = feats[i](cur)
            if i in (3, 5, 7):
                hi = (3, 5, 7).index(i)
                hs = head_streams[hi]
                hs.wait_stream(main)
                with torch.cuda.stream(hs):
                    model.pred_convs.run_head(hi, cur, locs, scores, offs[hi], flag)
        for hs in head_streams:
            main.wait_stream(hs)
    res["f1..f7 + forked heads"] = graph_time(net_forked)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(REPS):
        plan.graph.replay()
    b.record()
    torch.cuda.synchronize()
    res["plan graph (everything but the stem)"] = a.elapsed_time(b) * 1e3 / REPS
    a.record()
    for i in range(REPS):
        plan.stem(xs[i % 4], out=plan.stem_out)
        plan.graph.replay()
    b.record()
    torch.cuda.synchronize()
    res["stem + plan graph, no readback"] = a.elapsed_time(b) * 1e3 / REPS

for kk, v in res.items():
    print("%-40s %8.1f us" % (kk, v))
if len(sys.argv) > 1:
    json.dump(res, open(sys.argv[1], "w"), indent=1)
