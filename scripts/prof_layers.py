"""Per-kernel timing of the SSD3D step at the benchmark shapes (2ch 128^3, batch 8): every layer launched alone,
CUDA events on the launch stream, L2 flushed between repetitions.  Prints a table with algorithmic bytes, GB/s and
fraction of the measured HBM peak; `--only NAME` restricts it (used under ncu)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mslesions3d_b200 import ops  # noqa: E402
from mslesions3d_b200.ssd3d import LSSD3D  # noqa: E402
from mslesions3d_b200 import synthetic as O  # noqa: E402  (random weights)

ap = argparse.ArgumentParser()
ap.add_argument("--only", default="")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=128)
ap.add_argument("--json", default="")
args = ap.parse_args()

dev = torch.device("cuda")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
N, S = args.batch, args.size
model = LSSD3D(n_classes=2, input_channels=2, input_size=(S, S, S))
model.load_state_dict(O.random_state_dict(2, seed=0))
model = model.to(dev).eval()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flag = torch.zeros(1, dtype=torch.int32, device=dev)


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


rows = []


def run(name, fn, nbytes, flops=0):
    if args.only and args.only not in name:
        return
    us = timeit(fn, args.reps)
    gbs = nbytes / us / 1e3
    rows.append(dict(kernel=name, us=us, MB=nbytes / 1e6, GBs=gbs, frac_hbm=gbs / peak, GFLOP=flops / 1e9,
                     TFLOPs=flops / us / 1e6))
    print("%-34s %8.1f us  %8.1f MB  %7.0f GB/s  %5.1f%% HBM  %7.2f GFLOP %7.1f TF/s" % (
        name, us, nbytes / 1e6, gbs, 100 * gbs / peak, flops / 1e9, flops / us / 1e6), flush=True)


x = torch.randn(N, 2, S, S, S, device=dev).to(torch.bfloat16)
feats = model.base.features
with torch.no_grad():
    stem = feats[0]
    w, sc, sh = stem._pack()
    cur = ops.stem_conv_bn_relu(x, w, sc, sh, 2)
    vout = cur.numel() // 32
    run("stem_tc 2->32", lambda: ops.stem_conv_bn_relu(x, w, sc, sh, 2), x.numel() * 2 + cur.numel() * 2,
        2 * 54 * 32 * vout)
    fmaps = {}
    for i in range(1, len(feats)):
        blk = feats[i]
        wd, s1, b1, wp, s2, b2 = blk._pack()
        st = blk.conv1.stride[0]
        c = cur.shape[1]
        inp = cur
        mid = ops.dwconv3d_bn_relu(inp, wd, s1, b1, st)
        run("f%d dw C=%d s%d out %d^3" % (i, c, st, mid.shape[2]),
            lambda inp=inp, wd=wd, s1=s1, b1=b1, st=st: ops.dwconv3d_bn_relu(inp, wd, s1, b1, st),
            inp.numel() * 2 + mid.numel() * 2 + 54 * c, 54 * mid.numel())
        out = ops.pwconv_bn_relu(mid, wp, s2, b2, flag)
        co = out.shape[1]
        run("f%d pw %d->%d M=%d" % (i, c, co, mid.numel() // c),
            lambda mid=mid, wp=wp, s2=s2, b2=b2: ops.pwconv_bn_relu(mid, wp, s2, b2, flag),
            mid.numel() * 2 + out.numel() * 2 + 2 * c * co, 2 * (mid.numel() // c) * c * co)
        cur = out
        if i in (3, 5, 7):
            fmaps[i] = cur
    P = model.priors_cxcycz.shape[0]
    locs = torch.empty(N, P, 6, device=dev)
    scores = torch.empty(N, P, 2, device=dev)
    packed = model.pred_convs._pack()
    off = 0
    for hi, k in enumerate((3, 5, 7)):
        f = fmaps[k]
        wpk, bpk = packed[hi]
        vox = f.numel() // f.shape[1]
        for algo in (2, 1, 3, 0):
            if algo == 3 and vox < 256:
                continue
            run("head f%d C=%d %d^3 algo%d" % (k, f.shape[1], f.shape[2], algo),
                lambda f=f, wpk=wpk, bpk=bpk, off=off, algo=algo: ops.head_conv(f, wpk, bpk, locs, scores, 2, 2, off,
                                                                              flag, algo=algo),
                f.numel() * 2 + vox * 16 * 4 + wpk.numel() * 2, 2 * 27 * f.shape[1] * 16 * vox)
        off += vox // N * 2
    pri = model.priors_cxcycz
    run("detect_objects (4 kernels)", lambda: ops.detect_objects_padded(locs, scores, pri, 0.5, 0.5, 100),
        N * P * (24 + 8 + 24 + 24))
if args.json:
    json.dump(rows, open(args.json, "w"), indent=1)
