#!/usr/bin/env python
"""Benchmark of the SSD3D hot path: volumes/sec for forward + decode + NMS (= ``LSSD3D.predict_step``).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): SSD3D-MobileNet inference on synthetic 2-channel 128^3 cube-lesion
volumes (generate_artificial_dataset.py distribution), batch 8 per GPU, bf16 activations, random-init
weights with randomised BN statistics.  One step = one predict_step over one batch.  With N > 1 every rank
runs the same per-GPU batch (independent volumes, no data-path collective): weak scaling.

The single JSON line printed by rank 0 carries
  value        whole-job volumes/s with the input batches already resident in HBM (device-timed, max over ranks)
  e2e          the same through the public streaming API (LSSD3D.predict_batches, the analogue of the
               reference's Trainer.predict loop) from PINNED HOST buffers: H2D of every batch and D2H of its
               detections inside the timed region; the copy of batch i+1 overlaps the compute of batch i
  roofline     the dominant kernel timed alone with CUDA events, algorithmic bytes / time vs the measured
               HBM peak of MEASURED_PEAKS.json
  cpu_baseline the CPU oracle (a restatement of the reference's torch-CPU path) on a bounded sample of the
               same workload, best thread count
  extras       (N = 1) a secondary number outside the headline metric: greedy 3-D NMS over 2.5 M score-sorted
               candidates without truncation (configs[3]/[4]); a failure there is recorded, never fatal
``--impl reference`` times that CPU path alone (rank 0 only) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "SSD3D volumes/sec (fwd+decode+NMS)"
UNIT = "volumes/s"
WORKLOAD = "SSD3D-MobileNet inference, 2ch 128^3 synthetic cube-lesion volumes, batch 8 per GPU, bf16"
CHANNELS, SIZE, BATCH = 2, (128, 128, 128), 8
MIN_SCORE, MAX_OVERLAP, TOP_K = 0.5, 0.5, 100
N_ROTATE = 4   # distinct resident input batches: 4 x 67 MB > 126 MB of L2


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(name):
    """dram read + write bytes per launch from a committed ncu --set full summary (profiles/), or None."""
    try:
        rows = json.load(open(os.path.join(ROOT, "profiles", name)))
        vals = []
        for r in rows:
            rd, wr = r["dram__bytes_read.sum"].split(), r["dram__bytes_write.sum"].split()
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals.append(float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]])
        return float(np.mean(vals)) if vals else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# clocks sampled DURING the timed region
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------------
def make_inputs(n_batches: int, rank: int):
    """Host fp32 batches (n_batches, BATCH, C, D, H, W) of normalised synthetic volumes."""
    from mslesions3d_b200 import synthetic
    out = []
    for b in range(n_batches):
        out.append(synthetic.make_batch(BATCH, CHANNELS, SIZE, first_idx=(rank * n_batches + b) * BATCH))
    return out


def cpu_reference_run(volumes, threads: int, repeats: int = 1):
    """The reference's CPU path (oracle restatement, fp32 torch-CPU ops): forward + detect_objects per volume.
    Returns seconds per volume."""
    from oracle import ssd3d_oracle as O
    torch.set_num_threads(threads)
    sd = O.random_state_dict(CHANNELS, seed=0)
    priors = O.prior_boxes_fast(SIZE, in_channels=CHANNELS)
    x = torch.from_numpy(volumes)
    with torch.no_grad():
        O.forward(sd, x[:1])  # warm-up (thread pool, mkldnn primitives)
        t0 = time.perf_counter()
        for _ in range(repeats):
            for i in range(x.shape[0]):
                locs, scores = O.forward(sd, x[i:i + 1])
                O.detect_objects(locs, scores, priors, MIN_SCORE, MAX_OVERLAP, TOP_K)
        dt = time.perf_counter() - t0
    return dt / (repeats * x.shape[0])


def best_cpu_baseline(sample_volumes):
    """Sweep thread counts (oversubscription can be far slower than 1 thread) and keep the best."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (1, 2, 4, 8, 16, 32, 64, ncpu) if t <= ncpu})
    best = None
    t_start = time.perf_counter()
    for t in cands:
        if time.perf_counter() - t_start > 40:
            break
        spv = cpu_reference_run(sample_volumes, t)
        if best is None or spv < best[0]:
            best = (spv, t)
    return best


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from mslesions3d_b200 import synthetic
    sample = synthetic.make_batch(2, CHANNELS, SIZE)
    spv_probe, threads = best_cpu_baseline(sample[:1])
    torch.set_num_threads(threads)
    for _ in range(args.warmup):
        cpu_reference_run(sample[:1], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(sample, threads)
    dt = time.perf_counter() - t0
    vps = args.steps * sample.shape[0] / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": vps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "step": "2 volumes (bounded sample of the batch-8 workload), batch 1 each, "
                   "fp32 torch-CPU forward + detect_objects"},
        "cpu_baseline": {"value": vps, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "2 volumes per step x %d steps, oracle port of the reference's torch-CPU path, "
                                   "best of thread sweep (host has %d cpus)" % (args.steps, os.cpu_count() or 1)},
        "e2e": {"value": vps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true",
                    help="launch every kernel individually (no CUDA graph): for ncu launch lists, not for numbers")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from mslesions3d_b200.parallel import rank_world, max_over_ranks
    rank, local_rank, world = rank_world()

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from mslesions3d_b200 import _lib, ops
    from mslesions3d_b200.ssd3d import LSSD3D

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    _lib.load()
    if world > 1:
        from mslesions3d_b200.parallel import bind_to_gpu_cpus
        bind_to_gpu_cpus(local_rank)      # pinned staging buffers next to the GPU they feed
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- model + inputs -------------------------------------------------------------------------
    from mslesions3d_b200 import synthetic
    sd = synthetic.random_state_dict(CHANNELS, seed=0)
    model = LSSD3D(n_classes=2, input_channels=CHANNELS, input_size=SIZE, min_score=MIN_SCORE,
                   max_overlap=MAX_OVERLAP, top_k=TOP_K)
    model.load_state_dict(sd)
    model = model.to(dev).eval()
    host_f32 = make_inputs(N_ROTATE, rank)
    host_bf16 = [torch.from_numpy(h).to(torch.bfloat16).pin_memory() for h in host_f32]
    dev_bf16 = [h.to(dev) for h in host_bf16]
    in_bytes = host_bf16[0].numel() * 2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1), dev)    # the job is as slow as its slowest rank

    d2h_bytes = [0]

    def run_resident(steps):
        """`steps` device-resident batches through LSSD3D.predict_batches (results stay on the device)."""
        batches = ({"img": dev_bf16[i % N_ROTATE]} for i in range(steps))
        if args.eager:
            model.use_cuda_graph = False
            with torch.no_grad():
                for b in batches:
                    model.predict_step(b, 0)
            return
        with torch.no_grad():
            for _ in model.predict_batches(batches):
                pass

    def run_e2e(steps):
        """`steps` batches through the public streaming API (LSSD3D.predict_batches): every batch is copied
        from pinned host memory inside the timed region and its detections are read back to the host."""
        batches = ({"img": host_bf16[i % N_ROTATE]} for i in range(steps))
        with torch.no_grad():
            for b, l, s in model.predict_batches(batches, to_host=True):
                assert not b[0].is_cuda
        # per step: padded boxes/labels/scores (top_k rows per volume) + the count/status/flag words
        d2h_bytes[0] = BATCH * TOP_K * (24 + 4 + 8 + 8) + 4 * (BATCH + 2)

    # ---- resident-input throughput (value) --------------------------------------------------------
    # Build (capture) the inference plan of every pipeline slot first: that is one-off setup, not a step.
    depth = int(model.pipeline_depth)
    run_resident(depth)
    run_resident(args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.LAUNCHES[0] = 0
    ms = timed(lambda i: run_resident(args.steps) if i == 0 else None, args.steps)
    launches = ops.LAUNCHES[0]
    clocks = sampler.stop() if rank == 0 else None
    value = world * BATCH * args.steps / (ms / 1000.0)

    # ---- end to end from pinned host memory --------------------------------------------------------
    run_e2e(depth)
    run_e2e(args.warmup)
    ms_e2e = timed(lambda i: run_e2e(args.steps) if i == 0 else None, args.steps)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1000.0)

    # ---- dominant kernel alone: roofline ------------------------------------------------------------
    peak, peak_src = load_peaks()
    stem = model.base.features[0]
    w, scale, shift = stem._pack()
    sd_stride = 2
    def kernel_ms(fn, n_launch):
        """Average duration of n_launch back-to-back launches (one event pair around all of them, so the
        per-launch event/launch overhead is not attributed to the kernel); inputs rotate over > L2 bytes."""
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n_launch):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n_launch

    stem_out = ops.stem_conv_bn_relu(dev_bf16[0], w, scale, shift, sd_stride)
    n_launch = max(20, args.steps)
    k_ms = kernel_ms(lambda i: ops.stem_conv_bn_relu(dev_bf16[i % N_ROTATE], w, scale, shift, sd_stride, out=stem_out),
                     n_launch)
    vox_out = BATCH * (SIZE[0] // 2) * (SIZE[1] // 2) * (SIZE[2] // 2)
    algo_bytes = in_bytes + vox_out * 32 * 2 + 27 * CHANNELS * 32 * 4
    achieved = algo_bytes / (k_ms / 1000.0) / 1e9
    roofline = {"kernel": "stem_tc_kernel<bf16,2> (dense 3x3x3 conv 2->32 + BN + ReLU, tcgen05 implicit GEMM)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("r01_ncu_full_stem_final.json"),
                "traffic_source": "profiles/r01_ncu_full_stem_final.json (ncu --set full, dram read+write per launch; "
                                  "part of the 134 MB output is still in the 126 MB L2 when the kernel ends)",
                "timing": "%d back-to-back launches between one CUDA event pair, inputs rotating over 268 MB" % n_launch,
                "algorithmic_bytes_per_launch": algo_bytes, "kernel_ms": k_ms, "peak_source": peak_src}

    # second HBM-bound kernel the north star names: the first depthwise conv (32 ch, stride 2, 64^3 -> 32^3)
    blk = model.base.features[1]
    wd, s1, b1 = blk._pack()[:3]
    f0 = stem_out
    f0s = [f0, f0.clone()]      # 2 x 134 MB > L2
    dw_ms = kernel_ms(lambda i: ops.dwconv3d_bn_relu(f0s[i % 2], wd, s1, b1, 2), n_launch)
    dw_bytes = f0.numel() * 2 + (f0.numel() // 8) * 2 + 54 * 32
    roofline_dw = {"kernel": "dw_tma_kernel<2,2,4,4,8> (depthwise 3x3x3, 32 ch, stride 2 + BN + ReLU, TMA halo tiles)",
                   "bound": "hbm", "achieved": dw_bytes / (dw_ms / 1000.0) / 1e9, "peak": peak, "unit": "GB/s",
                   "frac": dw_bytes / (dw_ms / 1000.0) / 1e9 / peak, "traffic": ncu_traffic("r01_ncu_full_dw_final.json"),
                   "algorithmic_bytes_per_launch": dw_bytes, "kernel_ms": dw_ms}
    del f0s, f0

    # ---- secondary measurement (rank 0, N = 1 only): any-length greedy NMS, BASELINE configs[3]/[4] ----------
    # Never allowed to break the bench line: any failure is recorded instead.
    extras = None
    if rank == 0 and world == 1:
        try:
            n_nms = 2500000
            gen = torch.Generator(device=dev).manual_seed(n_nms)
            ctr = torch.rand((n_nms, 3), device=dev, generator=gen)
            side = 0.02 + 0.08 * torch.rand((n_nms, 1), device=dev, generator=gen)
            nms_boxes = torch.cat([ctr - side / 2, ctr + side / 2], 1).contiguous()
            keep, kept = ops.nms3d_sorted_chunked(nms_boxes, MAX_OVERLAP, return_count=True)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ops.nms3d_sorted_chunked(nms_boxes, MAX_OVERLAP)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            extras = {"nms_any_length": {"candidates": n_nms, "kept": int(kept.item()), "ms": sorted(ts)[1],
                                         "workload": "greedy 3-D NMS over ALL score-sorted candidates (no 10*top_k "
                                                     "truncation), cubic boxes of side 0.02-0.1 at uniform centres, "
                                                     "threshold 0.5 (SURVEY 8d C5); the reference's n x n IoU matrix "
                                                     "would be 25 TB"}}
            del nms_boxes, keep
        except Exception as exc:      # noqa: BLE001
            extras = {"nms_any_length_error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        spv, threads = best_cpu_baseline(host_f32[0][:1])
        spv = min(spv, cpu_reference_run(host_f32[0][:2], threads))
        cpu = {"value": 1.0 / spv, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "2 volumes of the same 2ch 128^3 workload, batch 1, fp32 torch-CPU oracle port "
                         "(forward + detect_objects), best of a thread sweep on %d host cpus" % (os.cpu_count() or 1)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                       "l2": "inputs rotate over %d resident batches (%d MB) > 126 MB L2" %
                             (N_ROTATE, N_ROTATE * in_bytes // 2 ** 20),
                       "min_score": MIN_SCORE, "max_overlap": MAX_OVERLAP, "top_k": TOP_K, "priors": 9344,
                       "pipeline": "%d batches in flight (LSSD3D.predict_batches: one captured plan per slot, own "
                                   "stream each; a step = one batch through stem + graph replay + result read-back)" % depth},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": d2h_bytes[0], "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "roofline_depthwise": roofline_dw,
            "cpu_baseline": cpu,
            "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
