#!/usr/bin/env python
"""Forward throughput of encoder='cnn' (torchvision ResNet-50 + KTD) on one B200 — the literal stage-1 shape
(configs/config_stage1.yaml: 128 images per GPU, T = 1) and the BASELINE shape (8 clips x T = 16 = 128 frames: the same
engine work).  Not part of the driver's bench contract (bench.py keeps the headline 'ste' metric).

    python scripts/bench_cnn.py [--steps 20] [--warmup 5] [--no-cpu-baseline]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GFLOP_PER_FRAME = 8.174          # 2 x MACs of torchvision resnet50 minus fc (torch.utils.flop_counter), decoder ~0.01


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--images", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_cnn.py: no CUDA device (the maed_b200 path has no CPU fallback)")
    import bench
    from maed_b200 import build, ops
    from maed_b200.models import MAED
    from oracle import maed_oracle as O
    from maed_b200 import synth
    build.build()
    dev = torch.device("cuda", 0)
    m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
    synth.fill_module_(m, 0)
    sd = {k: v.detach().clone() for k, v in list(m.named_parameters()) + list(m.named_buffers())}
    m = m.to(dev).eval()
    n = args.images
    xs = [synth.synth_frames(n, 1, 500 + i).to(dev) for i in range(4)]       # 4 x 77 MB rotating batches (> L2)
    for i in range(max(args.warmup, 3)):
        m(xs[i % 4])
    torch.cuda.synchronize(dev)
    sampler = bench.ClockSampler(0)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        out = m(xs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    launches = (ops.launch_count() - l0) // args.steps
    clocks = sampler.stop()
    peaks, src = bench.load_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    tflops = n * GFLOP_PER_FRAME / 1000.0 / (ms / 1000.0)
    line = {"metric": "images/sec (224x224, %d images/gpu, T=1), MAED cnn (ResNet-50) + ktd forward" % n, "value": n / (ms / 1000.0),
            "unit": "images/s", "clips_per_s_T16": n / 16.0 / (ms / 1000.0), "n_gpus": 1, "steps": args.steps, "ms_per_step": ms,
            "dtype": "f16 hi/lo split operands (3 tcgen05 MMAs per K step); BatchNorm folded in fp32", "data": "synthetic",
            "config": {"workload": "configs/config_stage1.yaml shape: 128 images per GPU, encoder='cnn', decoder='ktd', eval"},
            "gpu_launches_per_step": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": tflops, "peak": peak, "frac": tflops / peak,
                         "peak_source": src, "note": "whole step, algorithmic FLOPs (8.17 GFLOP per frame)"}}
    if not args.no_cpu_baseline:
        x = synth.synth_frames(2, 1, 1)
        torch.set_num_threads(min(16, os.cpu_count() or 1))
        with torch.no_grad():
            O.maed_forward(x, sd, "vanilla", "ktd", encoder="cnn")
            t0 = time.perf_counter()
            for _ in range(3):
                ref = O.maed_forward(x, sd, "vanilla", "ktd", encoder="cnn")
            dt = (time.perf_counter() - t0) / 3
        line["cpu_baseline"] = {"value": 2 / dt, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "3 x (2 images) forward, oracle/maed_oracle.py cnn_encoder"}
        got = m(x.to(dev))
        line["parity_theta_rel_err"] = float(((got["theta"].cpu() - ref["theta"]).norm() / ref["theta"].norm()).item())
    print(json.dumps(line))


if __name__ == "__main__":
    main()
