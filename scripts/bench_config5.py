#!/usr/bin/env python
"""BASELINE.json configs[4] shape on one GPU: 4 clips x T = 32 frames, 'parallel' + KTD, forward (the long-clip stress of the
temporal attention; the model carries a 32-row temp_embed, see DESIGN.md / SURVEY.md 8d config 5).  Not part of the driver's
bench contract.

    python scripts/bench_config5.py [--steps 20] [--warmup 5] [--train]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLIPS, T = 4, 32
GFLOP_PER_CLIP = 32 * 25.709 + 6 * 3 * 0.1549        # per-frame work x 32 + the extra temporal-attention work of T = 32 (T^2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--train", action="store_true", help="train step (fwd + bwd + Adam, MSE on theta) instead of the forward")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_config5.py: no CUDA device (the maed_b200 path has no CPU fallback)")
    import bench
    from maed_b200 import build, ops, train
    from maed_b200.models import MAED
    from maed_b200 import synth
    build.build()
    dev = torch.device("cuda", 0)
    m = MAED("ste", 6, 12, "parallel", "ktd", 1024, temp_frames=32)
    synth.fill_module_(m, 0)
    m = m.to(dev)
    xs = [synth.synth_frames(CLIPS, T, 600 + i).to(dev) for i in range(4)]
    if args.train:
        m.train().enable_training(True)
        opt = train.FusedAdam.for_model(m, lr=1e-4, weight_decay=1e-5)
        target = torch.zeros(CLIPS, T, 85, device=dev)
        target[..., 0] = 1.0

        def step(x):
            opt.zero_grad(set_to_none=True)
            loss = ((m(x)["theta"] - target) ** 2).mean()
            loss.backward()
            opt.step()
    else:
        m.eval()

        def step(x):
            m(x)
    for i in range(max(args.warmup, 3)):
        step(xs[i % 4])
    torch.cuda.synchronize(dev)
    sampler = bench.ClockSampler(0)
    sampler.start()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(xs[i % 4])
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / args.steps
    peaks, src = bench.load_peaks()
    peak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1590.0)))
    tflops = (3.0 if args.train else 1.0) * CLIPS * GFLOP_PER_CLIP / 1000.0 / (ms / 1000.0)
    print(json.dumps({
        "metric": "clips/sec (T=32, 224x224, bs=4/gpu), MAED ste-parallel+ktd %s" % ("train step" if args.train else "forward"),
        "value": CLIPS / (ms / 1000.0), "unit": "clips/s", "n_gpus": 1, "steps": args.steps, "ms_per_step": ms, "data": "synthetic",
        "config": {"workload": "BASELINE configs[4] shape on one GPU: bs=4 T=32 (temp_embed with 32 rows)"},
        "gpu_launches_per_step": int((ops.launch_count() - l0) // args.steps), "clocks": sampler.stop(),
        "roofline": {"bound": "tensor", "unit": "TFLOP/s", "achieved": tflops, "peak": peak, "frac": tflops / peak,
                     "peak_source": src, "note": "whole step, algorithmic FLOPs"}}))


if __name__ == "__main__":
    main()
