#!/usr/bin/env python
"""Per-kernel timing of the training path's kernels at the bench shapes (8 clips x T=16: 25 216 token rows, 128 frames) with
CUDA events: algorithmic bytes / FLOPs per launch against the measured peaks (MEASURED_PEAKS.json).  For round 2: which backward
kernels are worth optimising first.  Every call goes through the C ABI like the tests do.

    python scripts/bench_train_kernels.py [--reps 10]
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench_train_kernels.py: no CUDA device")
    import bench
    from maed_b200 import _lib, build, ops
    build.build()
    lib = _lib.load()
    peaks, _ = bench.load_peaks()
    hbm, tf = float(peaks.get("hbm_gbs", 6550.0)), float(peaks.get("bf16_tflops_sustained", 1400.0))
    dev = "cuda"
    st = _lib.stream_ptr
    rows, Cm, heads, BT, N, T, ntok = 8 * 16 * 197, 768, 12, 128, 8, 16, 197
    out = []

    def rec(name, ms, gbytes=None, gflop=None, note=""):
        r = {"kernel": name, "ms": round(ms, 4), "note": note}
        if gbytes is not None:
            r["GB/s"] = round(gbytes / (ms / 1e3), 1)
            r["frac_hbm"] = round(r["GB/s"] / hbm, 3)
        if gflop is not None:
            r["TFLOP/s"] = round(gflop / ms, 2)
            r["frac_tensor"] = round(r["TFLOP/s"] / tf, 4)
        out.append(r)
        print(json.dumps(r))

    # ---- LayerNorm backward [rows, 768]
    x, dy = torch.randn(rows, Cm, device=dev), torch.randn(rows, Cm, device=dev)
    gamma = torch.ones(Cm, device=dev)
    pr = lib.maed_bwd_layernorm_partial_rows()
    partial, scratch = torch.empty(pr, 2 * Cm, device=dev), torch.empty(lib.maed_bwd_colsum_chunks() * 2 * Cm, device=dev)
    dx, dg, db = torch.empty_like(x), torch.empty(Cm, device=dev), torch.empty(Cm, device=dev)
    ms = timed(lambda: _lib.call("maed_bwd_layernorm", _lib.ptr(dy), C.c_longlong(Cm), _lib.ptr(x), C.c_longlong(Cm), _lib.ptr(gamma),
                                 rows, Cm, C.c_float(1e-6), None, _lib.ptr(dx), C.c_longlong(Cm), _lib.ptr(partial),
                                 _lib.ptr(scratch), _lib.ptr(dg), _lib.ptr(db), st()), args.reps)
    rec("layernorm_bwd", ms, gbytes=3 * rows * Cm * 4 / 1e9)
    # ---- GELU backward [rows, 3072]
    pre, d = torch.randn(rows, 4 * Cm, device=dev), torch.randn(rows, 4 * Cm, device=dev)
    o = torch.empty(2, rows, 4 * Cm, dtype=torch.float16, device=dev)
    ms = timed(lambda: _lib.call("maed_bwd_gelu", _lib.ptr(d), _lib.ptr(pre), C.c_longlong(pre.numel()), _lib.ptr(o),
                                 C.c_longlong(o[0].numel()), st()), args.reps)
    rec("gelu_bwd", ms, gbytes=rows * 4 * Cm * 12 / 1e9)
    del pre, d, o
    # ---- transpose_planes [rows, 3072] (the weight-gradient operands)
    p = ops.split(torch.randn(rows, 4 * Cm, device=dev))
    ld = (rows + 7) // 8 * 8
    pt = torch.empty(2, 4 * Cm, ld, dtype=torch.float16, device=dev)
    ms = timed(lambda: _lib.call("maed_bwd_transpose_planes", _lib.ptr(p), C.c_longlong(p[0].numel()), rows, 4 * Cm, 4 * Cm, _lib.ptr(pt),
                                 C.c_longlong(pt[0].numel()), ld, st()), args.reps)
    rec("transpose_planes 25216x3072", ms, gbytes=rows * 4 * Cm * 8 / 1e9)
    # ---- split-K weight gradient: fc1 (3072 x 768 over 25216 rows)
    a = ops.split(torch.randn(4 * Cm, ld, device=dev) * 0.05)
    b = ops.split(torch.randn(Cm, ld, device=dev) * 0.05)
    slabs = torch.empty(lib.maed_bwd_wgrad_slab_floats(4 * Cm, Cm, rows), device=dev)
    D = torch.empty(4 * Cm, Cm, device=dev)
    ms = timed(lambda: _lib.call("maed_bwd_wgrad_splitk", _lib.ptr(a), C.c_longlong(a[0].numel()), ld, _lib.ptr(b), C.c_longlong(b[0].numel()),
                                 ld, 4 * Cm, Cm, rows, 3, C.c_float(1.0), 0, _lib.ptr(slabs), _lib.ptr(D), Cm, st()), args.reps)
    rec("gemm_splitk (fc1 wgrad)", ms, gflop=2.0 * rows * 4 * Cm * Cm / 1e9, note="algorithmic FLOPs; 3 MMAs issued per K step")
    del a, b, slabs, p, pt
    # ---- attention backward
    qkv = ops.split(torch.randn(rows, 3 * heads * 64, device=dev))
    d_out = torch.randn(rows, heads * 64, device=dev)
    d_qkv = torch.empty(rows, 3 * heads * 64, device=dev)
    for kind, name, flop in ((0, "attn_spatial_bwd", 5 * 2.0 * BT * heads * ntok * ntok * 64), (1, "attn_temporal_bwd", 5 * 2.0 * N * heads * ntok * T * T * 64)):
        ms = timed(lambda: _lib.call("maed_bwd_attention", kind, _lib.ptr(qkv), C.c_longlong(qkv[0].numel()), _lib.ptr(d_out), N, T, ntok,
                                     heads, C.c_float(0.125), 0, _lib.ptr(d_qkv), None, st()), args.reps)
        rec(name, ms, gbytes=(rows * 2304 * 8 + rows * 768 * 4) / 1e9, gflop=flop / 1e9, note="CUDA-core fp32 (first version)")
    del qkv, d_out, d_qkv
    # ---- GroupNorm backward: stage-0 output [128, 3136, 256]
    n, HW, Cc = BT, 3136, 256
    x, dy = torch.randn(n, HW, Cc, device=dev), torch.randn(n, HW, Cc, device=dev)
    gamma = torch.ones(Cc, device=dev)
    stats = torch.empty(n * 64, dtype=torch.float64, device=dev)
    red, dgb = torch.empty(n * (64 + 32 * Cc), device=dev), torch.empty(n, 2, Cc, device=dev)
    dxp = torch.empty(2, n, HW, Cc, dtype=torch.float16, device=dev)
    ms = timed(lambda: _lib.call("maed_bwd_groupnorm", _lib.ptr(dy), _lib.ptr(x), n, HW, Cc, _lib.ptr(gamma), C.c_float(1e-5), _lib.ptr(stats),
                                 _lib.ptr(red), _lib.ptr(dgb), _lib.ptr(dxp), C.c_longlong(dxp[0].numel()), None, 0, st()), args.reps)
    rec("groupnorm_bwd (+ stats) 128x3136x256", ms, gbytes=n * HW * Cc * (4 + 4 + 4 + 4 + 4) / 1e9)
    # ---- BatchNorm train forward + backward on the same map ('cnn' layer1 output)
    M = n * HW
    beta = torch.zeros(Cc, device=dev)
    rm, rv = torch.zeros(Cc, device=dev), torch.ones(Cc, device=dev)
    mean, rstd, dg, db = [torch.zeros(Cc, device=dev) for _ in range(4)]
    y = torch.empty(2, M, Cc, dtype=torch.float16, device=dev)
    scr = torch.zeros(lib.maed_bwd_batchnorm_scratch_doubles(M, Cc), dtype=torch.float64, device=dev)
    xf, dyf = x.reshape(M, Cc), dy.reshape(M, Cc)
    ms = timed(lambda: _lib.call("maed_bwd_batchnorm", _lib.ptr(xf), C.c_longlong(M), Cc, _lib.ptr(gamma), _lib.ptr(beta), C.c_float(1e-5),
                                 C.c_float(0.1), _lib.ptr(rm), _lib.ptr(rv), 1, None, C.c_longlong(0), _lib.ptr(y), C.c_longlong(M * Cc),
                                 _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dyf), C.c_float(1.0), _lib.ptr(dg), _lib.ptr(db), _lib.ptr(dxp),
                                 C.c_longlong(M * Cc), _lib.ptr(scr), st()), args.reps)
    rec("batchnorm train fwd + bwd 401408x256", ms, gbytes=M * Cc * (4 + 4 + 4 + 8 + 4 + 4) / 1e9)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "train_kernels.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
