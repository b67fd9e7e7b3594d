#!/usr/bin/env python
"""Micro-benchmark behind a design decision: how fast does the plain tcgen05 GEMM stream when its epilogue adds a
residual given as fp16 planes and writes fp16 planes (conv3 + norm + shortcut + ReLU of a bottleneck as ONE plain GEMM launch) at
the memory-bound backbone shapes?  Compared with the fused conv+GroupNorm kernel's time for the same layer.

    python scripts/bench_gemm_epilogue.py
"""
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps=10):
    for _ in range(3):
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
    from maed_b200 import _lib, build, ops
    build.build()
    dev = "cuda"
    st = _lib.stream_ptr
    # (name, M, N, K, residual?)  — stage-0 / 1 / 2 shortcut layers at 128 frames
    shapes = [("stage0.c3 64->256 +res", 401408, 256, 64, True), ("stage0.ds 64->256", 401408, 256, 64, False),
              ("stage1.c3 128->512 +res", 100352, 512, 128, True), ("stage2.c3 256->1024 +res", 25088, 1024, 256, True),
              ("stage2.c1 1024->256", 25088, 256, 1024, False)]
    for name, M, N, K, has_res in shapes:
        a = ops.split(torch.randn(M, K, device=dev))
        b = ops.split(torch.randn(N, K, device=dev) * 0.05)
        bias = torch.zeros(N, device=dev)
        res = ops.split(torch.randn(M, N, device=dev)) if has_res else None
        out = torch.empty(2, M, N, dtype=torch.float16, device=dev)
        outf = torch.empty(M, N, device=dev)

        def planes():
            _lib.call("maed_op_gemm_bottleneck", _lib.ptr(a), C.c_longlong(a[0].numel()), K, _lib.ptr(b), C.c_longlong(b[0].numel()), K,
                      M, N, K, 3, _lib.ptr(bias), _lib.ptr(res), C.c_longlong(res[0].numel() if has_res else 0), 2, ops.OUT_F16_SPLIT,
                      _lib.ptr(out), C.c_longlong(out[0].numel()), N, st())

        def f32():
            _lib.call("maed_op_gemm", _lib.ptr(a), C.c_longlong(a[0].numel()), K, _lib.ptr(b), C.c_longlong(b[0].numel()), K,
                      M, N, K, 3, _lib.ptr(bias), None, 0, ops.OUT_F32, _lib.ptr(outf), C.c_longlong(0), N, 0, st())

        ms_p, ms_f = timed(planes), timed(f32)
        by_p = (M * K * 4 + M * N * 4 + (M * N * 4 if has_res else 0)) / 1e9
        by_f = (M * K * 4 + M * N * 4) / 1e9
        print(json.dumps({"layer": name, "M": M, "N": N, "K": K, "planes_epilogue_us": round(1e3 * ms_p, 1),
                          "planes_GBps": round(by_p / ms_p * 1e3, 0), "f32_out_us": round(1e3 * ms_f, 1),
                          "f32_GBps": round(by_f / ms_f * 1e3, 0)}))
        del a, b, res, out, outf
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
