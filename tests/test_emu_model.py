"""Whole-engine runs on the CUDA-on-CPU test build (tests/emu/): the host orchestration of engine.cu / train.cu and every
CUDA-core kernel are the REAL sources (compiled by g++ against the shim); the tcgen05 / TMA kernels are replaced by CPU
restatements of their contracts (tests/emu/tc_stubs.cpp).  CPU-only, part of the default `-m "not gpu"` suite.

  * forward: the emulated engine reproduces the golden outputs of the unmodified reference — this pins the emulator itself
    (the same files gate the product library on the GPU, tests/test_model_gpu.py);
  * training: tape forward == inference forward, and ALL parameter gradients of train_backward match the gradient digests of
    the unmodified reference (tests/golden/grads_*.npz) for parallel / series / vanilla — i.e. the buffer plumbing, operand
    layouts (transposed / flipped weight packs, im2col, dilation), the backward walk and every CUDA-core backward kernel
    are checked without a GPU.  What this cannot check: the tensor-core kernels themselves (split-K wgrad kernel, GEMM /
    conv kernels on the new operand layouts), races, and launch limits other than those the shim enforces.
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, build_model, load_golden, rel_err
from oracle import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


@pytest.fixture(scope="module")
def harness():
    import harness as h
    h.load()
    return h


@pytest.mark.parametrize("name", ["parallel_ktd_T1", "series_iterative", "coupling_ktd", "temporal_ktd"])
def test_emulated_forward_matches_reference_golden(harness, name):
    g, meta = load_golden(name)
    em = harness.EmuModel(build_model(meta))
    o = em.forward(synth.synth_frames(meta["N"], meta["T"], meta["seed"]))
    for k, gk in (("feat", "tap_feat"), ("pose6d", "tap_pose6d"), ("shape", "tap_shape"), ("cam", "tap_cam"),
                  ("rotmat", "out_rotmat"), ("theta", "out_theta"), ("kp_2d", "out_kp_2d")):
        assert rel_err(o[k].reshape(g[gk].shape), g[gk]) < 2e-4, (name, k)


def _digest(g, nsamp=8):
    g = g.detach().double().reshape(-1)
    idx = np.unique(np.linspace(0, g.numel() - 1, nsamp).round().astype(np.int64))
    return g.norm().item(), g[torch.from_numpy(idx)].numpy()


@pytest.mark.parametrize("name", ["grads_vanilla_ktd", "grads_series_ktd", "grads_parallel_ktd", "grads_vanilla_iterative",
                                  "grads_temporal_ktd", "grads_coupling_ktd"])
def test_emulated_training_matches_reference_gradients(harness, name):
    # default suite: the stage-2 mode and the iterative decoder; the other modes (series is covered through the product modules
    # in tests/test_train.py; every case runs on the GPU in `-m gpu`) with MAED_EMU_FULL=1 — keeps the CPU suite at a few minutes
    if name not in ("grads_parallel_ktd", "grads_vanilla_iterative") and not os.environ.get("MAED_EMU_FULL"):
        pytest.skip("long on the emulator: set MAED_EMU_FULL=1")
    from maed_b200.models import MAED
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, T, seed = [int(v) for v in z["meta"]]
    mode = str(z["mode"])
    m = MAED("ste", 6, 12, mode, str(z["decoder"]), 1024, mean_params=synth.mean_params())
    synth.fill_module_(m, seed)
    em = harness.EmuModel(m)
    x = synth.synth_frames(N, T, seed)
    A, B, Cc = [synth.synth_tensor("grad_probe.%s" % k, (N * T, n), seed) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
    out = em.train_forward(x)
    inf = em.forward(x)
    for k in ("feat", "pose6d", "shape", "cam"):
        assert rel_err(out[k], inf[k]) < 1e-5, k                       # tape forward == inference forward
    loss = (out["pose6d"] * A).sum() + (out["shape"] * B).sum() + (out["cam"] * Cc).sum()
    assert abs(loss.item() - float(z["loss"])) < 1e-3 * max(1.0, abs(float(z["loss"])))
    grads = em.train_backward(A, B, Cc, loss_scale=4096.0)
    names = [str(s) for s in z["names"]]
    assert sorted(names) == sorted(k for k in em.names if k in dict(m.named_parameters()))
    worst = ("", 0.0)
    for k in names:
        g = grads[k]
        assert not torch.isnan(g).any(), "gradient of %s not (fully) written" % k
        norm, samp = _digest(g)
        ref_norm, ref_samp = float(z["g_stats/" + k][0]), z["g_samp/" + k]
        err = abs(norm - ref_norm) / max(ref_norm, 1e-12)
        rms = ref_norm / np.sqrt(g.numel())
        serr = np.abs(samp - ref_samp).max() / max(rms, 1e-20)
        if max(err, serr / 25) > worst[1]:
            worst = (k, max(err, serr / 25))
        # fp32 noise floor: the oracle's own fp32 vs fp64 autograd differ by up to 1.7e-2 per parameter (median 2e-3) on this
        # random weight-standardised network; observed here: worst 5.8e-3
        assert err < 2e-2, "%s: |g| %.6e vs reference %.6e" % (k, norm, ref_norm)
        assert serr < 0.25, "%s: sampled entries off by %.3f rms" % (k, serr)
    print("%s: worst %s %.2e" % (name, worst[0], worst[1]))
    # a different loss scale must give the same gradients (scale enters and leaves exactly once everywhere)
    if name == "grads_vanilla_ktd":
        em.train_forward(x)
        g2 = em.train_backward(A, B, Cc, loss_scale=256.0)
        for k in names:
            assert rel_err(g2[k], grads[k]) < 1e-3, k
