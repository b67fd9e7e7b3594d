"""The training-path oracle (autograd over oracle/maed_oracle.py) against gradient digests produced by the unmodified
reference (tests/golden/make_golden_grads.py).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import maed_oracle as O
from oracle import synth
from helpers import GOLDEN_DIR, reference_shapes

CASES = ["grads_parallel_ktd", "grads_series_ktd", "grads_vanilla_ktd", "grads_vanilla_iterative", "grads_temporal_ktd", "grads_coupling_ktd"]


def load_grad_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, T, seed = [int(v) for v in z["meta"]]
    return z, str(z["mode"]), str(z["decoder"]), N, T, seed


def digest(g, nsamp=8):
    g = g.detach().double().reshape(-1).cpu()
    idx = np.unique(np.linspace(0, g.numel() - 1, nsamp).round().astype(np.int64))
    return np.array([g.norm().item(), g.sum().item()]), g[torch.from_numpy(idx)].numpy()


def check_against_golden(z, grads, tol):
    worst = 0.0
    for k in [str(s) for s in z["names"]]:
        assert k in grads, "no gradient for %s" % k
        stats, samp = digest(grads[k])
        ref_stats, ref_samp = z["g_stats/" + k], z["g_samp/" + k]
        scale = max(ref_stats[0], 1e-12)
        err_n = abs(stats[0] - ref_stats[0]) / scale
        assert err_n < tol, "%s: |g| %.6e vs reference %.6e" % (k, stats[0], ref_stats[0])
        # samples: compare relative to the RMS entry of the reference gradient
        rms = ref_stats[0] / np.sqrt(grads[k].numel())
        assert np.abs(samp - ref_samp).max() <= tol * 50 * max(rms, 1e-20) + 1e-12, k
        worst = max(worst, err_n)
    return worst


@pytest.mark.parametrize("name", CASES)
def test_oracle_grads_match_reference(name):
    z, mode, dec, N, T, seed = load_grad_case(name)
    shapes = reference_shapes(mode, dec)
    sd = synth.synth_state_dict(shapes, seed)
    x = synth.synth_frames(N, T, seed)
    nt = N * T
    A = synth.synth_tensor("grad_probe.pose", (nt, 144), seed)
    B = synth.synth_tensor("grad_probe.shape", (nt, 10), seed)
    C = synth.synth_tensor("grad_probe.cam", (nt, 3), seed)
    L, grads, _ = O.maed_param_grads(x, sd, A, B, C, mode, dec)
    assert abs(L.item() - float(z["loss"])) <= 1e-5 * max(1.0, abs(float(z["loss"])))
    worst = check_against_golden(z, grads, 2e-4)
    print("%s: worst relative |grad| error vs reference %.2e" % (name, worst))
