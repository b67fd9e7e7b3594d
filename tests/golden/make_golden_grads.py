"""Generates tests/golden/grads_*.npz from the UNMODIFIED reference (CPU, fp32, autograd) — build container only:

    python tests/golden/make_golden_grads.py

The reference model (through oracle/ref_shim.py, synthetic weights from oracle/synth.py) runs in eval() mode (the
KTD dropout of ktd.py:54-56 is the only train/eval difference and is random) with autograd enabled.  The scalar

    L = sum(pose6d * A) + sum(shape * B) + sum(cam * C)          A, B, C = synth_tensor("grad_probe.*", ...)

is back-propagated; for each of the 305 parameters the file stores a digest of dL/dp (L2 norm, sum, 8 samples at
fixed positions) — 72 M gradient values cannot be committed.  The decoder outputs pose6d/shape/cam are the boundary
of the training engine (the geometry tail behind them is differentiated by autograd in the product too).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, synth  # noqa: E402

# name -> (st_mode, decoder, N, T, seed)
CASES = {
    "grads_parallel_ktd": ("parallel", "ktd", 1, 4, 11),
    "grads_series_ktd": ("series", "ktd", 1, 2, 12),
    "grads_vanilla_ktd": ("vanilla", "ktd", 2, 1, 13),
    # encoder='cnn' in train() mode: BatchNorm normalises with the statistics of the batch and updates its running buffers
    # (dropout modules stay in eval mode: random); also stores the running statistics after the step
    "grads_cnn_ktd": ("vanilla", "ktd", 2, 2, 14, "cnn"),
    # iterative regressor (spin.py:51-74): three passes through the same fc1 / fc2 / heads, state fed back into fc1
    "grads_vanilla_iterative": ("vanilla", "iterative", 1, 2, 15),
    "grads_cnn_iterative": ("vanilla", "iterative", 3, 1, 16, "cnn"),
    "grads_temporal_ktd": ("temporal", "ktd", 1, 3, 17),          # token mean -> attention across frames only
    "grads_coupling_ktd": ("coupling", "ktd", 1, 2, 18),          # joint attention over the T * 197 tokens of the clip
}
NSAMP = 8


def sample_index(numel):
    return np.unique(np.linspace(0, numel - 1, NSAMP).round().astype(np.int64))


def grad_digest(g):
    g = g.detach().double().reshape(-1)
    idx = sample_index(g.numel())
    return np.array([g.norm().item(), g.sum().item()], np.float64), g[torch.from_numpy(idx)].numpy()


def probes(nt, seed):
    return (synth.synth_tensor("grad_probe.pose", (nt, 144), seed), synth.synth_tensor("grad_probe.shape", (nt, 10), seed),
            synth.synth_tensor("grad_probe.cam", (nt, 3), seed))


def run_case(name):
    mode, dec, N, T, seed = CASES[name][:5]
    encoder = CASES[name][5] if len(CASES[name]) > 5 else "ste"
    out_dir = os.path.dirname(os.path.abspath(__file__))
    model = ref_shim.build_reference_model(mode, dec, encoder=encoder)
    synth.fill_module_(model, seed)
    model.eval()
    if encoder == "cnn":
        model.train()
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.eval()
    x = synth.synth_frames(N, T, seed)
    grabbed = {}
    orig = model.decoder.get_output

    def rec(pose, shape, cam, J):
        grabbed.update(pose=pose, shape=shape, cam=cam)
        return orig(pose, shape, cam, J)

    model.decoder.get_output = rec
    model(x)
    A, B, C = probes(N * T, seed)
    L = (grabbed["pose"] * A).sum() + (grabbed["shape"] * B).sum() + (grabbed["cam"] * C).sum()
    L.backward()
    rec_out = {"meta": np.array([N, T, seed], np.int64), "mode": np.array(mode), "decoder": np.array(dec),
               "loss": np.array(L.item(), np.float64)}
    names = []
    for k, p in model.named_parameters():
        if "smpl" in k:
            continue
        assert p.grad is not None, k
        stats, samp = grad_digest(p.grad)
        names.append(k)
        rec_out["g_stats/" + k] = stats
        rec_out["g_samp/" + k] = samp
    rec_out["names"] = np.array(names)
    if encoder == "cnn":
        rec_out["encoder"] = np.array(encoder)
        for k in ("pose", "shape", "cam"):
            rec_out["out_" + k] = grabbed[k].detach().numpy()
        bufs = [k for k, b in model.named_buffers() if k.endswith(("running_mean", "running_var"))]
        rec_out["buffers"] = np.array(bufs)
        for k, b in model.named_buffers():
            if k in bufs:
                rec_out["buf/" + k] = b.detach().numpy()
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **rec_out)
    print("%-22s %s/%s N=%d T=%d loss=%.6f params=%d -> %.0f KB" % (name, mode, dec, N, T, L.item(), len(names),
                                                                  os.path.getsize(path) / 1024))


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        run_case(n)
