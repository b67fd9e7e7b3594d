"""Shared test helpers: golden-file access, synthetic weights, error metrics."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["c1_parallel_ktd", "series_ktd", "vanilla_ktd", "coupling_ktd", "temporal_ktd", "parallel_iterative",
                "series_iterative", "parallel_ktd_T1", "parallel_ktd_T32", "parallel_ktd_T16"]


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, T, seed, tf = [int(v) for v in g["meta"]]
    return g, dict(N=N, T=T, seed=seed, temp_frames=tf, mode=str(g["mode"]), decoder=str(g["decoder"]),
                   encoder=str(g["encoder"]) if "encoder" in g.files else "ste")


def rel_err(a, b):
    a = torch.as_tensor(np.asarray(a) if not torch.is_tensor(a) else a).double().cpu()
    b = torch.as_tensor(np.asarray(b) if not torch.is_tensor(b) else b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_model(meta, device="cpu", precision="split"):
    """B200 MAED module filled with the synthetic weights of the golden case."""
    from maed_b200.models import MAED
    from oracle import synth
    m = MAED(meta.get("encoder", "ste"), 6, 12, meta["mode"], meta["decoder"], 1024, precision=precision,
             temp_frames=meta["temp_frames"])
    synth.fill_module_(m, meta["seed"])
    return m.to(device)


def state_dict_of(model):
    sd = {k: v.detach() for k, v in model.named_parameters()}
    sd.update({k: v.detach() for k, v in model.named_buffers()})
    return sd


def reference_shapes(mode, decoder):
    """{state_dict key: shape} of the reference model for (st_mode, decoder), from tests/golden/state_dict_keys.json."""
    import json
    spec = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))
    return {k: tuple(v) for k, v in spec["%s/%s" % (mode, decoder)].items() if "smpl" not in k}
