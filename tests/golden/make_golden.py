"""Generates tests/golden/*.npz from the UNMODIFIED reference (CPU, fp32) — run in the build container:

    python tests/golden/make_golden.py            # all cases
    python tests/golden/make_golden.py c1_parallel_ktd

The reference's own `lib.models.MAED` is imported through oracle/ref_shim.py, its parameters are
overwritten with oracle/synth.py's deterministic values (key+shape+seed -> tensor), the synthetic clip
from synth_frames() is pushed through `model(x)`, and outputs + intermediate taps are stored.  The tests
re-create the same weights/inputs from the seed, so only these small files have to travel.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_shim, synth  # noqa: E402

# name -> (st_mode, decoder, N, T, seed, temp_frames)
CASES = {
    "c1_parallel_ktd":     ("parallel", "ktd", 1, 8, 0, 16),      # BASELINE.json configs[0]
    "series_ktd":          ("series", "ktd", 1, 4, 1, 16),
    "vanilla_ktd":         ("vanilla", "ktd", 1, 2, 2, 16),
    "coupling_ktd":        ("coupling", "ktd", 1, 2, 3, 16),
    "temporal_ktd":        ("temporal", "ktd", 1, 3, 4, 16),
    "parallel_iterative":  ("parallel", "iterative", 2, 2, 5, 16),
    "series_iterative":    ("series", "iterative", 1, 2, 6, 16),
    "parallel_ktd_T1":     ("parallel", "ktd", 3, 1, 7, 16),      # image batches: seqlen 1 (trainer.py:177-179)
    "parallel_ktd_T32":    ("parallel", "ktd", 1, 32, 8, 32),     # BASELINE configs[4]: temp_embed data -> 32 rows
    "parallel_ktd_T16":    ("parallel", "ktd", 1, 16, 9, 16),
    # encoder='cnn' (torchvision ResNet-50, config_stage1.yaml:68): 7-tuples, st_mode is ignored by the reference
    "cnn_ktd":             ("vanilla", "ktd", 1, 2, 10, 16, "cnn"),
    "cnn_iterative":       ("vanilla", "iterative", 3, 1, 11, 16, "cnn"),   # stage 1 trains on image batches (T = 1)
}


def run_case(name):
    mode, dec, N, T, seed, tf = CASES[name][:6]
    encoder = CASES[name][6] if len(CASES[name]) > 6 else "ste"
    out_dir = os.path.dirname(os.path.abspath(__file__))   # before load_reference() chdirs away
    model = ref_shim.build_reference_model(mode, dec, temp_frames=tf, encoder=encoder)
    synth.fill_module_(model, seed)
    x = synth.synth_frames(N, T, seed)
    taps = {}

    def hook(key):
        def f(_m, _i, o):
            taps[key] = o.detach()
        return f

    enc = model.encoder
    if encoder == "cnn":
        hs = [enc.maxpool.register_forward_hook(hook("stem"))]
        for i, st in enumerate((enc.layer1, enc.layer2, enc.layer3, enc.layer4)):
            hs.append(st.register_forward_hook(hook("stage%d" % i)))
    else:
        hs = [enc.patch_embed.backbone.stem.register_forward_hook(hook("stem"))]
        for i, st in enumerate(enc.patch_embed.backbone.stages):
            hs.append(st.register_forward_hook(hook("stage%d" % i)))
        for i, blk in enumerate(enc.blocks):
            hs.append(blk.register_forward_hook(hook("block%d" % i)))
    hs.append(enc.register_forward_hook(hook("feat")))
    orig_get_output = model.decoder.get_output

    def rec_get_output(pose, shape, cam, J):
        taps["pose6d"], taps["shape"], taps["cam"] = pose.detach(), shape.detach(), cam.detach()
        return orig_get_output(pose, shape, cam, J)

    model.decoder.get_output = rec_get_output
    t0 = time.time()
    with torch.no_grad():
        out = model(x)
    dt = time.time() - t0
    for h in hs:
        h.remove()
    rec = {"meta": np.array([N, T, seed, tf], np.int64), "mode": np.array(mode), "decoder": np.array(dec)}
    if encoder != "ste":
        rec["encoder"] = np.array(encoder)
        rec["state_dict_keys"] = np.array(sorted(k for k in model.state_dict() if "smpl" not in k))
        rec["state_dict_shapes"] = np.array([",".join(str(d) for d in model.state_dict()[k].shape)
                                             for k in sorted(k for k in model.state_dict() if "smpl" not in k)])
    for k in ("theta", "rotmat", "kp_2d"):
        rec["out_" + k] = out[k].numpy()
    rec["out_verts_absmax"] = np.array(out["verts"].abs().max().item(), np.float32)
    rec["out_kp_3d_absmax"] = np.array(out["kp_3d"].abs().max().item(), np.float32)
    for k in ("feat", "pose6d", "shape", "cam"):
        rec["tap_" + k] = taps[k].numpy()
    for k, v in taps.items():
        if k in ("feat", "pose6d", "shape", "cam"):
            continue
        sub, stats = synth.tap_digest(v)
        rec["dig_%s_sub" % k] = sub
        rec["dig_%s_stats" % k] = stats
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, **rec)
    print("%-22s %s/%s N=%d T=%d  %.1fs  -> %s (%.0f KB)" % (
        name, mode, dec, N, T, dt, os.path.basename(path), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    names = sys.argv[1:] or list(CASES)
    for n in names:
        run_case(n)
