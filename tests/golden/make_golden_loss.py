"""Generates tests/golden/loss_*.npz from the UNMODIFIED reference loss classes (lib/core/loss.py), CPU fp32:

    python tests/golden/make_golden_loss.py

Inputs come from oracle/loss_oracle.py::synth_loss_case (seeded); stored: every loss_dict value, the total, and the
gradients of the total with respect to preds['kp_2d'], preds['kp_3d'], preds['theta'] (reference autograd).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import loss_oracle, ref_shim  # noqa: E402

# name -> (kind, n2, n3, T, seed, constructor kwargs)
CASES = {
    "loss_video_stage2": ("video", 1, 2, 4, 0, dict(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60.,
                                                      e_shape_loss_weight=0.06, e_smpl_norm_loss=1., e_smpl_accl_loss=0.)),
    "loss_video_accl":   ("video", 0, 2, 5, 1, dict(e_smpl_accl_loss=2.0)),
    "loss_video_novalid": ("video", 0, 1, 3, 2, dict()),
    "loss_image_stage1": ("image", 0, 6, 1, 3, dict(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60.,
                                                      e_shape_loss_weight=0.06)),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    ref_shim.load_reference()
    import lib.core.loss as ref_loss
    for name, (kind, n2, n3, T, seed, kw) in CASES.items():
        preds, d3, d2 = loss_oracle.synth_loss_case(n2, n3, T, seed, image=(kind == "image"))
        if name == "loss_video_novalid":
            d3["w_smpl"].zero_()
        preds = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
        if kind == "video":
            total, d = ref_loss.LossVideo(device="cpu", **kw)(preds, d3, d2)
        else:
            total, d = ref_loss.LossImage(device="cpu", **kw)(preds, d3)
        grads = torch.autograd.grad(total, [preds["kp_2d"], preds["kp_3d"], preds["theta"]], allow_unused=True)
        rec = {"meta": np.array([n2, n3, T, seed], np.int64), "kind": np.array(kind), "total": total.detach().numpy(),
               "kw_keys": np.array(list(kw.keys())), "kw_vals": np.array(list(kw.values()), np.float64),
               "keys": np.array(list(d.keys()))}
        for k, v in d.items():
            rec["term_" + k] = torch.as_tensor(v).detach().numpy()
        for k, gten, p in zip(("kp_2d", "kp_3d", "theta"), grads, (preds["kp_2d"], preds["kp_3d"], preds["theta"])):
            rec["grad_" + k] = (gten if gten is not None else torch.zeros_like(p)).numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **rec)
        print(name, float(total), {k: float(v) for k, v in d.items()})


if __name__ == "__main__":
    main()
