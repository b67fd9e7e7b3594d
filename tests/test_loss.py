"""Fused training loss (maed_b200/loss.py + csrc/loss.cu) against the reference's lib/core/loss.py.

CPU (default suite):
  * oracle/loss_oracle.py against golden values AND gradients of the UNMODIFIED reference classes (tests/golden/loss_*.npz,
    tests/golden/make_golden_loss.py) -> the oracle is pinned;
  * the real loss.cu on the CUDA-on-CPU test build through the PRODUCT module maed_b200.loss (LossVideo / LossImage / Loss):
    every loss_dict entry, the total and the three gradients against the same golden files, and against the oracle on
    bench-sized inputs (8 clips x 16 frames).
GPU (`-m gpu`): the same checks on the product library; gated like the training path until a B200 has run them.
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, rel_err
from oracle import loss_oracle as LO

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
CASES = ["loss_video_stage2", "loss_video_accl", "loss_video_novalid", "loss_image_stage1"]
KW = {"e_loss_weight": "w_kp2d", "e_3d_loss_weight": "w_kp3d", "e_pose_loss_weight": "w_pose", "e_shape_loss_weight": "w_shape",
      "e_smpl_norm_loss": "w_norm", "e_smpl_accl_loss": "w_accl"}


def _case(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n2, n3, T, seed = [int(v) for v in z["meta"]]
    kind = str(z["kind"])
    preds, d3, d2 = LO.synth_loss_case(n2, n3, T, seed, image=(kind == "image"))
    if name == "loss_video_novalid":
        d3["w_smpl"].zero_()
    kw = {str(k): float(v) for k, v in zip(z["kw_keys"], z["kw_vals"])}
    return z, kind, preds, d3, d2, kw


def _check(z, total, d, grads, tol=2e-5):
    assert [str(k) for k in z["keys"]] == list(d.keys())
    assert abs(float(total.detach()) - float(z["total"])) <= tol * abs(float(z["total"]))
    for k in d:
        ref = float(z["term_" + k])
        assert abs(float(d[k].detach()) - ref) <= tol * max(abs(ref), 1e-6), (k, float(d[k].detach()), ref)
    for k, g in zip(("kp_2d", "kp_3d", "theta"), grads):
        ref = z["grad_" + k]
        if np.abs(ref).max() == 0:
            assert float(g.abs().max()) == 0.0, k
        else:
            assert rel_err(g.reshape(ref.shape), ref) < 5e-5, k


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    z, kind, preds, d3, d2, kw = _case(name)
    preds = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    okw = {KW[k]: v for k, v in kw.items()}
    total, d = LO.loss_video(preds, d3, d2, **okw) if kind == "video" else LO.loss_image(preds, d3, **okw)
    grads = torch.autograd.grad(total, [preds["kp_2d"], preds["kp_3d"], preds["theta"]])
    _check(z, total, d, grads, tol=1e-6)


def _product(kind, preds, d3, d2, kw, device):
    from maed_b200 import loss as L
    mv = lambda t: {k: v.to(device) for k, v in t.items()} if t else t  # noqa: E731
    preds = {k: v.to(device).clone().requires_grad_(True) for k, v in preds.items()}
    if kind == "video":
        total, d = L.Loss(device=device, **kw)(preds, target_3d=mv(d3), target_2d=mv(d2))
    else:
        kw = {k: v for k, v in kw.items() if k != "e_smpl_accl_loss"}
        total, d = L.LossImage(device=device, **kw)(preds, mv(d3))
    (3.0 * total).backward()                                   # a non-unit upstream gradient must scale through
    grads = [preds[k].grad / 3.0 for k in ("kp_2d", "kp_3d", "theta")]
    assert all(not v.requires_grad for v in d.values())
    return total.detach(), d, grads


@pytest.fixture(scope="module")
def harness():
    import harness as h
    h.load()
    return h


@pytest.mark.parametrize("name", CASES)
def test_emulated_fused_loss_matches_reference_golden(harness, name):
    z, kind, preds, d3, d2, kw = _case(name)
    with harness.product_on_cpu():
        total, d, grads = _product(kind, preds, d3, d2, kw, "cpu")
    _check(z, total, d, grads)


def test_emulated_fused_loss_bench_shape_vs_oracle(harness):
    """8 clips x 16 frames (3-D labels) + 4 clips with 2-D labels only, stage-2 weights with the accl term switched on."""
    preds, d3, d2 = LO.synth_loss_case(4, 8, 16, 7)
    kw = dict(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06, e_smpl_norm_loss=1.,
              e_smpl_accl_loss=0.5)
    p = {k: v.clone().requires_grad_(True) for k, v in preds.items()}
    ref_total, ref_d = LO.loss_video(p, d3, d2, **{KW[k]: v for k, v in kw.items()})
    ref_g = torch.autograd.grad(ref_total, [p["kp_2d"], p["kp_3d"], p["theta"]])
    with harness.product_on_cpu():
        total, d, grads = _product("video", preds, d3, d2, kw, "cpu")
        again = _product("video", preds, d3, d2, kw, "cpu")
    assert abs(float(total) - float(ref_total)) < 2e-5 * float(ref_total)
    for k in ref_d:
        assert abs(float(d[k]) - float(ref_d[k])) <= 2e-5 * max(abs(float(ref_d[k])), 1e-6), k
    for g, r in zip(grads, ref_g):
        assert rel_err(g, r) < 5e-5
    assert float(again[0]) == float(total) and all(torch.equal(a, b) for a, b in zip(again[2], grads))   # bit-reproducible


def test_merge_loss_and_dispatch():
    from maed_b200.loss import Loss
    crit = Loss(device="cpu")
    assert crit({}, other=1) == (0, {})
    total, d = crit.merge_loss(torch.tensor(2.0), {"a": torch.tensor(1.0)}, torch.tensor(4.0), {"a": torch.tensor(3.0), "b": torch.tensor(5.0)},
                               vid_w=0.25, img_w=0.75)
    assert float(total) == 3.5 and float(d["a"]) == 2.5 and float(d["b"]) == 3.75


def test_cpu_tensors_fail_loudly():
    from maed_b200.loss import LossImage
    preds, d3, _ = LO.synth_loss_case(0, 2, 1, 0, image=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LossImage(device="cpu")(preds, d3)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fused_loss_matches_reference_golden_gpu(lib, name):
    z, kind, preds, d3, d2, kw = _case(name)
    total, d, grads = _product(kind, preds, d3, d2, kw, "cuda")
    _check(z, total.cpu(), {k: v.cpu() for k, v in d.items()}, [g.cpu() for g in grads])


def test_stage2_iteration_with_the_fused_loss_on_the_emulator(harness):
    """reference trainer.py:188,240-245 with the product modules: 'ste' model in train() mode -> LossVideo (3-D labelled clips +
    2-D-only clips in one batch, trainer.py:181-187) -> backward -> FusedAdam; the loss must be finite and go down."""
    from oracle import synth
    with harness.product_on_cpu():
        from maed_b200.loss import Loss
        from maed_b200.models import MAED
        from maed_b200.train import FusedAdam
        m = MAED("ste", 1, 12, "vanilla", "ktd", 1024)
        synth.fill_module_(m, 4)
        m = m.train().enable_training(True, dropout_p=0.0)
        crit = Loss(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06, device="cpu")
        opt = FusedAdam.for_model(m, lr=2e-4, weight_decay=1e-5)
        _, d3, d2 = LO.synth_loss_case(1, 1, 2, 9)                               # 1 clip with 2-D labels only + 1 clip with 3-D labels
        x = synth.synth_frames(2, 2, 9)                                          # the 2-D clip comes first (loss.py:171-176)
        losses = []
        for _ in range(2):
            opt.zero_grad()
            loss, terms = crit(m(x), target_3d=d3, target_2d=d2)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
        assert all(np.isfinite(losses)) and losses[1] < losses[0], losses
        assert list(terms) == ["loss_kp_2d", "loss_kp_3d", "loss_shape", "loss_pose", "loss_norm"]
