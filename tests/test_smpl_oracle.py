"""Known-answer tests of the SMPL restatement (oracle/smpl_oracle.py) — properties of linear blend skinning that hold for
any asset pack.  Parity with smplx itself is unpinned (smplx and the SMPL assets are absent); CPU only."""
import math

import torch

from oracle import smpl_oracle as S


def _rot(axis, angle):
    axis = torch.tensor(axis, dtype=torch.float64)
    axis = axis / axis.norm()
    K = torch.tensor([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + math.sin(angle) * K + (1 - math.cos(angle)) * (K @ K)


def test_rest_pose_returns_template_and_regressed_joints():
    a = S.synthetic_assets(0, torch.float64)
    betas = torch.zeros(2, 10, dtype=torch.float64)
    R = torch.eye(3, dtype=torch.float64).expand(2, 24, 3, 3).contiguous()
    verts, j24 = S.lbs(betas, R, a)
    assert torch.allclose(verts, a["v_template"].expand(2, -1, -1), atol=1e-12)
    assert torch.allclose(j24, (a["J_regressor"] @ a["v_template"]).expand(2, -1, -1), atol=1e-12)


def test_shape_blend_is_linear_in_betas():
    a = S.synthetic_assets(1, torch.float64)
    R = torch.eye(3, dtype=torch.float64).expand(1, 24, 3, 3).contiguous()
    b1, b2 = torch.randn(1, 10, dtype=torch.float64), torch.randn(1, 10, dtype=torch.float64)
    v0, _ = S.lbs(torch.zeros(1, 10, dtype=torch.float64), R, a)
    v1, _ = S.lbs(b1, R, a)
    v2, _ = S.lbs(b2, R, a)
    v12, _ = S.lbs(b1 + b2, R, a)
    assert torch.allclose(v12 - v0, (v1 - v0) + (v2 - v0), atol=1e-10)       # rest pose: skinning is the identity


def test_pure_global_rotation_rotates_about_the_root_joint():
    a = S.synthetic_assets(2, torch.float64)
    betas = 0.5 * torch.randn(1, 10, dtype=torch.float64)
    Rg = _rot([0.3, -1.0, 0.5], 1.1)
    R = torch.eye(3, dtype=torch.float64).expand(1, 24, 3, 3).clone()
    R[0, 0] = Rg
    rest, j_rest = S.lbs(betas, torch.eye(3, dtype=torch.float64).expand(1, 24, 3, 3).contiguous(), a)
    verts, j24 = S.lbs(betas, R, a)
    J0 = j_rest[0, 0]
    assert torch.allclose(verts[0], (rest[0] - J0) @ Rg.t() + J0, atol=1e-10)
    assert torch.allclose(j24[0], (j_rest[0] - J0) @ Rg.t() + J0, atol=1e-10)


def test_joint_selection_and_regressor_override():
    a = S.synthetic_assets(3, torch.float64)
    betas = 0.3 * torch.randn(2, 10, dtype=torch.float64)
    R = torch.stack([torch.stack([_rot([1, j % 3, 0.2 * j], 0.1 * j + 0.05 * b) for j in range(24)]) for b in range(2)])
    verts, j49 = S.smpl_forward(betas, R, a)
    _, j24 = S.lbs(betas, R, a)
    assert j49.shape == (2, 49, 3)
    assert torch.equal(j49[:, 8], j24[:, 0])                                    # 'OP MidHip' -> SMPL joint 0
    assert torch.equal(j49[:, 0], verts[:, S.EXTRA_VERTEX_IDS[0]])              # 'OP Nose' -> joint 24 = first selected vertex
    assert torch.allclose(j49[:, 39], a["J_regressor_extra"][4] @ verts, atol=1e-12)   # 'Pelvis (MPII)' -> 49 = 45 + 4
    _, j17 = S.smpl_forward(betas, R, a, a["J_regressor_h36m"])
    assert j17.shape == (2, 17, 3) and torch.allclose(j17, torch.einsum("jv,bvk->bjk", a["J_regressor_h36m"], verts))


def test_training_tail_body_model_matches_oracle_and_is_differentiable():
    """The autograd tail of the training path — rot6d -> rotmat (csrc/decode_bwd.cu), the body model (csrc/smpl.cu, forward
    and backward kernels behind maed_b200.train._SmplBody), the projection — on the CUDA-on-CPU test build against the oracle
    restatement, values and gradients (float64 autograd over the oracle)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import harness
    from maed_b200 import train
    from maed_b200.models.modules import SMPLHead
    a = S.synthetic_assets(4)
    head = SMPLHead().load_assets(a)
    betas = (0.4 * torch.randn(3, 10)).requires_grad_(True)
    pose6d = torch.randn(3, 144, requires_grad=True)
    cam = torch.tensor([[0.9, 0.1, -0.1]]).repeat(3, 1)
    with harness.product_on_cpu():
        out = train.decode_outputs(pose6d, betas, cam, 49, head)
        v_ref, j_ref = S.smpl_forward(betas.detach().double(), out["rotmat"].detach().double(),
                                      {k: (v.double() if v.dtype.is_floating_point else v) for k, v in a.items()})
        assert (out["verts"].double() - v_ref).abs().max() < 1e-5 and (out["kp_3d"].double() - j_ref).abs().max() < 1e-5
        (out["kp_2d"].square().sum() + out["kp_3d"].sum()).backward()
        assert pose6d.grad is not None and betas.grad is not None and torch.isfinite(pose6d.grad).all() and pose6d.grad.abs().sum() > 0
        # the same scalar through the oracle in float64
        from oracle import maed_oracle as O
        p64, b64 = pose6d.detach().double().requires_grad_(True), betas.detach().double().requires_grad_(True)
        a64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in a.items()}
        _, j64 = S.smpl_forward(b64, O.rot6d_to_rotmat(p64).reshape(3, 24, 3, 3), a64)
        (O.project_keypoints(j64, cam.double()).square().sum() + j64.sum()).backward()
        err = lambda x, y: ((x.double() - y).norm() / y.norm()).item()  # noqa: E731
        assert err(pose6d.grad, p64.grad) < 1e-4 and err(betas.grad, b64.grad) < 1e-4
        out17 = train.decode_outputs(pose6d, betas, cam, 17, head, a["J_regressor_h36m"])
        assert out17["kp_3d"].shape == (3, 17, 3)
