"""Geometry tail of the training path (maed_b200.train._PoseTail / _Project over csrc/decode_bwd.cu) against autograd over the
oracle's torch restatement of lib/utils/geometry.py:320-334,58-223 and lib/models/spin.py:113-157.  Two backends like
tests/test_bwd_ops.py: ``emu`` (CPU, default suite) and ``cuda`` (`-m gpu`, gated until a B200 has run it)."""
import os
import sys

import pytest
import torch

from helpers import rel_err
from oracle import maed_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
DEV = "cuda"
_CUDA_MARKS = [pytest.mark.gpu]


@pytest.fixture(params=["emu", pytest.param("cuda", marks=_CUDA_MARKS)])
def backend(request):
    global DEV
    if request.param == "emu":
        import harness
        DEV = "cpu"
        with harness.product_on_cpu():
            yield
    else:
        from maed_b200 import build
        build.build()
        DEV = "cuda"
        yield


def _inputs(R, seed):
    g = torch.Generator().manual_seed(seed)
    pose = torch.randn(R, 144, generator=g)                       # random 6-D vectors: rotations of all four quaternion cases
    pose[0] = torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24) + 1e-3 * torch.randn(144, generator=g)   # near identity
    shape, cam = torch.randn(R, 10, generator=g), torch.randn(R, 3, generator=g) * 0.2 + torch.tensor([0.9, 0.0, 0.0])
    return pose, shape, cam


def test_pose_tail_matches_oracle_autograd(backend):
    from maed_b200.train import _PoseTail
    R = 57
    pose, shape, cam = _inputs(R, 3)
    g = torch.Generator().manual_seed(4)
    w_theta, w_rot = torch.randn(R, 85, generator=g), torch.randn(R, 24, 3, 3, generator=g)
    # reference (float64 autograd over the oracle's restatement)
    pr, sr, cr = [t.double().requires_grad_(True) for t in (pose, shape, cam)]
    rot_r = O.rot6d_to_rotmat(pr).reshape(R, 24, 3, 3)
    aa_r = O.rotmat_to_angle_axis(rot_r.reshape(-1, 3, 3)).reshape(R, 72)
    theta_r = torch.cat([cr, aa_r, sr], dim=1)
    cases = {int(c) for c in ((rot_r[:, :, 2, 2] < 1e-6).long() * 2 + (rot_r[:, :, 0, 0] > rot_r[:, :, 1, 1]).long()).reshape(-1)}
    assert len(cases) >= 3                                        # the branches of the quaternion extraction are exercised
    gr = torch.autograd.grad((theta_r * w_theta.double()).sum() + (rot_r * w_rot.double()).sum(), [pr, sr, cr])
    # product
    p, s, c = [t.clone().to(DEV).requires_grad_(True) for t in (pose, shape, cam)]
    theta, rot = _PoseTail.apply(p, s, c)
    assert rel_err(theta, theta_r.detach()) < 2e-5 and rel_err(rot, rot_r.detach()) < 2e-6
    ((theta * w_theta.to(DEV)).sum() + (rot * w_rot.to(DEV)).sum()).backward()
    assert rel_err(p.grad, gr[0]) < 2e-4
    assert rel_err(s.grad, gr[1]) < 1e-6 and rel_err(c.grad, gr[2]) < 1e-6
    # only one of the two outputs used downstream
    p2 = pose.clone().to(DEV).requires_grad_(True)
    _PoseTail.apply(p2, shape.to(DEV), cam.to(DEV))[1].sum().backward()
    pr2 = pose.clone().double().requires_grad_(True)
    O.rot6d_to_rotmat(pr2).sum().backward()
    assert rel_err(p2.grad, pr2.grad) < 2e-4


@pytest.mark.parametrize("with_joints", [True, False])
def test_projection_matches_oracle_autograd(backend, with_joints):
    from maed_b200.train import _Project
    R, J = 33, 49
    _, _, cam = _inputs(R, 5)
    g = torch.Generator().manual_seed(6)
    kp3d = 0.3 * torch.randn(R, J, 3, generator=g)
    w = torch.randn(R, J, 2, generator=g)
    cr = cam.double().requires_grad_(True)
    kr = kp3d.double().requires_grad_(True) if with_joints else torch.zeros(R, J, 3, dtype=torch.float64)
    ref = O.project_keypoints(kr, cr)
    gr = torch.autograd.grad((ref * w.double()).sum(), [cr] + ([kr] if with_joints else []))
    c = cam.clone().to(DEV).requires_grad_(True)
    k = kp3d.clone().to(DEV).requires_grad_(True) if with_joints else None
    out = _Project.apply(k, c, J)
    assert rel_err(out, ref.detach()) < 2e-6
    (out * w.to(DEV)).sum().backward()
    assert rel_err(c.grad, gr[0]) < 2e-5
    if with_joints:
        assert rel_err(k.grad, gr[1]) < 2e-5
