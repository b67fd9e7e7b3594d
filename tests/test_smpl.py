"""SMPL forward kernels (csrc/smpl.cu) against oracle/smpl_oracle.py on the seeded synthetic asset pack, and the wiring of
verts / kp_3d / kp_2d into MAED.forward.  Two backends (see tests/test_bwd_ops.py): ``emu`` = the real kernel sources on
the CUDA-on-CPU shim (default CPU suite), ``cuda`` = the product library on a B200 (`-m gpu`; green on hardware since round 2)."""
import ctypes as C
import os
import sys

import pytest
import torch

from helpers import rel_err
from oracle import maed_oracle as O
from oracle import smpl_oracle as S
from oracle import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

_CUDA_MARKS = [pytest.mark.gpu]


@pytest.fixture(params=["emu", pytest.param("cuda", marks=_CUDA_MARKS)])
def dev(request):
    if request.param == "emu":
        import harness
        with harness.activate():
            yield "cpu"
    else:
        from maed_b200 import build
        build.build()
        yield "cuda"


def _inputs(R, seed):
    g = torch.Generator().manual_seed(seed)
    pose6d = torch.randn(R, 144, generator=g)
    pose6d[0] = torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24)               # rest pose
    betas = 0.5 * torch.randn(R, 10, generator=g)
    betas[0] = 0
    return betas, O.rot6d_to_rotmat(pose6d).reshape(R, 24, 3, 3)


@pytest.mark.parametrize("R,use_reg", [(5, False), (130, False), (7, True)])
def test_smpl_forward_matches_oracle(dev, R, use_reg):
    from maed_b200 import _lib
    from maed_b200.models.modules import SMPLHead
    a = S.synthetic_assets(0)
    head = SMPLHead().load_assets(a).to(dev)
    betas, rot = _inputs(R, 50)
    reg = a["J_regressor_h36m"] if use_reg else None
    v_ref, j_ref = S.smpl_forward(betas.double(), rot.double(), {k: (v.double() if v.dtype.is_floating_point else v) for k, v in a.items()},
                                  reg.double() if use_reg else None)
    assets = _lib.MaedSmplAssets(*[_lib.ptr(getattr(head, k)) for k in (
        "v_template", "shapedirs", "posedirs", "J_template", "J_shapedirs", "lbs_weights", "J_regressor_extra", "parents",
        "extra_vertex_ids", "joint_map")])
    nj = 17 if use_reg else 49
    verts = torch.empty(R, 6890, 3, device=dev)
    joints = torch.empty(R, nj, 3, device=dev)
    nbytes = _lib.load().maed_smpl_scratch_bytes(R)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    regd = reg.to(dev).contiguous() if use_reg else None
    betas_d, rot_d = betas.to(dev).contiguous(), rot.to(dev).contiguous()       # must outlive the call
    _lib.call("maed_smpl_forward", C.byref(assets), _lib.ptr(betas_d), _lib.ptr(rot_d), R, _lib.ptr(regd),
              17 if use_reg else 0, _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(scratch), C.c_size_t(nbytes), _lib.stream_ptr())
    assert rel_err(verts, v_ref) < 1e-5 and rel_err(joints, j_ref) < 1e-5
    assert rel_err(verts[0], a["v_template"]) < 1e-6                             # rest pose, zero betas
    # projection of the joints (maed_op_decode_outputs), as MAED._smpl chains it
    g = torch.Generator().manual_seed(51)
    pose6d, cam = torch.randn(R, 144, generator=g).to(dev), (torch.rand(R, 3, generator=g) + 0.5).to(dev)
    rot_o, theta, kp2d = torch.empty(R, 24, 3, 3, device=dev), torch.empty(R, 85, device=dev), torch.empty(R, nj, 2, device=dev)
    _lib.call("maed_op_decode_outputs", _lib.ptr(pose6d), _lib.ptr(betas_d), _lib.ptr(cam), R, _lib.ptr(joints), nj, _lib.ptr(rot_o),
              _lib.ptr(theta), _lib.ptr(kp2d), _lib.stream_ptr())
    assert rel_err(kp2d, O.project_keypoints(j_ref.float(), cam.cpu())) < 1e-5


@pytest.mark.parametrize("R,use_reg", [(3, False), (2, True), (37, False)])
def test_smpl_backward_matches_oracle_autograd(dev, R, use_reg):
    """maed_smpl_backward (d_verts, d_joints -> d_betas, d_rotmat) against float64 autograd over the oracle restatement; the
    49-joint map repeats joints (summed), the vertex-selected joints scatter into d_verts, J_regressor rows replace them."""
    from maed_b200 import _lib
    from maed_b200.models.modules import SMPLHead
    if R > 8 and dev == "cpu":
        pytest.skip("larger batch: hardware only")
    a = S.synthetic_assets(2)
    a64 = {k: (v.double() if v.dtype.is_floating_point else v) for k, v in a.items()}
    head = SMPLHead().load_assets(a).to(dev)
    betas, rot = _inputs(R, 60)
    g = torch.Generator().manual_seed(61)
    nj = 17 if use_reg else 49
    d_verts, d_joints = torch.randn(R, 6890, 3, generator=g), torch.randn(R, nj, 3, generator=g) * 10.0
    reg = a["J_regressor_h36m"] if use_reg else None
    bd, rd = betas.double().requires_grad_(True), rot.double().requires_grad_(True)
    v_ref, j_ref = S.smpl_forward(bd, rd, a64, reg.double() if use_reg else None)
    ((v_ref * d_verts.double()).sum() + (j_ref * d_joints.double()).sum()).backward()
    assets = _lib.MaedSmplAssets(*[_lib.ptr(getattr(head, k)) for k in (
        "v_template", "shapedirs", "posedirs", "J_template", "J_shapedirs", "lbs_weights", "J_regressor_extra", "parents",
        "extra_vertex_ids", "joint_map")])
    nbytes = _lib.load().maed_smpl_backward_scratch_bytes(R)
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    t = [x.to(dev).contiguous() for x in (betas, rot, d_verts, d_joints)]
    regd = reg.to(dev).contiguous() if use_reg else None
    d_betas, d_rot = torch.empty(R, 10, device=dev), torch.empty(R, 24, 3, 3, device=dev)
    _lib.call("maed_smpl_backward", C.byref(assets), _lib.ptr(t[0]), _lib.ptr(t[1]), R, _lib.ptr(regd), 17 if use_reg else 0,
              _lib.ptr(t[2]), _lib.ptr(t[3]), _lib.ptr(d_betas), _lib.ptr(d_rot), _lib.ptr(scratch), C.c_size_t(nbytes),
              _lib.stream_ptr())
    assert rel_err(d_betas, bd.grad) < 2e-5 and rel_err(d_rot, rd.grad) < 2e-5
    # without a vertex gradient (what the reference's losses produce: only the joints are penalised)
    bd.grad = rd.grad = None
    _, j_ref = S.smpl_forward(bd, rd, a64, reg.double() if use_reg else None)
    (j_ref * d_joints.double()).sum().backward()
    _lib.call("maed_smpl_backward", C.byref(assets), _lib.ptr(t[0]), _lib.ptr(t[1]), R, _lib.ptr(regd), 17 if use_reg else 0,
              None, _lib.ptr(t[3]), _lib.ptr(d_betas), _lib.ptr(d_rot), _lib.ptr(scratch), C.c_size_t(nbytes), _lib.stream_ptr())
    assert rel_err(d_betas, bd.grad) < 2e-5 and rel_err(d_rot, rd.grad) < 2e-5


@pytest.mark.gpu
def test_model_outputs_with_body_model(lib):
    from maed_b200.models import MAED
    a = S.synthetic_assets(1)
    m = MAED("ste", 6, 12, "vanilla", "ktd", 1024, smpl_assets=a)
    synth.fill_module_(m, 3)
    m = m.cuda().eval()
    x = synth.synth_frames(1, 2, 3)
    out = m(x.cuda())
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    taps = {}
    O.maed_forward(x, sd, "vanilla", "ktd", taps=taps)
    rot = O.rot6d_to_rotmat(taps["pose6d"]).reshape(2, 24, 3, 3)
    v_ref, j_ref = S.smpl_forward(taps["shape"], rot, a)
    assert rel_err(out["verts"].reshape(2, 6890, 3), v_ref) < 1e-3
    assert rel_err(out["kp_3d"].reshape(2, 49, 3), j_ref) < 1e-3
    assert rel_err(out["kp_2d"].reshape(2, 49, 2), O.project_keypoints(j_ref, taps["cam"])) < 1e-3
    out17 = m(x.cuda(), J_regressor=a["J_regressor_h36m"])
    assert out17["kp_3d"].shape == (1, 2, 17, 3) and out17["kp_2d"].shape == (1, 2, 17, 2)
