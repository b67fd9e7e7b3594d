"""Backward kernels of the training path against torch autograd (float64) — one test per C-ABI entry, two backends:

  * ``emu``  (CPU, runs in the default `-m "not gpu"` suite): the REAL kernel sources of maed_b200/csrc compiled by g++
    against the CUDA-on-CPU shim (tests/emu/): checks the index logic, reductions and formulas of every CUDA-core kernel;
    the tcgen05 split-K kernel cannot run there (its case is skipped, the conv data-gradient case uses the contract stub);
  * ``cuda`` (`-m gpu`): the product library on a B200 (green on hardware since round 2).
"""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

DEV = "cuda"
_CUDA_MARKS = [pytest.mark.gpu]


@pytest.fixture(params=["emu", pytest.param("cuda", marks=_CUDA_MARKS)])
def L(request):
    global DEV
    from maed_b200 import _lib, ops
    if request.param == "emu":
        import harness
        DEV = "cpu"
        with harness.activate():
            yield _lib, ops
    else:
        from maed_b200 import build
        build.build()
        DEV = "cuda"
        yield _lib, ops


def _is_emu():
    return DEV == "cpu"


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=DEV, dtype=torch.float32) * scale


def _planes(x, ops):
    return ops.split(x)


def _join(p):
    return p[0].double() + p[1].double()


def test_transpose_planes(L):
    _lib, ops = L
    x = _rand(1000, 200, seed=1)
    p = _planes(x, ops)
    ld = 1000
    out = torch.zeros(2, 200, ld, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_transpose_planes", _lib.ptr(p), C.c_longlong(p[0].numel()), 1000, 200, 200, _lib.ptr(out),
              C.c_longlong(out[0].numel()), ld, _lib.stream_ptr())
    assert torch.equal(out[0], p[0].t()) and torch.equal(out[1], p[1].t())


@pytest.mark.parametrize("R,Cc,ld", [(3000, 333, 333), (3001, 768, 768), (130, 64, 128), (7, 1536, 1536)])
def test_colsum(L, R, Cc, ld):
    """scalar path (odd widths) and the 16-byte vectorised path (C, ld multiples of 4); scratch = 256 chunks x C floats"""
    _lib, _ = L
    x = _rand(R, ld, seed=2)
    scratch = torch.empty(_lib.load().maed_bwd_colsum_chunks() * Cc, device=DEV)
    out = torch.full((Cc,), 5.0, device=DEV)
    _lib.call("maed_bwd_colsum", _lib.ptr(x), C.c_longlong(ld), R, Cc, C.c_float(0.5), 1, _lib.ptr(scratch), _lib.ptr(out),
              _lib.stream_ptr())
    assert rel_err(out, 5.0 + 0.5 * x[:, :Cc].double().sum(0)) < 1e-6


@pytest.mark.parametrize("rows,C_", [(777, 768), (64, 1024), (5, 128)])
def test_layernorm_bwd(L, rows, C_):
    _lib, _ = L
    x, dy, add = _rand(rows, C_, seed=3), _rand(rows, C_, seed=4), _rand(rows, C_, seed=5)
    gamma = 1 + 0.1 * _rand(C_, seed=6)
    xd = x.double().requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = torch.zeros(C_, dtype=torch.float64, device=DEV, requires_grad=True)
    F.layer_norm(xd, (C_,), gd, bd, 1e-6).backward(dy.double())
    pr = _lib.load().maed_bwd_layernorm_partial_rows()
    partial = torch.empty(pr, 2 * C_, device=DEV)
    scratch = torch.empty(_lib.load().maed_bwd_colsum_chunks() * 2 * C_, device=DEV)
    dx, dg, db = torch.empty_like(x), torch.empty(C_, device=DEV), torch.empty(C_, device=DEV)
    _lib.call("maed_bwd_layernorm", _lib.ptr(dy), C.c_longlong(C_), _lib.ptr(x), C.c_longlong(C_), _lib.ptr(gamma), rows, C_,
              C.c_float(1e-6), _lib.ptr(add), _lib.ptr(dx), C.c_longlong(C_), _lib.ptr(partial), _lib.ptr(scratch), _lib.ptr(dg),
              _lib.ptr(db), _lib.stream_ptr())
    assert rel_err(dx, xd.grad + add.double()) < 2e-6
    assert rel_err(dg, gd.grad) < 2e-6 and rel_err(db, bd.grad) < 2e-6


@pytest.mark.parametrize("n,HW,C_,relu,res", [(3, 196, 256, 1, 0), (2, 3136, 64, 1, 0), (3, 196, 1024, 1, 1), (5, 784, 128, 0, 0),
                                              (2, 784, 512, 1, 1), (1, 50, 256, 0, 1)])
def test_groupnorm_train_forward(L, n, HW, C_, relu, res):
    """GroupNorm of the training forward (cluster-per-image kernel on the GPU, the multi-kernel path on the emulator): planes
    and the (sum, sumsq) statistics it leaves on the tape, vs torch float64 (reference resnetv2.py:35-49)."""
    _lib, ops = L
    x = _rand(n, HW, C_, scale=2.0, seed=21) + 0.3
    gamma, beta = 1 + 0.1 * _rand(C_, seed=22), 0.1 * _rand(C_, seed=23)
    r = _rand(n, HW, C_, seed=24) if res else None
    rp = ops.split(r) if res else None
    out = torch.empty(2, n, HW, C_, dtype=torch.float16, device=DEV)
    stats = torch.full((n, 32, 2), -1.0, dtype=torch.float64, device=DEV)
    _lib.call("maed_op_groupnorm_train", _lib.ptr(x), n, HW, C_, _lib.ptr(gamma), _lib.ptr(beta), C.c_float(1e-5), relu,
              _lib.ptr(rp) if res else None, C.c_longlong(rp[0].numel() if res else 0), _lib.ptr(out), C.c_longlong(out[0].numel()),
              _lib.ptr(stats), _lib.stream_ptr())
    ref = F.group_norm(x.permute(0, 2, 1).double(), 32, gamma.double(), beta.double(), 1e-5).permute(0, 2, 1)
    if res:
        ref = ref + r.double()
    if relu:
        ref = F.relu(ref)
    assert rel_err(_join(out), ref) < 1e-6
    xg = x.double().reshape(n, HW, 32, C_ // 32)
    assert rel_err(stats[:, :, 0], xg.sum((1, 3))) < 1e-6 and rel_err(stats[:, :, 1], (xg * xg).sum((1, 3))) < 1e-6


@pytest.mark.parametrize("n,HW,C_,relu,order", [(3, 196, 256, 0, 0), (2, 3136, 64, 1, 1), (2, 49, 1024, 0, 2), (5, 784, 128, 1, 2),
                                                (1, 200, 512, 1, 0), (4, 196, 256, 1, 1)])
def test_groupnorm_bwd(L, n, HW, C_, relu, order):
    """relu: dy is the gradient behind the ReLU that follows the norm (mask recomputed inside the kernels, reference
    resnetv2.py:35-49 GroupNormAct); order: the zig-zag image orders of the two passes (results must not depend on it)."""
    _lib, _ = L
    x, dy = _rand(n, HW, C_, seed=7), _rand(n, HW, C_, seed=8)
    gamma = 1 + 0.1 * _rand(C_, seed=9)
    beta = 0.2 * _rand(C_, seed=10)
    xd = x.double().permute(0, 2, 1).reshape(n, C_, HW, 1).requires_grad_(True)
    gd = gamma.double().requires_grad_(True)
    bd = beta.double().requires_grad_(True)
    y = F.group_norm(xd, 32, gd, bd, 1e-5)
    if relu:
        # keep the comparison away from the kink: elements whose pre-activation is within fp32 rounding of 0 get no gradient
        safe = (y.detach().abs() > 1e-4).to(dy.dtype).reshape(n, C_, HW).permute(0, 2, 1)
        dy = dy * safe.float()
        y = F.relu(y)
    y.backward(dy.double().permute(0, 2, 1).reshape(n, C_, HW, 1))
    stats = torch.empty(n * 64, dtype=torch.float64, device=DEV)
    red = torch.empty(n * (64 + 32 * C_), device=DEV)
    dgb = torch.empty(n, 2, C_, device=DEV)
    dx = torch.empty(2, n, HW, C_, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_groupnorm", _lib.ptr(dy), _lib.ptr(x), n, HW, C_, _lib.ptr(gamma), C.c_float(1e-5), _lib.ptr(stats),
              _lib.ptr(red), _lib.ptr(dgb), _lib.ptr(dx), C.c_longlong(dx[0].numel()), _lib.ptr(beta) if relu else None, order,
              _lib.stream_ptr())
    ref_dx = xd.grad.reshape(n, C_, HW).permute(0, 2, 1)
    assert rel_err(_join(dx), ref_dx) < 1e-5
    assert rel_err(dgb[:, 0].double().sum(0), gd.grad) < 1e-5 and rel_err(dgb[:, 1].double().sum(0), bd.grad) < 1e-5


@pytest.mark.parametrize("Cout,Cin,k", [(64, 64, 3), (256, 64, 1), (64, 3, 7)])
def test_wstd_bwd(L, Cout, Cin, k):
    _lib, _ = L
    w = _rand(Cout, Cin, k, k, scale=0.1, seed=10)
    kc = k * k * Cin
    kp = (kc + 31) // 32 * 32
    g = torch.zeros(Cout, kp, device=DEV)
    g_oihw = _rand(Cout, Cin, k, k, seed=11)
    g[:, :kc] = g_oihw.permute(0, 2, 3, 1).reshape(Cout, kc)         # packed layout [co][(kh,kw),ci]
    wd = w.double().requires_grad_(True)
    std, mean = torch.std_mean(wd, dim=[1, 2, 3], keepdim=True, unbiased=False)
    ((wd - mean) / (std + 1e-5)).backward(g_oihw.double())
    dw = torch.empty_like(w)
    _lib.call("maed_bwd_wstd", _lib.ptr(g), kp, _lib.ptr(w), Cout, Cin, k, k, C.c_float(1e-5), C.c_float(0.25), _lib.ptr(dw),
              _lib.stream_ptr())
    assert rel_err(dw, 0.25 * wd.grad) < 1e-5


def test_gelu_bwd_and_relu_mask(L):
    _lib, ops = L
    pre, d = _rand(1000, 3072, seed=12), _rand(1000, 3072, seed=13)
    pd = pre.double().requires_grad_(True)
    F.gelu(pd).backward(d.double())
    out = torch.empty(2, 1000, 3072, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_gelu", _lib.ptr(d), _lib.ptr(pre), C.c_longlong(pre.numel()), _lib.ptr(out), C.c_longlong(out[0].numel()),
              _lib.stream_ptr())
    assert rel_err(_join(out), pd.grad) < 1e-5
    act = _planes(torch.relu(pre), ops)
    dd = d.clone()
    _lib.call("maed_bwd_relu_mask", _lib.ptr(dd), _lib.ptr(act), C.c_longlong(dd.numel()), _lib.stream_ptr())
    assert torch.equal(dd, d * (pre > 0))


def test_maxpool_fwd_idx_and_bwd(L):
    _lib, _ = L
    n, H, W, C_ = 2, 112, 112, 64
    x = _rand(n, H, W, C_, seed=14)
    gamma, beta = 1 + 0.1 * _rand(C_, seed=15), 0.1 * _rand(C_, seed=16)
    d_pool = _rand(n, 56, 56, C_, seed=17)
    xd = x.double().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    y = torch.relu(F.group_norm(xd, 32, gamma.double(), beta.double(), 1e-5))
    y.retain_grad()
    yp = F.pad(y, [0, 1, 0, 1], value=float("-inf"))                     # TF-SAME: extra pixel bottom/right
    pooled = F.max_pool2d(yp, 3, 2)
    pooled.backward(d_pool.double().permute(0, 3, 1, 2))
    stats = torch.empty(n * 64, dtype=torch.float64, device=DEV)
    out = torch.empty(2, n, 56, 56, C_, dtype=torch.float16, device=DEV)
    idx = torch.empty(n, 56, 56, C_, dtype=torch.uint8, device=DEV)
    d_y = torch.empty(n, H, W, C_, device=DEV)
    _lib.call("maed_bwd_maxpool", _lib.ptr(x), n, H, W, C_, _lib.ptr(gamma), _lib.ptr(beta), C.c_float(1e-5), _lib.ptr(stats),
              _lib.ptr(out), C.c_longlong(out[0].numel()), _lib.ptr(idx), _lib.ptr(d_pool), _lib.ptr(d_y), _lib.stream_ptr())
    assert rel_err(_join(out), pooled.permute(0, 2, 3, 1)) < 1e-5
    # gradient w.r.t. the GN output with the ReLU mask applied (all-zero windows route d_pool to a masked position)
    assert rel_err(d_y, (y.grad * (y > 0)).permute(0, 2, 3, 1)) < 1e-6


def test_dilate_and_scatter(L):
    _lib, ops = L
    n, OH, C_ = 2, 14, 64
    src = _rand(n, OH, OH, C_, seed=18)
    p = _planes(src, ops)
    out = torch.empty(2, n, 2 * OH, 2 * OH, C_, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_dilate2", _lib.ptr(p), C.c_longlong(p[0].numel()), n, OH, OH, C_, 2 * OH, 2 * OH, _lib.ptr(out),
              C.c_longlong(out[0].numel()), _lib.stream_ptr())
    ref = torch.zeros(n, 2 * OH, 2 * OH, C_, dtype=torch.float64, device=DEV)
    ref[:, ::2, ::2] = _join(p)
    assert torch.equal(_join(out), ref)
    add = _rand(n, 2 * OH, 2 * OH, C_, seed=19)
    d_in = torch.empty_like(add)
    _lib.call("maed_bwd_scatter_stride2", _lib.ptr(src), n, OH, OH, C_, 2 * OH, 2 * OH, _lib.ptr(add), _lib.ptr(d_in), _lib.stream_ptr())
    ref2 = add.clone()
    ref2[:, ::2, ::2] += src
    assert torch.equal(d_in, ref2)


def test_blend_bwd(L):
    _lib, _ = L
    BT, ntok, C_ = 6, 197, 768
    xs, xt, d_ao = _rand(BT, ntok, C_, seed=20), _rand(BT, ntok, C_, seed=21), _rand(BT, ntok, C_, seed=22)
    logits, d_pool = _rand(BT, 2 * C_, seed=23), _rand(BT, 2 * C_, seed=24)
    xsd, xtd = xs.double().requires_grad_(True), xt.double().requires_grad_(True)
    ld = logits.double().requires_grad_(True)
    alpha = torch.softmax(ld.reshape(BT, 1, C_, 2), dim=-1)
    ao = xtd * alpha[..., 1] + xsd * alpha[..., 0]
    pooled = torch.cat([xsd.mean(1), xtd.mean(1)], dim=-1)
    ((ao * d_ao.double()).sum() + (pooled * d_pool.double()).sum()).backward()
    dl, dxs, dxt = torch.empty_like(logits), torch.empty_like(xs), torch.empty_like(xt)
    _lib.call("maed_bwd_blend", _lib.ptr(d_ao), _lib.ptr(xs), _lib.ptr(xt), _lib.ptr(logits), _lib.ptr(d_pool), BT, ntok, C_,
              _lib.ptr(dl), _lib.ptr(dxs), _lib.ptr(dxt), _lib.stream_ptr())
    assert rel_err(dl, ld.grad) < 1e-5 and rel_err(dxs, xsd.grad) < 1e-6 and rel_err(dxt, xtd.grad) < 1e-6


@pytest.mark.parametrize("ta,tb", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_sgemm(L, ta, tb):
    _lib, _ = L
    M, N, K = 130, 1030, 77
    A = _rand(K, M, seed=25) if ta else _rand(M, K, seed=25)
    B = _rand(N, K, seed=26) if tb else _rand(K, N, seed=26)
    Cm = _rand(M, N, seed=27)
    ref = 0.5 * (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double()) + 2.0 * Cm.double()
    _lib.call("maed_bwd_sgemm", ta, tb, M, N, K, C.c_float(0.5), _lib.ptr(A), A.shape[1], _lib.ptr(B), B.shape[1], C.c_float(2.0),
              _lib.ptr(Cm), N, _lib.stream_ptr())
    assert rel_err(Cm, ref) < 1e-5


def test_ktd_tree_bwd(L):
    _lib, _ = L
    from maed_b200.models.modules import ANCESTOR_INDEX
    R = 37
    base = _rand(R, 192, seed=28)
    blocks = [_rand(6, 6 * len(a), scale=0.3, seed=100 + j) for j, a in enumerate(ANCESTOR_INDEX)]
    w_anc = torch.cat([b.reshape(-1) for b in blocks if b.numel()])
    bd = base.double().requires_grad_(True)
    wd = [b.double().requires_grad_(True) for b in blocks]
    pose = []
    for j, a in enumerate(ANCESTOR_INDEX):
        o = bd[:, 6 * j:6 * j + 6]
        if a:
            o = o + torch.cat([pose[i] for i in a], dim=1) @ wd[j].t()
        pose.append(o)
    pose = torch.cat(pose, dim=1)
    d_pose, d_shape, d_cam = _rand(R, 144, seed=29), _rand(R, 10, seed=30), _rand(R, 3, seed=31)
    ((pose * d_pose.double()).sum() + (bd[:, 144:154] * d_shape.double()).sum() + (bd[:, 154:157] * d_cam.double()).sum()).backward()
    g_total, d_base = torch.empty(R, 144, device=DEV), torch.empty(R, 192, device=DEV)
    d_w = torch.empty(36 * 95, device=DEV)
    pose_f32 = pose.detach().float().contiguous()                       # must outlive the call (ptr() keeps no reference)
    _lib.call("maed_bwd_ktd_tree", _lib.ptr(d_pose), _lib.ptr(d_shape), _lib.ptr(d_cam), _lib.ptr(w_anc),
              _lib.ptr(pose_f32), R, C.c_float(1.0), _lib.ptr(g_total), _lib.ptr(d_base), 192, _lib.ptr(d_w),
              _lib.stream_ptr())
    assert rel_err(d_base, bd.grad) < 1e-5
    ref_w = torch.cat([w.grad.reshape(-1) for w in wd if w.numel()])
    assert rel_err(d_w, ref_w) < 1e-5


def _attention_ref(qkv, B, T, ntok, heads, scale, kind):
    BT = B * T
    q, k, v = qkv.reshape(BT, ntok, 3, heads, 64).permute(2, 0, 3, 1, 4)
    if kind == "spatial":
        a = (q @ k.transpose(-2, -1) * scale).softmax(-1)
        return (a @ v).transpose(1, 2).reshape(BT * ntok, heads * 64)
    if kind == "coupling":                      # all T * ntok tokens of a clip attend to each other
        rc = lambda t: t.reshape(B, T, heads, ntok, 64).transpose(1, 2).reshape(B, heads, T * ntok, 64)  # noqa: E731
        a = (rc(q) @ rc(k).transpose(-2, -1) * scale).softmax(-1)
        return (a @ rc(v)).reshape(B, heads, T, ntok, 64).permute(0, 2, 3, 1, 4).reshape(BT * ntok, heads * 64)
    r = lambda t: t.reshape(B, T, heads, ntok, 64).permute(0, 2, 3, 1, 4)  # noqa: E731
    a = (r(q) @ r(k).transpose(-2, -1) * scale).softmax(-1)
    return (a @ r(v)).permute(0, 3, 2, 1, 4).reshape(BT * ntok, heads * 64)


@pytest.mark.parametrize("kind,B,T,ntok", [("spatial", 2, 2, 197), ("spatial", 1, 3, 60), ("spatial_tc", 2, 2, 197),
                                           ("spatial_tc", 1, 3, 60), ("spatial_tc", 8, 16, 197), ("spatial_tc", 1, 2, 128),
                                           ("spatial_tc", 1, 1, 113), ("spatial_tc", 2, 1, 208), ("temporal", 2, 16, 197),
                                           ("temporal_tc", 2, 16, 197), ("temporal_tc", 8, 16, 197), ("temporal_tc", 3, 8, 50),
                                           ("temporal_tc", 1, 32, 64), ("temporal_tc", 2, 4, 33),
                                           ("spatial_fast", 2, 2, 197), ("spatial_fast", 8, 16, 197), ("spatial_fast", 1, 3, 60),
                                           ("temporal_fast", 2, 16, 197), ("temporal_fast", 8, 16, 197), ("temporal_fast", 1, 32, 64),
                                           ("temporal_fast", 3, 8, 50),
                                           ("temporal", 3, 5, 33), ("temporal", 2, 1, 20), ("temporal", 1, 32, 50),
                                           ("coupling", 2, 3, 37), ("coupling", 1, 2, 197)])
def test_attention_bwd(L, kind, B, T, ntok):
    _lib, ops = L
    heads = 12
    qkv = _rand(B * T * ntok, 3 * heads * 64, scale=1.2, seed=32)
    d_out = _rand(B * T * ntok, heads * 64, seed=33)
    p = _planes(qkv, ops)
    qd = _join(p).requires_grad_(True)
    _attention_ref(qd, B, T, ntok, heads, 0.125, kind.split("_")[0]).backward(d_out.double())
    d_qkv = torch.full_like(qkv, 1.0)
    scratch = torch.empty(B * heads * T * ntok * 3, device=DEV) if kind == "coupling" else None
    if kind in ("spatial_tc", "temporal_tc", "spatial_fast", "temporal_fast"):
        # tcgen05 kernels (the emulator runs contract stubs); *_fast: with the forward's row statistics, as the train step does
        if _is_emu() and B * T > 40:
            pytest.skip("bench-sized case: hardware only")
        rows = B * T * ntok
        scratch = torch.empty(rows * heads * 64 * (2 if kind.endswith("fast") else 1) + 2 * rows * heads, device=DEV)
    _lib.call("maed_bwd_attention", {"spatial": 0, "temporal": 1, "coupling": 2, "spatial_tc": 3, "temporal_tc": 4, "spatial_fast": 5, "temporal_fast": 6}[kind], _lib.ptr(p), C.c_longlong(p[0].numel()),
              _lib.ptr(d_out), B, T, ntok, heads, C.c_float(0.125), 1, _lib.ptr(d_qkv), _lib.ptr(scratch), _lib.stream_ptr())
    assert rel_err(d_qkv, qd.grad + 1.0) < 2e-5, kind


@pytest.mark.parametrize("Mo,No,R", [(768, 3072, 25216), (64, 64, 12544), (2304, 768, 1970), (64, 160, 25088), (1024, 512, 392)])
def test_wgrad_splitk(L, Mo, No, R):
    _lib, ops = L
    if _is_emu():
        pytest.skip("tcgen05 split-K kernel: hardware only (the emulator holds a contract stub)")
    ld = (R + 7) // 8 * 8
    A, B = _rand(Mo, ld, scale=0.05, seed=34), _rand(No, ld, scale=0.05, seed=35)
    pa, pb = _planes(A, ops), _planes(B, ops)
    ref = 0.5 * _join(pa)[:, :R] @ _join(pb)[:, :R].t() + 1.0
    slabs = torch.empty(_lib.load().maed_bwd_wgrad_slab_floats(Mo, No, R), device=DEV)
    D = torch.ones(Mo, No, device=DEV)
    _lib.call("maed_bwd_wgrad_splitk", _lib.ptr(pa), C.c_longlong(pa[0].numel()), ld, _lib.ptr(pb), C.c_longlong(pb[0].numel()), ld,
              Mo, No, R, 3, C.c_float(0.5), 1, _lib.ptr(slabs), _lib.ptr(D), No, _lib.stream_ptr())
    assert rel_err(D, ref) < 2e-5


@pytest.mark.parametrize("Mo,No,No_x,R", [(768, 3072, 3072, 25216), (3072, 768, 768, 25216), (64, 64, 64, 12544), (2304, 768, 768, 1970),
                                          (64, 160, 152, 25088), (1024, 512, 512, 392), (256, 576, 576, 3136), (192, 96, 72, 70)])
def test_wgrad_rows(L, Mo, No, No_x, R):
    """dW = scale * dY^T X on row-major operands (MN-major tcgen05 descriptors): what train.cu calls for every weight gradient."""
    _lib, ops = L
    if _is_emu():
        pytest.skip("tcgen05 split-K kernel: hardware only (the emulator holds a contract stub)")
    dY, X = _rand(R, Mo, scale=0.05, seed=44), _rand(R, No_x, scale=0.05, seed=45)
    py, px = _planes(dY, ops), _planes(X, ops)
    ref = torch.ones(Mo, No, dtype=torch.float64, device=DEV)
    ref[:, :No_x] += 0.5 * _join(py).t() @ _join(px)
    slabs = torch.empty(_lib.load().maed_bwd_wgrad_slab_floats(Mo, No, R), device=DEV)
    D = torch.ones(Mo, No, device=DEV)
    _lib.call("maed_bwd_wgrad_rows", _lib.ptr(py), C.c_longlong(py[0].numel()), Mo, _lib.ptr(px), C.c_longlong(px[0].numel()), No_x,
              No_x, Mo, No, R, 3, C.c_float(0.5), 1, _lib.ptr(slabs), _lib.ptr(D), No, _lib.stream_ptr())
    assert rel_err(D, ref) < 2e-5


@pytest.mark.parametrize("n,H,Cin,Cout,k", [(3, 56, 64, 64, 3), (2, 28, 128, 128, 3), (5, 14, 256, 256, 3), (2, 7, 512, 512, 3),
                                            (1, 14, 64, 128, 1)])
def test_wgrad_conv_implicit(L, n, H, Cin, Cout, k):
    """dW of a stride-1 k x k convolution with the implicit im2col operand (tap-shifted 5-D TMA patches) against the weight
    gradient of torch's conv2d in float64 ([Cout][kh][kw][Cin] layout, padding k // 2)."""
    _lib, ops = L
    if _is_emu() and n * H * H > 2000:
        pytest.skip("larger maps: hardware only (the emulator holds a contract stub)")
    x = _rand(n, H, H, Cin, seed=46)                                             # NHWC
    dy = _rand(n * H * H, Cout, scale=0.1, seed=47)
    px, py = _planes(x.reshape(-1, Cin), ops), _planes(dy, ops)
    xd = _join(px).reshape(n, H, H, Cin).permute(0, 3, 1, 2)
    wd = torch.zeros(Cout, Cin, k, k, dtype=torch.float64, device=DEV, requires_grad=True)
    y = torch.nn.functional.conv2d(xd, wd, padding=k // 2)
    y.backward(_join(py).reshape(n, H, H, Cout).permute(0, 3, 1, 2))
    ref = 0.5 * wd.grad.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin) + 1.0
    No = k * k * Cin
    slabs = torch.empty(_lib.load().maed_bwd_wgrad_slab_floats(Cout, No, n * H * H), device=DEV)
    D = torch.ones(Cout, No, device=DEV)
    _lib.call("maed_bwd_wgrad_conv", _lib.ptr(py), C.c_longlong(py[0].numel()), _lib.ptr(px), C.c_longlong(px[0].numel()), n, H, H,
              Cin, Cout, k, k, k // 2, C.c_float(0.5), 1, _lib.ptr(slabs), _lib.ptr(D), No, _lib.stream_ptr())
    assert rel_err(D, ref) < 2e-5


def test_split_transposed(L):
    _lib, _ = L
    w = _rand(2304, 768, scale=0.05, seed=36)
    out = torch.empty(2, 768, 2304, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_split_transposed", _lib.ptr(w), 2304, 768, _lib.ptr(out), C.c_longlong(out[0].numel()), _lib.stream_ptr())
    assert rel_err(_join(out), w.double().t()) < 1e-6


@pytest.mark.parametrize("stride", [1, 2])
def test_conv_dgrad_through_flipped_weights(L, stride):
    """dX of a standardised 3x3 SAME conv = implicit-GEMM conv of (dilated) dY with prep_conv_weight_dgrad's operand."""
    _lib, ops = L
    n, H, Cin, Cout = 2, 28, 64, 128
    Ho = H // stride
    w = _rand(Cout, Cin, 3, 3, scale=0.1, seed=37)
    x = _rand(n, Cin, H, H, seed=38)
    dy = _rand(n, Ho, Ho, Cout, seed=39)
    xd = x.double().requires_grad_(True)
    std, mean = torch.std_mean(w.double(), dim=[1, 2, 3], keepdim=True, unbiased=False)
    what = (w.double() - mean) / (std + 1e-5)
    pad_total = max((Ho - 1) * stride + 3 - H, 0)
    xp = F.pad(xd, [pad_total // 2, pad_total - pad_total // 2, pad_total // 2, pad_total - pad_total // 2])
    F.conv2d(xp, what, stride=stride).backward(dy.double().permute(0, 3, 1, 2))
    wt = torch.empty(2, Cin, 9 * Cout, dtype=torch.float16, device=DEV)
    _lib.call("maed_bwd_prep_conv_weight_dgrad", _lib.ptr(w), Cout, Cin, 3, 3, 1, _lib.ptr(wt), C.c_longlong(wt[0].numel()),
              _lib.stream_ptr())
    p = _planes(dy, ops)
    if stride == 2:
        dil = torch.empty(2, n, H, H, Cout, dtype=torch.float16, device=DEV)
        _lib.call("maed_bwd_dilate2", _lib.ptr(p), C.c_longlong(p[0].numel()), n, Ho, Ho, Cout, H, H, _lib.ptr(dil),
                  C.c_longlong(dil[0].numel()), _lib.stream_ptr())
        p = dil
    out = torch.empty(n, H, H, Cin, device=DEV)
    pad = 2 - pad_total // 2
    _lib.call("maed_op_conv_gemm", _lib.ptr(p), C.c_longlong(p[0].numel()), _lib.ptr(wt), C.c_longlong(wt[0].numel()), n, H, H,
              Cout, Cin, 3, 3, pad, pad, 3, 0, _lib.ptr(out), C.c_longlong(0), 0, _lib.stream_ptr())
    assert rel_err(out, xd.grad.permute(0, 2, 3, 1)) < 2e-5


def test_adam_matches_torch(L):
    _lib, _ = L
    p0, gs = _rand(10007, seed=40), [_rand(10007, seed=41 + i) for i in range(3)]
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-2, weight_decay=1e-3)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for i, g in enumerate(gs):
        ref.grad = g.clone()
        opt.step()
        _lib.call("maed_adam_step", _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(v), C.c_longlong(p.numel()), C.c_double(1e-2),
                  C.c_double(0.9), C.c_double(0.999), C.c_double(1e-8), C.c_double(1e-3), i + 1, C.c_float(1.0), _lib.stream_ptr())
    assert rel_err(p, ref.detach()) < 1e-6


def test_dropout(L):
    _lib, _ = L
    x = torch.ones(1 << 20, device=DEV)
    d = torch.ones(1 << 20, device=DEV)
    mask = torch.empty(1 << 20, dtype=torch.uint8, device=DEV)
    _lib.call("maed_bwd_dropout", _lib.ptr(x), C.c_longlong(x.numel()), C.c_float(0.5), C.c_ulonglong(1234), _lib.ptr(mask),
              _lib.ptr(d), _lib.stream_ptr())
    keep = mask.float().mean().item()
    assert abs(keep - 0.5) < 5e-3
    assert torch.equal(x, mask.float() * 2.0) and torch.equal(d, x)


# ------------------------------------------------------------------------------------ 'cnn' encoder training kernels
@pytest.mark.parametrize("M,Cc,relu,res", [(1000, 64, 1, 0), (37, 256, 0, 0), (6272, 128, 1, 1)])
def test_batchnorm_train_forward_backward(L, M, Cc, relu, res):
    """nn.BatchNorm2d.train() + (residual) + ReLU against torch autograd in float64, incl. the running-buffer update."""
    _lib, ops = L
    x = _rand(M, Cc, seed=40) * 1.7 + 0.3
    gamma, beta = _rand(Cc, seed=41) * 0.2 + 1.0, _rand(Cc, seed=42) * 0.1
    rm, rv = _rand(Cc, seed=43) * 0.1, _rand(Cc, seed=44).abs() + 0.5
    dy = _rand(M, Cc, seed=45)
    resid = _planes(_rand(M, Cc, seed=46), ops) if res else None
    rm0, rv0 = rm.clone(), rv.clone()
    y = torch.zeros(2, M, Cc, dtype=torch.float16, device=DEV)
    dx = torch.zeros_like(y)
    mean, rstd, dg, db = [torch.zeros(Cc, device=DEV) for _ in range(4)]
    lib = _lib.load()
    scratch = torch.zeros(lib.maed_bwd_batchnorm_scratch_doubles(M, Cc), dtype=torch.float64, device=DEV)
    # reference (float64 autograd): the ReLU mask is applied to dy by the caller in the engine, so compare with relu'(y) * dy
    xd = x.double().requires_grad_(True)
    gd, bd = gamma.double().requires_grad_(True), beta.double().requires_grad_(True)
    rmd, rvd = rm0.double().clone(), rv0.double().clone()
    yr = F.batch_norm(xd.reshape(M, Cc, 1, 1), rmd, rvd, gd, bd, True, 0.1, 1e-5).reshape(M, Cc)
    if res:
        yr = yr + _join(resid)
    pre = yr
    if relu:
        yr = torch.relu(yr)
    dy_eff = dy.double() * ((pre > 0).double() if relu else 1.0)
    yr_sum = (yr * dy.double()).sum()
    gx, gg, gb = torch.autograd.grad(yr_sum, [xd, gd, bd])
    dy_in = dy_eff.float().contiguous()
    _lib.call("maed_bwd_batchnorm", _lib.ptr(x), C.c_longlong(M), Cc, _lib.ptr(gamma), _lib.ptr(beta), C.c_float(1e-5),
              C.c_float(0.1), _lib.ptr(rm), _lib.ptr(rv), relu, _lib.ptr(resid), C.c_longlong(M * Cc if res else 0), _lib.ptr(y),
              C.c_longlong(M * Cc), _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dy_in), C.c_float(0.5), _lib.ptr(dg), _lib.ptr(db),
              _lib.ptr(dx), C.c_longlong(M * Cc), _lib.ptr(scratch), _lib.stream_ptr())
    assert rel_err(_join(y), yr.detach()) < 2e-6
    assert rel_err(rm, rmd) < 1e-6 and rel_err(rv, rvd) < 1e-6
    assert rel_err(_join(dx), gx) < 2e-5
    assert rel_err(dg, 0.5 * gg) < 2e-5 and rel_err(db, 0.5 * gb) < 2e-5


@pytest.mark.parametrize("n,H,W,Cc", [(2, 12, 10, 8), (1, 7, 9, 4), (3, 112, 112, 64)])
def test_maxpool3x3s2_forward_backward(L, n, H, W, Cc):
    _lib, ops = L
    if _is_emu() and H > 100:
        pytest.skip("large case: GPU only")
    x = _rand(n, H, W, Cc, seed=50) - 1.0
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    d_out = _rand(n, OH, OW, Cc, seed=51)
    out = torch.zeros(2, n, OH, OW, Cc, dtype=torch.float16, device=DEV)
    idx = torch.zeros(n, OH, OW, Cc, dtype=torch.uint8, device=DEV)
    d_x = torch.full((n, H, W, Cc), float("nan"), device=DEV)
    _lib.call("maed_bwd_maxpool3x3s2", _lib.ptr(x), n, H, W, Cc, _lib.ptr(out), C.c_longlong(out[0].numel()), _lib.ptr(idx),
              _lib.ptr(d_out), _lib.ptr(d_x), _lib.stream_ptr())
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(xr, 3, 2, 1)
    (gr,) = torch.autograd.grad((yr * d_out.double().permute(0, 3, 1, 2)).sum(), xr)
    assert rel_err(_join(out), yr.detach().permute(0, 2, 3, 1)) < 1e-6
    assert rel_err(d_x, gr.permute(0, 2, 3, 1)) < 1e-6
