"""GPU: every CUDA kernel against a plain PyTorch fp32 reference of the same op (through the C ABI)."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    assert torch.cuda.is_available()
    from maed_b200 import ops as o
    return o


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def test_split_roundtrip(ops):
    x = _rand(1000, 37, seed=1)
    p = ops.split(x)
    assert rel_err(ops.join(p), x) < 2e-7          # hi+lo carries ~22 bits
    assert torch.equal(p[0], x.half())


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 0), (300, 128, 192, 0), (25216, 768, 768, 0), (1000, 256, 152, 0),
                                      (4096, 2304, 768, 256), (512, 3072, 768, 128), (260, 768, 3072, 64)])
@pytest.mark.parametrize("nsplit", [3, 1])
def test_gemm_plain(ops, M, N, K, bn, nsplit):
    a, b = _rand(M, K, seed=2), _rand(N, K, scale=0.05, seed=3)
    pa, pb = ops.split(a), ops.split(b)
    out = ops.gemm(pa, pb, nsplit=nsplit, block_n=bn)
    if nsplit == 3:
        ref = a.double() @ b.double().t()
        assert rel_err(out, ref) < 2e-5          # fp32 tensor-core accumulation over K up to 3072
    else:
        ref = pa[0].double() @ pb[0].double().t()
        assert rel_err(out, ref) < 2e-5
        assert rel_err(out, a.double() @ b.double().t()) < 2e-3


def test_gemm_epilogues(ops):
    M, N, K = 777, 768, 256
    a, b = _rand(M, K, seed=4), _rand(N, K, scale=0.1, seed=5)
    bias, res = _rand(N, seed=6), _rand(M, N, seed=7)
    pa, pb = ops.split(a), ops.split(b)
    ref = a.double() @ b.double().t() + bias.double()
    out = ops.gemm(pa, pb, bias=bias, act=ops.ACT_GELU, out_mode=ops.OUT_F16_SPLIT)
    assert rel_err(ops.join(out), F.gelu(ref)) < 3e-6
    out = ops.gemm(pa, pb, bias=bias, residual=res)
    assert rel_err(out, ref + res.double()) < 3e-6
    out = ops.gemm(pa, pb, act=ops.ACT_RELU, out_mode=ops.OUT_F16)
    assert rel_err(out[0].float(), F.relu(a.double() @ b.double().t())) < 1e-3
    # in-place residual (out aliases residual), as used for the STE residual stream
    x = res.clone()
    from maed_b200._lib import call, ptr, stream_ptr
    call("maed_op_gemm", ptr(pa), pa[0].numel(), K, ptr(pb), pb[0].numel(), K, M, N, K, 3, ptr(bias), ptr(x), 0, 0, ptr(x), 0,
         N, 0, stream_ptr())
    assert rel_err(x, ref + res.double()) < 3e-6


@pytest.mark.parametrize("n,H,Cin,Cout,k", [(3, 56, 64, 64, 3), (2, 28, 128, 128, 3), (5, 14, 256, 256, 3), (2, 14, 64, 128, 1)])
def test_conv_implicit_gemm(ops, n, H, Cin, Cout, k):
    x = _rand(n, Cin, H, H, seed=8)
    w = _rand(Cout, Cin, k, k, scale=0.1, seed=9)
    a = ops.split(x.permute(0, 2, 3, 1).contiguous())
    wp = ops.split(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous())
    out = ops.conv_gemm(a, wp, k, k, k // 2, k // 2)
    ref = F.conv2d(x.double(), w.double(), padding=k // 2).permute(0, 2, 3, 1).reshape(-1, Cout)
    assert rel_err(out, ref) < 2e-5          # fp32 tensor-core accumulation over K up to 2304


def test_prep_conv_weight_standardizes(ops):
    w = _rand(64, 32, 3, 3, scale=0.3, seed=10) + 0.05
    p = ops.prep_conv_weight(w)
    std, mean = torch.std_mean(w, dim=[1, 2, 3], keepdim=True, unbiased=False)
    ref = ((w - mean) / (std + 1e-5)).permute(0, 2, 3, 1).reshape(64, -1)
    assert rel_err(ops.join(p), ref) < 1e-6
    p2 = ops.prep_conv_weight(_rand(64, 3, 7, 7, seed=11), k_pad=152)
    assert p2.shape == (2, 64, 152) and float(p2[:, :, 147:].abs().max()) == 0.0


def test_im2col_stem_and_stem_conv(ops):
    x = _rand(2, 3, 224, 224, seed=12)
    w = _rand(64, 3, 7, 7, scale=0.2, seed=13)
    col = ops.im2col_stem(x)
    xp = F.pad(x, [2, 3, 2, 3])
    ref_col = F.unfold(xp, 7, stride=2).reshape(2, 3, 49, -1).permute(0, 3, 2, 1).reshape(-1, 147)   # (m, tap, c)
    assert rel_err(ops.join(col)[:, :147], ref_col) < 2e-7
    wp = ops.prep_conv_weight(w, k_pad=152, standardize=False)
    out = ops.gemm(col, wp)
    ref = F.conv2d(xp.double(), w.double(), stride=2).permute(0, 2, 3, 1).reshape(-1, 64)
    assert rel_err(out, ref) < 3e-6


def test_stem_conv_fused(ops):
    """tcgen05 stem conv with in-kernel im2col + fused GroupNorm statistics vs F.conv2d (SAME pad 2/3)."""
    n = 3
    x = _rand(n, 3, 224, 224, seed=40)
    w = _rand(64, 3, 7, 7, scale=0.2, seed=41)
    wp = ops.prep_conv_weight(w, k_pad=152, standardize=False)
    out, stats = ops.stem_conv(x, wp)
    ref = F.conv2d(F.pad(x, [2, 3, 2, 3]).double(), w.double(), stride=2).permute(0, 2, 3, 1)     # (n,112,112,64)
    assert rel_err(out.reshape(n, 112, 112, 64), ref) < 3e-6
    g = ref.reshape(n, 112 * 112, 32, 2)
    assert rel_err(stats[:, :, 0], g.sum(dim=(1, 3))) < 1e-5
    assert rel_err(stats[:, :, 1], (g * g).sum(dim=(1, 3))) < 1e-5


@pytest.mark.parametrize("H,C,k,s", [(56, 128, 3, 2), (56, 256, 1, 2), (28, 256, 3, 2)])
def test_im2col_nhwc(ops, H, C, k, s):
    x = _rand(2, C, H, H, seed=14)
    a = ops.split(x.permute(0, 2, 3, 1).contiguous())
    OH = H // s
    pad = max((OH - 1) * s + k - H, 0)
    col = ops.im2col_nhwc(a, k, k, s, pad // 2, pad // 2, OH, OH)
    xp = F.pad(x, [pad // 2, pad - pad // 2, pad // 2, pad - pad // 2])
    ref = F.unfold(xp, k, stride=s).reshape(2, C, k * k, -1).permute(0, 3, 2, 1).reshape(-1, k * k * C)
    assert rel_err(ops.join(col), ref) < 2e-7


@pytest.mark.parametrize("HW,C,relu,res", [(3136, 64, True, False), (3136, 256, True, True), (784, 512, False, False),
                                           (196, 1024, True, True), (196, 256, True, False)])
def test_groupnorm(ops, HW, C, relu, res):
    n = 3
    x = _rand(n, HW, C, scale=2.0, seed=15) + 0.3
    gamma, beta = 1 + 0.1 * _rand(C, seed=16), 0.1 * _rand(C, seed=17)
    r = _rand(n, HW, C, seed=18) if res else None
    rp = ops.split(r) if res else None
    out = ops.join(ops.groupnorm(x, gamma, beta, relu, rp))
    ref = F.group_norm(x.permute(0, 2, 1).double(), 32, gamma.double(), beta.double(), 1e-5).permute(0, 2, 1)
    if res:
        ref = ref + r.double()
    if relu:
        ref = F.relu(ref)
    assert rel_err(out, ref) < 1e-6


@pytest.mark.parametrize("n,H,Cin,C,k,relu,res", [(5, 56, 64, 256, 1, True, True), (3, 56, 64, 64, 3, True, False),
                                                  (6, 28, 128, 512, 1, False, False), (4, 28, 128, 128, 3, True, False),
                                                  (9, 14, 256, 1024, 1, True, True), (7, 14, 256, 256, 3, True, False),
                                                  (2, 56, 256, 128, 1, True, False), (40, 14, 1024, 256, 1, True, False),
                                                  # many images: every cluster works through several items per epilogue group, so
                                                  # the slot / barrier phases that carry over from item to item are exercised
                                                  (48, 28, 128, 512, 1, True, True), (40, 14, 256, 1024, 1, True, True),
                                                  (24, 56, 64, 256, 1, True, True), (64, 14, 256, 256, 3, True, False)])
def test_conv_gn_fused(ops, n, H, Cin, C, k, relu, res):
    """Fused tcgen05 conv + GroupNorm (+shortcut, ReLU), incl. the 4- and 8-CTA cluster (DSMEM) variants, vs PyTorch."""
    x = _rand(n, Cin, H, H, seed=50)
    w = _rand(C, Cin, k, k, scale=0.1, seed=51)
    gamma, beta = 1 + 0.1 * _rand(C, seed=52), 0.1 * _rand(C, seed=53)
    r = _rand(n, C, H, H, seed=54) if res else None
    a = ops.split(x.permute(0, 2, 3, 1).contiguous())
    wp = ops.split(w.permute(0, 2, 3, 1).reshape(C, -1).contiguous())
    rp = ops.split(r.permute(0, 2, 3, 1).contiguous()) if res else None
    out = ops.join(ops.conv_gn(a, wp, k, k, gamma, beta, relu, rp))
    y = F.group_norm(F.conv2d(x.double(), w.double(), padding=k // 2), 32, gamma.double(), beta.double(), 1e-5)
    if res:
        y = y + r.double()
    if relu:
        y = F.relu(y)
    assert rel_err(out, y.permute(0, 2, 3, 1)) < 2e-5


def test_groupnorm_maxpool(ops):
    n, H, C = 2, 112, 64
    x = _rand(n, H, H, C, scale=1.5, seed=19)
    gamma, beta = 1 + 0.1 * _rand(C, seed=20), 0.1 * _rand(C, seed=21)
    out = ops.join(ops.groupnorm_maxpool(x, gamma, beta))
    y = F.relu(F.group_norm(x.permute(0, 3, 1, 2).double(), 32, gamma.double(), beta.double(), 1e-5))
    y = F.max_pool2d(F.pad(y, [0, 1, 0, 1], value=-float("inf")), 3, 2)
    assert rel_err(out, y.permute(0, 2, 3, 1)) < 1e-6


def test_layernorm(ops):
    x = _rand(1000, 768, scale=3.0, seed=22) + 0.5
    gamma, beta = 1 + 0.1 * _rand(768, seed=23), 0.1 * _rand(768, seed=24)
    out = ops.join(ops.layernorm(x, gamma, beta))
    assert rel_err(out, F.layer_norm(x.double(), (768,), gamma.double(), beta.double(), 1e-6)) < 1e-6


def _ref_attention(qkv, B, T, ntok, heads, scale, kind):
    BT = B * T
    q, k, v = qkv.double().reshape(BT, ntok, 3, heads, 64).permute(2, 0, 3, 1, 4)
    if kind == "spatial":
        a = (q @ k.transpose(-2, -1) * scale).softmax(-1)
        return (a @ v).transpose(1, 2).reshape(BT * ntok, heads * 64)
    if kind == "temporal":
        r = lambda t: t.reshape(B, T, heads, ntok, 64).permute(0, 2, 3, 1, 4)  # noqa: E731
        a = (r(q) @ r(k).transpose(-2, -1) * scale).softmax(-1)
        return (a @ r(v)).permute(0, 3, 2, 1, 4).reshape(BT * ntok, heads * 64)
    r = lambda t: t.reshape(B, T, heads, ntok, 64).transpose(1, 2).reshape(B, heads, T * ntok, 64)  # noqa: E731
    a = (r(q) @ r(k).transpose(-2, -1) * scale).softmax(-1)
    o = (a @ r(v)).reshape(B, heads, T, ntok, 64).transpose(1, 2).reshape(BT, heads, ntok, 64)
    return o.transpose(1, 2).reshape(BT * ntok, heads * 64)


@pytest.mark.parametrize("kind,B,T,ntok", [("generic", 2, 3, 197), ("temporal", 2, 16, 197), ("temporal", 3, 8, 50),
                                           ("temporal", 2, 5, 33), ("temporal", 1, 32, 64), ("temporal", 4, 1, 20),
                                           ("spatial", 2, 3, 197), ("spatial", 40, 4, 197), ("spatial", 1, 2, 100),
                                           ("spatial", 1, 1, 197)])
def test_attention(ops, kind, B, T, ntok):
    heads = 12
    qkv = _rand(B * T * ntok, 3 * heads * 64, scale=1.5, seed=25)
    out = ops.attention(kind, ops.split(qkv), B, T, ntok, heads, 0.125)
    ref = _ref_attention(qkv, B, T, ntok, heads, 0.125, kind)
    assert rel_err(out, ref) < 5e-6, kind


@pytest.mark.parametrize("kind,B,T,ntok", [("spatial", 3, 4, 197), ("spatial", 20, 8, 197), ("spatial", 1, 3, 128),
                                           ("spatial", 2, 2, 60), ("spatial", 1, 2, 129), ("temporal", 2, 16, 197)])
def test_attention_plane_output(ops, kind, B, T, ntok):
    """The fp16 hi/lo plane output (what the engine consumes; the spatial kernel writes it by TMA store)."""
    heads = 12
    qkv = _rand(B * T * ntok, 3 * heads * 64, scale=1.5, seed=32)
    out = ops.attention(kind, ops.split(qkv), B, T, ntok, heads, 0.125, planes=True)
    ref = _ref_attention(qkv, B, T, ntok, heads, 0.125, kind)
    assert torch.isfinite(out.float()).all()
    assert rel_err(out[0].double() + out[1].double(), ref) < 5e-6, kind


def test_attention_spatial_plain_fp16(ops):
    B, T, ntok, heads = 2, 2, 197, 12
    qkv = _rand(B * T * ntok, 3 * heads * 64, scale=1.0, seed=26)
    out = ops.attention("spatial", ops.split(qkv), B, T, ntok, heads, 0.125, nsplit=1)
    assert rel_err(out, _ref_attention(qkv, B, T, ntok, heads, 0.125, "spatial")) < 3e-3


def test_linear_f32(ops):
    x, W, b = _rand(130, 925, seed=27), _rand(1024, 925, scale=0.05, seed=28), _rand(1024, seed=29)
    out = ops.linear_f32(x, W, b, act=ops.ACT_TANH)
    assert rel_err(out, torch.tanh(x.double() @ W.double().t() + b.double())) < 1e-6
    r = _rand(130, 10, seed=30)
    out = ops.linear_f32(x, W[:10].contiguous(), b[:10].contiguous(), residual=r)
    assert rel_err(out, x.double() @ W[:10].double().t() + b[:10].double() + r.double()) < 1e-6


def test_decode_outputs_matches_oracle(ops):
    from oracle import maed_oracle as O
    g = torch.Generator().manual_seed(31)
    pose6d = torch.randn(200, 144, generator=g)
    pose6d[0] = torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24)          # identity
    pose6d[1, :6] = torch.tensor([-1., 0., 0., -1., 0., 0.])              # 180 degrees about z (branchy quaternion case)
    shape, cam = torch.randn(200, 10, generator=g), torch.randn(200, 3, generator=g) + 1.0
    rot, theta, kp2d = ops.decode_outputs(pose6d.cuda(), shape.cuda(), cam.cuda())
    ref = O.decode_outputs(pose6d, shape, cam)
    assert rel_err(rot, ref["rotmat"]) < 1e-6
    assert rel_err(theta[:, :3], ref["theta"][:, :3]) == 0.0
    assert rel_err(theta[:, 75:], ref["theta"][:, 75:]) == 0.0
    assert rel_err(theta[:, 3:75], ref["theta"][:, 3:75]) < 1e-5
    assert rel_err(kp2d, ref["kp_2d"]) < 1e-6
