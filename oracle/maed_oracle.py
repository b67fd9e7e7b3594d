"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32, plain PyTorch ops) of MAED's per-clip forward.

Nothing in the shipped product (`maed_b200/`) imports this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may.

It restates, function by function, what the reference computes for
``MAED(encoder in {'ste','cnn'}, ..., decoder in {'ktd','iterative'}).forward`` (reference `lib/models/maed.py:52-66`;
the 'cnn' encoder is torchvision's ResNet-50, restated from torchvision 0.26's `models/resnet.py`)
from a plain ``state_dict`` with the REFERENCE's key names, so it can run on the GPU box where
`/root/reference` does not exist.  It is pinned against the reference's own modules executed on CPU:
``tests/golden/*.npz`` were produced by ``tests/golden/make_golden.py`` from the unmodified reference
(through `oracle/ref_shim.py`) and ``tests/test_oracle.py`` checks this file against them.

Parity status: pinned for encoder features, pose6d/shape/cam, rotmat, theta.  **Parity unpinned** for
verts / kp_3d / kp_2d: they depend on `smplx==0.1.13` (requirements.txt:4, un-vendored, absent) and on
licensed SMPL assets; this oracle returns zeros for `verts`/`kp_3d` exactly like the shimmed reference.

`gemm_in` is an optional hook applied to every tensor-core operand (activations and weights of convs,
linears and attention matmuls).  It is used to *predict* the effect of fp16/bf16 operand rounding on
the 1e-3 parity gate; the default (identity) is the fp32 oracle.
"""
import math

import torch
import torch.nn.functional as F

# reference lib/models/ktd.py:10-35 — kinematic ancestors of each of the 24 SMPL joints
ANCESTORS = [
    [], [0], [0], [0], [0, 1], [0, 2], [0, 3], [0, 1, 4], [0, 2, 5], [0, 3, 6], [0, 1, 4, 7],
    [0, 2, 5, 8], [0, 3, 6, 9], [0, 3, 6, 9], [0, 3, 6, 9], [0, 3, 6, 9, 12], [0, 3, 6, 9, 13],
    [0, 3, 6, 9, 14], [0, 3, 6, 9, 13, 16], [0, 3, 6, 9, 14, 17], [0, 3, 6, 9, 13, 16, 18],
    [0, 3, 6, 9, 14, 17, 19], [0, 3, 6, 9, 13, 16, 18, 20], [0, 3, 6, 9, 14, 17, 19, 21],
]

_ident = lambda t: t  # noqa: E731


# --------------------------------------------------------------------------------------------------
# backbone: ResNetV2 (3,4,9), StdConv2dSame + GroupNorm(32) (+ReLU)      reference lib/models/resnetv2.py
# --------------------------------------------------------------------------------------------------
def same_pad_amount(size, k, s):
    """resnetv2.py:51-52 get_same_padding (dilation 1)."""
    return max((math.ceil(size / s) - 1) * s + (k - 1) + 1 - size, 0)


def pad_same(x, k, s, value=0.0):
    """resnetv2.py:54-59: TF 'SAME' padding, extra pixel goes bottom/right."""
    ph = same_pad_amount(x.shape[-2], k, s)
    pw = same_pad_amount(x.shape[-1], k, s)
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2], value=value)
    return x


def standardize_weight(w, eps=1e-5):
    """resnetv2.py:86-89: per-output-channel (w-mean)/(std_biased+eps); eps is added to std, not var."""
    std, mean = torch.std_mean(w, dim=[1, 2, 3], keepdim=True, unbiased=False)
    return (w - mean) / (std + eps)


def std_conv_same(x, w, stride, gemm_in=_ident):
    """resnetv2.py:91-93."""
    k = w.shape[-1]
    return F.conv2d(gemm_in(pad_same(x, k, stride)), gemm_in(standardize_weight(w)), None, stride)


def group_norm_act(x, w, b, relu):
    """resnetv2.py:45-49 (32 groups, eps 1e-5)."""
    x = F.group_norm(x, 32, w, b, 1e-5)
    return F.relu(x) if relu else x


def bottleneck(x, sd, p, stride, has_ds, gemm_in=_ident):
    """resnetv2.py:189-204 (non pre-activation Bottleneck; stride on conv2 and on the 1x1 downsample)."""
    sc = x
    if has_ds:
        sc = std_conv_same(x, sd[p + "downsample.conv.weight"], stride, gemm_in)
        sc = group_norm_act(sc, sd[p + "downsample.norm.weight"], sd[p + "downsample.norm.bias"], False)
    y = std_conv_same(x, sd[p + "conv1.weight"], 1, gemm_in)
    y = group_norm_act(y, sd[p + "norm1.weight"], sd[p + "norm1.bias"], True)
    y = std_conv_same(y, sd[p + "conv2.weight"], stride, gemm_in)
    y = group_norm_act(y, sd[p + "norm2.weight"], sd[p + "norm2.bias"], True)
    y = std_conv_same(y, sd[p + "conv3.weight"], 1, gemm_in)
    y = group_norm_act(y, sd[p + "norm3.weight"], sd[p + "norm3.bias"], False)
    return F.relu(y + sc)


def backbone(x, sd, pre="encoder.patch_embed.backbone.", gemm_in=_ident, taps=None):
    """resnetv2.py:245-274 (stem, 'same'), :218-242 (stages), :337-348."""
    y = std_conv_same(x, sd[pre + "stem.conv.weight"], 2, gemm_in)
    y = group_norm_act(y, sd[pre + "stem.norm.weight"], sd[pre + "stem.norm.bias"], True)
    y = F.max_pool2d(pad_same(y, 3, 2, value=-float("inf")), 3, 2)            # resnetv2.py:70-72
    if taps is not None:
        taps["stem"] = y
    for si, depth in enumerate((3, 4, 9)):
        for bi in range(depth):
            stride = 2 if (si > 0 and bi == 0) else 1
            y = bottleneck(y, sd, "%sstages.%d.blocks.%d." % (pre, si, bi), stride, bi == 0, gemm_in)
        if taps is not None:
            taps["stage%d" % si] = y
    return y


# --------------------------------------------------------------------------------------------------
# STE (hybrid ViT)                                             reference lib/models/vision_transformer.py
# --------------------------------------------------------------------------------------------------
def attn_spatial(q, k, v, scale, gemm_in=_ident):
    """vision_transformer.py:206-214.  q,k,v: (BT, H, N, d)."""
    a = (gemm_in(q) @ gemm_in(k).transpose(-2, -1)) * scale
    a = a.softmax(dim=-1)
    o = gemm_in(a) @ gemm_in(v)
    BT, H, N, d = q.shape
    return o.transpose(1, 2).reshape(BT, N, H * d)


def attn_temporal(q, k, v, scale, T, gemm_in=_ident):
    """vision_transformer.py:216-228: attention across the T frames of a clip, per (clip, head, token)."""
    BT, H, N, d = q.shape
    r = lambda t: t.reshape(-1, T, H, N, d).permute(0, 2, 3, 1, 4)  # noqa: E731  (B,H,N,T,d)
    qt, kt, vt = r(q), r(k), r(v)
    a = (gemm_in(qt) @ gemm_in(kt).transpose(-2, -1)) * scale
    a = a.softmax(dim=-1)
    o = gemm_in(a) @ gemm_in(vt)
    return o.permute(0, 3, 2, 1, 4).reshape(BT, N, H * d)


def attn_coupling(q, k, v, scale, T, gemm_in=_ident):
    """vision_transformer.py:180-204: joint attention over the T*N tokens of a clip."""
    BT, H, N, d = q.shape
    r = lambda t: t.reshape(-1, T, H, N, d).transpose(1, 2).reshape(-1, H, T * N, d)  # noqa: E731
    qc, kc, vc = r(q), r(k), r(v)
    a = (gemm_in(qc) @ gemm_in(kc).transpose(-2, -1)) * scale
    a = a.softmax(dim=-1)
    o = gemm_in(a) @ gemm_in(vc)                                   # (B,H,TN,d)
    o = o.reshape(-1, H, T, N, d).transpose(1, 2).reshape(BT, H, N, d)
    return o.transpose(1, 2).reshape(BT, N, H * d)


def linear(x, sd, p, gemm_in=_ident):
    return F.linear(gemm_in(x), gemm_in(sd[p + "weight"]), sd.get(p + "bias"))


def attention(x, sd, p, mode, T, H, gemm_in=_ident):
    """vision_transformer.py:135-178."""
    BT, N, C = x.shape
    d = C // H
    scale = d ** -0.5

    def qkv_of(t):
        n = t.shape[1]
        y = linear(t, sd, p + "qkv.", gemm_in).reshape(BT, n, 3, H, d).permute(2, 0, 3, 1, 4)
        return y[0], y[1], y[2]

    if mode == "series":                        # the SAME qkv weights are applied twice (:139-145)
        q, k, v = qkv_of(x)
        x = attn_spatial(q, k, v, scale, gemm_in)
        q, k, v = qkv_of(x)
        x = attn_temporal(q, k, v, scale, T, gemm_in)
    elif mode == "parallel":                    # :146-158
        q, k, v = qkv_of(x)
        x_t = attn_temporal(q, k, v, scale, T, gemm_in)
        x_s = attn_spatial(q, k, v, scale, gemm_in)
        alpha = torch.cat([x_s, x_t], dim=-1).mean(dim=1, keepdim=True)
        alpha = linear(alpha, sd, p + "ts_attn.", gemm_in).reshape(BT, 1, C, 2).softmax(dim=-1)
        x = x_t * alpha[..., 1] + x_s * alpha[..., 0]
    elif mode == "coupling":                    # :159-162
        q, k, v = qkv_of(x)
        x = attn_coupling(q, k, v, scale, T, gemm_in)
    elif mode == "vanilla":                     # :163-166
        q, k, v = qkv_of(x)
        x = attn_spatial(q, k, v, scale, gemm_in)
    elif mode == "temporal":                    # :167-173  (token mean first => N=1)
        x = x.mean(dim=1, keepdim=True)
        q, k, v = qkv_of(x)
        x = attn_temporal(q, k, v, scale, T, gemm_in)
    else:
        raise NotImplementedError(mode)
    return linear(x, sd, p + "proj.", gemm_in)


def ste_block(x, sd, p, mode, T, H, gemm_in=_ident):
    """vision_transformer.py:258-261 (pre-LN, eps 1e-6, DropPath = identity) and Mlp :105-112."""
    C = x.shape[-1]
    y = F.layer_norm(x, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)
    x = x + attention(y, sd, p + "attn.", mode, T, H, gemm_in)
    y = F.layer_norm(x, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)
    y = F.gelu(linear(y, sd, p + "mlp.fc1.", gemm_in))              # exact-erf GELU (nn.GELU default)
    return x + linear(y, sd, p + "mlp.fc2.", gemm_in)


def ste_encoder(x, sd, mode, T, num_blocks=6, H=12, gemm_in=_ident, taps=None):
    """vision_transformer.py:388-412 with HybridEmbed :306-311.  x: (BT,3,224,224) -> (BT,768)."""
    pre = "encoder."
    BT = x.shape[0]
    f = backbone(x, sd, pre + "patch_embed.backbone.", gemm_in, taps)
    f = F.conv2d(gemm_in(f), gemm_in(sd[pre + "patch_embed.proj.weight"]), sd[pre + "patch_embed.proj.bias"])
    tok = f.flatten(2).transpose(1, 2)                                         # (BT,196,768)
    tok = torch.cat([sd[pre + "cls_token"].expand(BT, -1, -1), tok], dim=1) + sd[pre + "pos_embed"]
    if mode in ("coupling", "parallel", "series"):                             # :396-399
        _, N, C = tok.shape
        tok = (tok.reshape(-1, T, N, C) + sd[pre + "temp_embed"][:, :T]).reshape(BT, N, C)
    if taps is not None:
        taps["embed"] = tok
    for i in range(num_blocks):
        tok = ste_block(tok, sd, "%sblocks.%d." % (pre, i), mode, T, H, gemm_in)
        if taps is not None:
            taps["block%d" % i] = tok
    C = tok.shape[-1]
    cls = F.layer_norm(tok, (C,), sd[pre + "norm.weight"], sd[pre + "norm.bias"], 1e-6)[:, 0]
    return torch.tanh(linear(cls, sd, pre + "pre_logits.fc.", gemm_in))        # :350-353


# --------------------------------------------------------------------------------------------------
# 'cnn' encoder: torchvision ResNet-50 with fc = Identity                 reference lib/models/maed.py:35-37
# --------------------------------------------------------------------------------------------------
def batch_norm_eval(x, sd, p):
    """nn.BatchNorm2d: eval() -> (x - running_mean) / sqrt(running_var + 1e-5) * weight + bias; with sd["__bn_train__"] set
    (training-path oracle) -> batch statistics, and the running buffers in `sd` are updated in place (momentum 0.1, unbiased
    variance) like nn.BatchNorm2d.train() does."""
    train = bool(sd.get("__bn_train__", False))
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"], train, 0.1, 1e-5)


def tv_bottleneck(x, sd, p, stride, has_ds, gemm_in=_ident):
    """torchvision.models.resnet.Bottleneck (v1.5: the stride sits on the 3x3 conv2); expansion 4."""
    idt = x
    if has_ds:
        idt = F.conv2d(gemm_in(x), gemm_in(sd[p + "downsample.0.weight"]), None, stride)
        idt = batch_norm_eval(idt, sd, p + "downsample.1.")
    y = F.relu(batch_norm_eval(F.conv2d(gemm_in(x), gemm_in(sd[p + "conv1.weight"])), sd, p + "bn1."))
    y = F.relu(batch_norm_eval(F.conv2d(gemm_in(y), gemm_in(sd[p + "conv2.weight"]), None, stride, 1), sd, p + "bn2."))
    y = batch_norm_eval(F.conv2d(gemm_in(y), gemm_in(sd[p + "conv3.weight"])), sd, p + "bn3.")
    return F.relu(y + idt)


def cnn_encoder(x, sd, gemm_in=_ident, taps=None):
    """torchvision resnet50._forward_impl with fc = nn.Identity() (maed.py:36-37): (BT,3,224,224) -> (BT,2048)."""
    pre = "encoder."
    y = F.conv2d(gemm_in(x), gemm_in(sd[pre + "conv1.weight"]), None, 2, 3)
    y = F.relu(batch_norm_eval(y, sd, pre + "bn1."))
    y = F.max_pool2d(y, 3, 2, 1)
    if taps is not None:
        taps["stem"] = y
    for li, depth in enumerate((3, 4, 6, 3)):
        for bi in range(depth):
            stride = 2 if (li > 0 and bi == 0) else 1
            y = tv_bottleneck(y, sd, "%slayer%d.%d." % (pre, li + 1, bi), stride, bi == 0, gemm_in)
        if taps is not None:
            taps["stage%d" % li] = y
    return y.mean(dim=(2, 3))                                                  # AdaptiveAvgPool2d(1) + flatten


# --------------------------------------------------------------------------------------------------
# decoders                                       reference lib/models/ktd.py, lib/models/spin.py
# --------------------------------------------------------------------------------------------------
def ktd_regress(xf, sd, gemm_in=_ident):
    """ktd.py:69-88 in eval() (dropout = identity): fc1, fc2 (no nonlinearity), shape, cam, then the 24
    tree-ordered joint regressors, each fed its ancestors' 6-D outputs."""
    p = "decoder."
    x = linear(linear(xf, sd, p + "fc1.", gemm_in), sd, p + "fc2.", gemm_in)
    shape = linear(x, sd, p + "decshape.", gemm_in)
    cam = linear(x, sd, p + "deccam.", gemm_in)
    pose = []
    for j, anc in enumerate(ANCESTORS):
        inp = torch.cat([x] + [pose[a] for a in anc], dim=1)
        pose.append(linear(inp, sd, "%sjoint_regs.%d." % (p, j), gemm_in))
    return torch.cat(pose, dim=1), shape, cam


def iterative_regress(xf, sd, n_iter=3, gemm_in=_ident):
    """spin.py:51-74 in eval()."""
    p = "decoder."
    nt = xf.shape[0]
    pose = sd[p + "init_pose"].expand(nt, -1)
    shape = sd[p + "init_shape"].expand(nt, -1)
    cam = sd[p + "init_cam"].expand(nt, -1)
    for _ in range(n_iter):
        xc = torch.cat([xf, pose, shape, cam], 1)
        xc = linear(linear(xc, sd, p + "fc1.", gemm_in), sd, p + "fc2.", gemm_in)
        pose = linear(xc, sd, p + "decpose.", gemm_in) + pose
        shape = linear(xc, sd, p + "decshape.", gemm_in) + shape
        cam = linear(xc, sd, p + "deccam.", gemm_in) + cam
    return pose, shape, cam


def rot6d_to_rotmat(x):
    """lib/utils/geometry.py:320-334: Gram-Schmidt on the two columns of x.view(-1,3,2)."""
    x = x.reshape(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = a1 / a1.norm(dim=1, keepdim=True).clamp_min(1e-6)
    u = a2 - (b1 * a2).sum(dim=1, keepdim=True) * b1
    b2 = u / u.norm(dim=1, keepdim=True).clamp_min(1e-6)
    b3 = torch.cross(b1, b2, dim=1)
    return torch.stack([b1, b2, b3], dim=-1)


def rotmat_to_angle_axis(R):
    """geometry.py:58-87 -> :143-223 (R -> quaternion, 4 masked cases on the TRANSPOSED matrix, eps 1e-6)
    -> :90-140 (quaternion -> angle-axis via atan2), NaN -> 0."""
    m = R.reshape(-1, 3, 3).transpose(1, 2)         # rmat_t
    m00, m01, m02 = m[:, 0, 0], m[:, 0, 1], m[:, 0, 2]
    m10, m11, m12 = m[:, 1, 0], m[:, 1, 1], m[:, 1, 2]
    m20, m21, m22 = m[:, 2, 0], m[:, 2, 1], m[:, 2, 2]
    d2 = m22 < 1e-6
    d0_d1 = m00 > m11
    d0_nd1 = m00 < -m11
    t0 = 1 + m00 - m11 - m22
    t1 = 1 - m00 + m11 - m22
    t2 = 1 - m00 - m11 + m22
    t3 = 1 + m00 + m11 + m22
    q0 = torch.stack([m12 - m21, t0, m01 + m10, m20 + m02], -1)
    q1 = torch.stack([m20 - m02, m01 + m10, t1, m12 + m21], -1)
    q2 = torch.stack([m01 - m10, m20 + m02, m12 + m21, t2], -1)
    q3 = torch.stack([t3, m12 - m21, m20 - m02, m01 - m10], -1)
    c0 = (d2 & d0_d1).unsqueeze(-1).float()
    c1 = (d2 & ~d0_d1).unsqueeze(-1).float()
    c2 = (~d2 & d0_nd1).unsqueeze(-1).float()
    c3 = (~d2 & ~d0_nd1).unsqueeze(-1).float()
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    t = t0.unsqueeze(-1) * c0 + t1.unsqueeze(-1) * c1 + t2.unsqueeze(-1) * c2 + t3.unsqueeze(-1) * c3
    q = 0.5 * q / torch.sqrt(t)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s2 = x * x + y * y + z * z
    s = torch.sqrt(s2)
    two_theta = 2.0 * torch.where(w < 0.0, torch.atan2(-s, -w), torch.atan2(s, w))
    kk = torch.where(s2 > 0.0, two_theta / s, torch.full_like(s, 2.0))
    aa = torch.stack([x * kk, y * kk, z * kk], -1)
    return torch.nan_to_num(aa, nan=0.0, posinf=float("inf"), neginf=-float("inf"))


def project_keypoints(joints, cam):
    """spin.py:113-157: weak-perspective cam -> translation, perspective projection f=5000, /112."""
    t = torch.stack([cam[:, 1], cam[:, 2], 2 * 5000.0 / (224.0 * cam[:, 0] + 1e-9)], dim=-1)
    pts = joints + t.unsqueeze(1)
    pts = pts / pts[:, :, -1:].clone()
    return (5000.0 * pts[:, :, :2]) / 112.0


def decode_outputs(pose6d, shape, cam, n_joints=49):
    """ktd.py:94-124 / spin.py:87-110 with the placeholder body model (verts = joints = 0)."""
    nt = pose6d.shape[0]
    R = rot6d_to_rotmat(pose6d).reshape(nt, 24, 3, 3)
    verts = pose6d.new_zeros(nt, 6890, 3)
    kp3d = pose6d.new_zeros(nt, n_joints, 3)
    kp2d = project_keypoints(kp3d, cam)
    aa = rotmat_to_angle_axis(R.reshape(-1, 3, 3)).reshape(nt, 72)
    theta = torch.cat([cam, aa, shape], dim=1)
    return {"theta": theta, "verts": verts, "kp_2d": kp2d, "kp_3d": kp3d, "rotmat": R}


# --------------------------------------------------------------------------------------------------
def maed_forward(x, sd, st_mode="parallel", decoder="ktd", num_blocks=6, num_heads=12,
                 gemm_in=_ident, taps=None, encoder="ste"):
    """lib/models/maed.py:52-66.  x: (N,T,3,224,224) fp32 -> dict like the reference's."""
    N, T = x.shape[:2]
    if encoder == "cnn":
        xf = cnn_encoder(x.reshape(N * T, *x.shape[2:]), sd, gemm_in, taps)
    else:
        xf = ste_encoder(x.reshape(N * T, *x.shape[2:]), sd, st_mode, T, num_blocks, num_heads, gemm_in, taps)
    if decoder == "ktd":
        pose6d, shape, cam = ktd_regress(xf, sd, gemm_in)
    elif decoder == "iterative":
        pose6d, shape, cam = iterative_regress(xf, sd, 3, gemm_in)
    else:
        raise NotImplementedError(decoder)
    out = decode_outputs(pose6d, shape, cam)
    if taps is not None:
        taps.update(feat=xf, pose6d=pose6d, shape=shape, cam=cam)
    return {
        "theta": out["theta"].reshape(N, T, -1), "verts": out["verts"].reshape(N, T, -1, 3),
        "kp_2d": out["kp_2d"].reshape(N, T, -1, 2), "kp_3d": out["kp_3d"].reshape(N, T, -1, 3),
        "rotmat": out["rotmat"].reshape(N, T, -1, 3, 3),
    }


# --------------------------------------------------------------------------------------------------
# gradients (training-path oracle): autograd over the restatement above
# --------------------------------------------------------------------------------------------------
def maed_param_grads(x, sd, probe_pose, probe_shape, probe_cam, st_mode="parallel", decoder="ktd", num_blocks=6,
                     num_heads=12, encoder="ste"):
    """dL/dp for every entry of `sd` with L = sum(pose6d*A) + sum(shape*B) + sum(cam*C) — the scalar
    tests/golden/make_golden_grads.py back-propagates through the reference (eval mode: no dropout).
    Returns (L, {key: grad}, {pose6d, shape, cam})."""
    is_buf = lambda k: k.endswith(("running_mean", "running_var", "num_batches_tracked"))  # noqa: E731
    sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not is_buf(k)) for k, v in sd.items()}
    if encoder == "cnn":
        sd["__bn_train__"] = torch.tensor(1)        # train()-mode BatchNorm (the running buffers of this copy get updated)
    taps = {}
    maed_forward(x, sd, st_mode, decoder, num_blocks, num_heads, taps=taps, encoder=encoder)
    L = (taps["pose6d"] * probe_pose).sum() + (taps["shape"] * probe_shape).sum() + (taps["cam"] * probe_cam).sum()
    keys = [k for k, v in sd.items() if v.requires_grad]
    grads = torch.autograd.grad(L, [sd[k] for k in keys], allow_unused=True)
    outs = {k: taps[k].detach() for k in ("pose6d", "shape", "cam")}
    if encoder == "cnn":
        outs["buffers"] = {k: v.detach() for k, v in sd.items() if k.endswith(("running_mean", "running_var"))}
    return L.detach(), {k: g for k, g in zip(keys, grads) if g is not None}, outs
