"""Torch-tensor wrappers over the per-op C-ABI entry points (unit tests and microbenchmarks).

"planes": a split-precision fp16 pair stored as one tensor of shape (2, ...) — hi = rn(x), lo = rn(x - hi).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr

OUT_F32, OUT_F16, OUT_F16_SPLIT = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU, ACT_TANH = 0, 1, 2, 3


def _planes_like(shape, device):
    return torch.empty((2,) + tuple(shape), dtype=torch.float16, device=device)


def plane_stride(p):
    return p[0].numel()


def split(x):
    """fp32 tensor -> planes (2, *x.shape) via the CUDA kernel."""
    x = x.contiguous().float()
    out = _planes_like(x.shape, x.device)
    call("maed_op_split_f32", ptr(x), ptr(out), plane_stride(out), x.numel(), stream_ptr())
    return out


def join(p):
    return p[0].float() + p[1].float()


def gemm(a, b, nsplit=3, bias=None, residual=None, act=ACT_NONE, out_mode=OUT_F32, block_n=0):
    """a: planes (2,M,K), b: planes (2,N,K)  ->  fp32 (M,N) or planes (2,M,N)."""
    M, K = a.shape[1:]
    N = b.shape[1]
    dev = a.device
    if out_mode == OUT_F32:
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        oplane = 0
    else:
        out = _planes_like((M, N), dev)
        if out_mode == OUT_F16:
            out[1].zero_()
        oplane = M * N
    call("maed_op_gemm", ptr(a), plane_stride(a), K, ptr(b), plane_stride(b), K, M, N, K, nsplit, ptr(bias),
         ptr(residual), act, out_mode, ptr(out), oplane, N, block_n, stream_ptr())
    return out


def conv_gemm(a, w, KH, KW, pad_h, pad_w, nsplit=3, block_n=0):
    """a: planes (2, n, H, W, Cin) NHWC; w: planes (2, Cout, KH*KW*Cin) -> fp32 (n*H*W, Cout)."""
    _, n, H, W, Cin = a.shape
    Cout = w.shape[1]
    out = torch.empty(n * H * W, Cout, dtype=torch.float32, device=a.device)
    call("maed_op_conv_gemm", ptr(a), plane_stride(a), ptr(w), plane_stride(w), n, H, W, Cin, Cout, KH, KW, pad_h, pad_w,
         nsplit, OUT_F32, ptr(out), 0, block_n, stream_ptr())
    return out


def prep_conv_weight(w, k_pad=None, standardize=True):
    Cout, Cin, KH, KW = w.shape
    k_pad = k_pad or Cin * KH * KW
    out = _planes_like((Cout, k_pad), w.device)
    call("maed_op_prep_conv_weight", ptr(w.contiguous()), Cout, Cin, KH, KW, k_pad, int(standardize), ptr(out),
         plane_stride(out), stream_ptr())
    return out


def im2col_stem(x, k_pad=152):
    n = x.shape[0]
    out = _planes_like((n * 112 * 112, k_pad), x.device)
    call("maed_op_im2col_stem", ptr(x.contiguous()), n, 3, 224, 224, 7, 7, 2, 2, 2, 112, 112, k_pad, ptr(out),
         plane_stride(out), stream_ptr())
    return out


def conv_gn(a, w, KH, KW, gamma, beta, relu, residual=None, nsplit=3, eps=1e-5, dbg=None):
    """Fused conv + GroupNorm(32) (+ residual) (+ ReLU).  a: planes (2, n, H, W, Cin); w: planes (2, C, KH*KW*Cin);
    residual: planes (2, n, H, W, C) -> planes (2, n, H, W, C)."""
    _, n, H, W, Cin = a.shape
    Cc = w.shape[1]
    out = _planes_like((n, H, W, Cc), a.device)
    call("maed_op_conv_gn", ptr(a), plane_stride(a), ptr(w), plane_stride(w), n, H, W, Cin, Cc, KH, KW, nsplit, ptr(gamma),
         ptr(beta), C.c_float(eps), int(relu), ptr(residual), plane_stride(residual) if residual is not None else 0, ptr(out),
         plane_stride(out), ptr(dbg), stream_ptr())
    return out


def stem_conv(x, w_planes, nsplit=3):
    """x: fp32 (n,3,224,224); w_planes: prep_conv_weight(w, k_pad=152) -> (fp32 (n*112*112, 64), stats (n,32,2) f64)."""
    n = x.shape[0]
    out = torch.empty(n * 112 * 112, 64, dtype=torch.float32, device=x.device)
    stats = torch.empty(n, 32, 2, dtype=torch.float64, device=x.device)
    call("maed_op_stem_conv", ptr(x.contiguous()), n, ptr(w_planes), plane_stride(w_planes), w_planes.shape[2], nsplit,
         ptr(out), ptr(stats), stream_ptr())
    return out, stats


def im2col_nhwc(a, KH, KW, stride, pad_t, pad_l, OH, OW):
    _, n, H, W, Cc = a.shape
    out = _planes_like((n * OH * OW, KH * KW * Cc), a.device)
    call("maed_op_im2col_nhwc", ptr(a), plane_stride(a), n, H, W, Cc, KH, KW, stride, pad_t, pad_l, OH, OW, ptr(out),
         plane_stride(out), stream_ptr())
    return out


def groupnorm(x, gamma, beta, relu, residual=None, eps=1e-5):
    """x: fp32 (n, HW, C) NHWC-flattened; residual: planes (2, n, HW, C) -> planes (2, n, HW, C)."""
    n, HW, Cc = x.shape
    out = _planes_like(x.shape, x.device)
    scratch = torch.empty(n * 64, dtype=torch.float64, device=x.device)
    call("maed_op_groupnorm", ptr(x), n, HW, Cc, ptr(gamma), ptr(beta), C.c_float(eps), int(relu), ptr(residual),
         plane_stride(residual) if residual is not None else 0, ptr(out), plane_stride(out), ptr(scratch), stream_ptr())
    return out


def groupnorm_maxpool(x, gamma, beta, eps=1e-5):
    """x: fp32 (n, H, W, C) -> planes (2, n, ceil(H/2), ceil(W/2), C)."""
    n, H, W, Cc = x.shape
    out = _planes_like((n, (H + 1) // 2, (W + 1) // 2, Cc), x.device)
    scratch = torch.empty(n * 64, dtype=torch.float64, device=x.device)
    call("maed_op_groupnorm_maxpool", ptr(x), n, H, W, Cc, ptr(gamma), ptr(beta), C.c_float(eps), ptr(out),
         plane_stride(out), ptr(scratch), stream_ptr())
    return out


def layernorm(x, gamma, beta, eps=1e-6):
    rows, Cc = x.shape
    out = _planes_like(x.shape, x.device)
    call("maed_op_layernorm", ptr(x), Cc, ptr(gamma), ptr(beta), rows, Cc, C.c_float(eps), ptr(out), plane_stride(out),
         stream_ptr())
    return out


def attention(kind, qkv, B, T, ntok, heads, scale, nsplit=3, planes=False):
    """qkv: planes (2, B*T*ntok, 3*heads*64) -> fp32 (B*T*ntok, heads*64), or with `planes=True` the fp16 hi/lo
    planes (2, rows, heads*64) the engine consumes.  kind: 'spatial'|'temporal'|'generic'."""
    rows = qkv.shape[1]
    k = {"spatial": 0, "temporal": 1, "generic": 2}[kind]
    if planes:
        out = torch.empty(2, rows, heads * 64, dtype=torch.float16, device=qkv.device)
        call("maed_op_attention", k, ptr(qkv), plane_stride(qkv), B, T, ntok, heads, C.c_float(scale), nsplit, None, ptr(out),
             plane_stride(out), stream_ptr())
        return out
    out = torch.empty(rows, heads * 64, dtype=torch.float32, device=qkv.device)
    call("maed_op_attention", k, ptr(qkv), plane_stride(qkv), B, T, ntok, heads, C.c_float(scale), nsplit, ptr(out), None, 0,
         stream_ptr())
    return out


def linear_f32(x, W, bias=None, act=ACT_NONE, residual=None):
    R, K = x.shape
    N = W.shape[0]
    out = torch.empty(R, N, dtype=torch.float32, device=x.device)
    call("maed_op_linear_f32", ptr(x), K, ptr(W), K, ptr(bias), R, N, K, act, ptr(residual), N, ptr(out), N, stream_ptr())
    return out


def decode_outputs(pose6d, shape, cam, n_joints=49):
    R = pose6d.shape[0]
    dev = pose6d.device
    rot = torch.empty(R, 24, 3, 3, dtype=torch.float32, device=dev)
    theta = torch.empty(R, 85, dtype=torch.float32, device=dev)
    kp2d = torch.empty(R, n_joints, 2, dtype=torch.float32, device=dev)
    call("maed_op_decode_outputs", ptr(pose6d), ptr(shape), ptr(cam), R, None, n_joints, ptr(rot), ptr(theta), ptr(kp2d),
         stream_ptr())
    return rot, theta, kp2d


def launch_count():
    return _lib.load().maed_launch_count()
