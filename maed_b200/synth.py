"""Deterministic synthetic parameters and inputs (bench.py, the scripts, the tests and the oracle: data only, no algorithm).

The reference ships no checkpoints we can reach (pretrained ViT weights are a download,
`lib/models/vision_transformer.py:36`; SMPL mean params are licensed data, `lib/models/spin.py:42`), and
a 72 M-parameter state_dict (288 MB) cannot be committed as a fixture.  Instead every parameter tensor is
a pure function of (its state_dict key, its shape, a seed): the golden-vector generator loads these
values into the REFERENCE model, the tests load the same values into the B200 model and into the
oracle, so all three compute on identical weights without any file travelling.

Values are deliberately *non-trivial* (norm weights != 1, all biases != 0, peaky attention logits) so
that a kernel which drops a bias or an affine term fails parity; the reference's own init
(`vision_transformer.py:366-375`: LN=(1,0), Linear bias 0) would hide such bugs.
"""
import zlib

import numpy as np
import torch


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def synth_tensor(key: str, shape, seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)
    shape = tuple(shape)
    r = torch.randn(shape, generator=g, dtype=torch.float32)
    leaf = key.split(".")[-1]
    parent = key.split(".")[-2] if "." in key else ""
    if key.endswith(("cls_token", "pos_embed", "temp_embed")):
        return 0.1 * r
    if leaf == "init_pose":       # 6-D identity + noise (stand-in for smpl_mean_params['pose'])
        base = torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24).reshape(shape)
        return base + 0.2 * r
    if leaf == "init_shape":
        return 0.3 * r
    if leaf == "init_cam":
        return torch.tensor([0.9, 0.0, 0.0]).reshape(shape) + 0.05 * r
    is_norm = parent.startswith("norm") or parent == "norm"
    # torchvision ResNet-50 ('cnn' encoder): bn1/bn2/bn3 and downsample.1 are BatchNorm2d with running statistics
    is_bn = parent.startswith("bn") or (parent == "1" and ".downsample." in key)
    if is_bn and len(shape) == 1:
        if leaf == "weight":                              # 0.6 inside the residual stages keeps the activations O(1)
            return (1.0 if key.startswith("encoder.bn1.") else 0.6) * (1.0 + 0.1 * r)
        if leaf == "running_var":
            return 0.6 + 0.4 * r.abs()
        return 0.1 * r                                    # bias, running_mean
    if len(shape) == 1:
        if is_norm and leaf == "weight":
            return 1.0 + 0.1 * r
        if is_norm and leaf == "bias":
            return 0.1 * r
        return 0.02 * r                                   # Linear / conv bias
    if len(shape) == 4:                                   # conv weight: kaiming-normal(fan_out)-like
        fan_out = shape[0] * shape[2] * shape[3]
        return r * float(np.sqrt(2.0 / fan_out))
    if len(shape) == 2:
        if parent == "qkv":
            return 0.05 * r                               # peaky-ish attention logits
        if parent == "ts_attn":
            return 0.05 * r
        return 0.02 * r
    return 0.02 * r


def fill_module_(module: torch.nn.Module, seed: int = 0, skip=("smpl",)):
    """In-place: overwrite every parameter and float buffer of `module` (reference or B200 model)."""
    with torch.no_grad():
        for k, v in list(module.named_parameters()) + list(module.named_buffers()):
            if any(s in k for s in skip) or not v.dtype.is_floating_point:
                continue
            v.copy_(synth_tensor(k, v.shape, seed).to(v.dtype))
    return module


def synth_state_dict(shapes: dict, seed: int = 0) -> dict:
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_frames(n: int, t: int, seed: int = 0, size: int = 224) -> torch.Tensor:
    """ImageNet-normalised frames are ~N(0,1) (`lib/data_utils/transforms/basic.py:6-7`)."""
    g = _gen("frames_%d_%d_%d" % (n, t, size), seed)
    return torch.randn((n, t, 3, size, size), generator=g, dtype=torch.float32)


def mean_params():
    """Synthetic stand-in for data/smpl_data/smpl_mean_params.npz (pose (144,), shape (10,), cam (3,))."""
    return {
        "pose": synth_tensor("decoder.init_pose", (144,), 0).numpy().astype(np.float32),
        "shape": synth_tensor("decoder.init_shape", (10,), 0).numpy().astype(np.float32),
        "cam": synth_tensor("decoder.init_cam", (3,), 0).numpy().astype(np.float32),
    }


def tap_digest(t: torch.Tensor, max_samples: int = 4096):
    """Compact fingerprint of an intermediate tensor for the golden files: a strided subsample of the
    flattened tensor plus (mean, std, absmax).  Used identically by make_golden.py and the tests."""
    f = t.detach().float().reshape(-1).cpu()
    stride = max(1, f.numel() // max_samples)
    sub = f[::stride][:max_samples].clone()
    stats = torch.stack([f.mean(), f.std(), f.abs().max()])
    return sub.numpy(), stats.numpy()
