"""Python side of the CUDA-on-CPU test build (tests/emu/README.md) — TEST INFRASTRUCTURE ONLY.

`activate()` points `maed_b200._lib` at tests/emu/_build/libmaed_emu.so for the duration of a test, so that the per-op
wrappers of `maed_b200.ops` and raw `_lib.call(...)` invocations run the real CUDA-core kernel sources (compiled by g++
against cuda_emu.h) on CPU tensors.  `EmuModel` drives the engine entry points (pack / forward / train_forward /
train_backward) directly through the C ABI with CPU tensors: the product module `maed_b200.models.MAED` refuses non-CUDA
inputs by design, and nothing under maed_b200/ knows about the emulator.
"""
import contextlib
import ctypes as C
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from maed_b200 import _lib, ops  # noqa: E402

_emu = None


def load():
    global _emu
    if _emu is None:
        sys.path.insert(0, HERE)
        import build_emu
        lib = C.CDLL(build_emu.build())
        for name, (res, args) in _lib.SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _emu = lib
    return _emu


@contextlib.contextmanager
def activate():
    """maed_b200._lib / maed_b200.ops talk to the emulator library inside the block (stream pointer = NULL)."""
    lib = load()
    saved = (_lib._lib, _lib.stream_ptr, ops.stream_ptr)
    _lib._lib, _lib.stream_ptr, ops.stream_ptr = lib, (lambda: None), (lambda: None)
    try:
        yield lib
    finally:
        _lib._lib, _lib.stream_ptr, ops.stream_ptr = saved


@contextlib.contextmanager
def product_on_cpu():
    """Runs the PRODUCT Python modules (maed_b200.models.MAED, maed_b200.train) on CPU tensors against the emulator library,
    so that the autograd boundary, the flat gradient buffer, FusedAdam and the geometry tail are exercised without a GPU.
    Test-only trickery: `torch.Tensor.is_cuda` is shadowed by a property that answers True and `torch.cuda.device` by a
    no-op context manager for the duration of the block — the product code keeps refusing CPU tensors everywhere else."""
    class _NoDevice:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    with activate() as lib:
        saved = torch.cuda.device
        torch.cuda.device = _NoDevice
        torch.Tensor.is_cuda = property(lambda self: True)
        try:
            yield lib
        finally:
            del torch.Tensor.is_cuda
            torch.cuda.device = saved


class EmuModel:
    """Engine driver on CPU tensors: the same C-ABI call sequence as maed_b200.models.MAED._run and
    maed_b200.train.MaedTrainFunction, minus torch.cuda."""

    def __init__(self, module):
        self.m = module
        self.lib = load()
        h = C.c_void_p()
        self._check(self.lib.maed_engine_create(C.byref(module._cfg), C.byref(h)), "engine_create")
        self.eng = h
        n = self.lib.maed_engine_num_params(h)
        self.names = [self.lib.maed_engine_param_name(h, i).decode() for i in range(n)]
        sd = dict(module.named_parameters())
        sd.update(dict(module.named_buffers()))
        self.tensors = [sd[k].detach().float().contiguous() for k in self.names]
        for t, k, i in zip(self.tensors, self.names, range(n)):
            assert t.numel() == self.lib.maed_engine_param_numel(h, i), k
        self.params = (C.c_void_p * n)(*[t.data_ptr() for t in self.tensors])
        self.packed = torch.full((self.lib.maed_engine_packed_bytes(h),), 0xFF, dtype=torch.uint8)            # poisoned
        self._check(self.lib.maed_engine_pack(h, self.params, _lib.ptr(self.packed), None), "engine_pack")
        self.tpack = None
        self.ws = None

    def __del__(self):
        try:
            self.lib.maed_engine_destroy(self.eng)
        except Exception:
            pass

    def _check(self, status, what):
        if status != 0:
            msg = self.lib.maed_last_error()
            raise RuntimeError("emulated %s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))

    def forward(self, x, taps=()):
        N, T = x.shape[:2]
        BT = N * T
        x = x.float().contiguous()
        # poisoned like uninitialised device memory (torch.empty on the GPU): 0xFF bytes are NaN in fp32 and in fp16
        ws = torch.full((self.lib.maed_engine_workspace_bytes(self.eng, BT),), 0xFF, dtype=torch.uint8)
        nj = 49
        o = {"feat": torch.zeros(BT, getattr(self.m, "feat_dim", 768)), "pose6d": torch.zeros(BT, 144), "shape": torch.zeros(BT, 10),
             "cam": torch.zeros(BT, 3), "rotmat": torch.zeros(BT, 24, 3, 3), "theta": torch.zeros(BT, 85),
             "kp_2d": torch.zeros(BT, nj, 2)}
        outs = _lib.MaedOutputs(_lib.ptr(o["feat"]), _lib.ptr(o["pose6d"]), _lib.ptr(o["shape"]), _lib.ptr(o["cam"]),
                                _lib.ptr(o["rotmat"]), _lib.ptr(o["theta"]), _lib.ptr(o["kp_2d"]), None, nj)
        shapes = {"stem": (BT, 56, 56, 64), "stage0": (BT, 56, 56, 256), "stage1": (BT, 28, 28, 512),
                  "stage2": (BT, 14, 14, 1024), "embed": (BT, 197, 768)}
        for i in range(8):
            shapes["block%d" % i] = (BT, 197, 768)
        if getattr(self.m, "encoder_type", "ste").lower() == "cnn":
            shapes["embed"] = (BT, 7, 7, 2048)                     # layer4's output arrives in the 'embed' slot
        tap_t, ptrs = {}, []
        for name in _lib.TAP_NAMES:
            if name in taps:
                tap_t[name] = torch.zeros(shapes[name])
                ptrs.append(tap_t[name].data_ptr())
            else:
                ptrs.append(None)
        taps_arr = (C.c_void_p * len(ptrs))(*ptrs) if taps else None
        self._check(self.lib.maed_engine_forward(self.eng, self.params, _lib.ptr(self.packed), _lib.ptr(x), N, T, _lib.ptr(ws),
                                                 C.c_size_t(ws.numel()), C.byref(outs), taps_arr, None), "engine_forward")
        o["taps"] = tap_t
        return o

    def train_forward(self, x, dropout_p=0.0, seed=1):
        N, T = x.shape[:2]
        BT = N * T
        self.x = x.float().contiguous()
        if self.tpack is None:
            self.tpack = torch.full((self.lib.maed_train_pack_bytes(self.eng),), 0xFF, dtype=torch.uint8)         # poisoned
            self._check(self.lib.maed_train_pack(self.eng, self.params, _lib.ptr(self.tpack), None), "train_pack")
        self.ws = torch.full((self.lib.maed_train_workspace_bytes(self.eng, BT),), 0xFF, dtype=torch.uint8)   # poisoned, see forward
        o = {"feat": torch.zeros(BT, getattr(self.m, "feat_dim", 768)), "pose6d": torch.zeros(BT, 144), "shape": torch.zeros(BT, 10),
             "cam": torch.zeros(BT, 3)}
        outs = _lib.MaedTrainOutputs(_lib.ptr(o["feat"]), _lib.ptr(o["pose6d"]), _lib.ptr(o["shape"]), _lib.ptr(o["cam"]))
        self._check(self.lib.maed_train_forward(self.eng, self.params, _lib.ptr(self.packed), _lib.ptr(self.x), N, T,
                                                _lib.ptr(self.ws), C.c_size_t(self.ws.numel()), C.c_float(dropout_p),
                                                C.c_ulonglong(seed), C.byref(outs), None), "train_forward")
        self._nt = (N, T)
        return o

    def train_backward(self, d_pose, d_shape, d_cam, loss_scale=4096.0, dropout_p=0.0):
        N, T = self._nt
        grads = [torch.full_like(t, float("nan")) for t in self.tensors]          # every entry must be written
        is_buf = [k.endswith(("running_mean", "running_var")) for k in self.names]  # buffers have no gradient: NULL slot
        gp = (C.c_void_p * len(grads))(*[None if b else g.data_ptr() for g, b in zip(grads, is_buf)])
        d_pose, d_shape, d_cam = [t.float().contiguous() for t in (d_pose, d_shape, d_cam)]
        self._check(self.lib.maed_train_backward(self.eng, self.params, _lib.ptr(self.packed), _lib.ptr(self.tpack),
                                                 _lib.ptr(self.x), N, T, _lib.ptr(self.ws), C.c_size_t(self.ws.numel()),
                                                 _lib.ptr(d_pose), _lib.ptr(d_shape), _lib.ptr(d_cam), C.c_float(loss_scale),
                                                 C.c_float(dropout_p), gp, None), "train_backward")
        return dict(zip(self.names, grads))
