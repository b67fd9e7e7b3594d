"""`MAED` — drop-in for the reference's `lib.models.MAED` (reference lib/models/maed.py:9-66).

Same constructor, same `forward(x, J_regressor=None) -> dict(theta, verts, kp_2d, kp_3d, rotmat)`, same
`extract_feature`, same `state_dict()` keys; the arithmetic runs in `libmaed_b200.so` (hand-written sm_100a
CUDA: tcgen05/TMA GEMMs and attention, fused norm kernels) through one C-ABI call per forward.

Differences a caller can observe (documented in INTEGRATION.md):
  * inputs must be CUDA float32 tensors — there is no CPU / eager fallback, by design;
  * in `train()` mode with autograd enabled `forward` runs the CUDA training path (maed_b200/train.py: activation tape,
    backward kernels to every parameter) exactly like the reference module under `loss.backward()`; `eval()` or
    `torch.no_grad()` runs the fused inference engine;
  * `encoder='cnn'` (torchvision ResNet-50, stage-1 config): eval() folds BatchNorm into the conv weights; train() uses batch
    statistics and, under `torch.distributed` with more than one rank, exchanges them like `nn.SyncBatchNorm`
    (reference train.py:95; `model.sync_batchnorm = False` keeps per-rank statistics);
  * `decoder.smpl.*` buffers do not exist (smplx and the SMPL assets are absent): `verts`/`kp_3d` are zeros unless a body
    model is installed with `load_smpl_assets()`.
"""
import ctypes as C
import os

import torch
import torch.nn as nn

from .. import _lib
from .modules import KTD, CNNEncoder, Regressor, STEncoder


def _default_precision():
    p = os.environ.get("MAED_B200_PRECISION", "split")
    if p not in ("split", "fp16"):
        raise ValueError("MAED_B200_PRECISION must be 'split' or 'fp16', got %r" % p)
    return p


class MAED(nn.Module):
    def __init__(self, encoder="ste", num_blocks=6, num_heads=12, st_mode="parallel", decoder="ktd", hidden_dim=1024,
                 precision=None, temp_frames=16, **kwargs):
        """`precision`: 'split' (default; every tensor-core operand is an fp16 hi/lo pair, 3 MMAs per K step —
        meets the reference's 1e-3 parity gate) or 'fp16' (single MMA; ~3x less tensor work, ~1e-2 error
        on random-weight models).  `temp_frames`: rows of `temp_embed` (reference: 16)."""
        super().__init__()
        self.encoder_type = encoder
        self.decoder_type = decoder
        if encoder.lower() == "cnn":
            # torchvision ResNet-50, fc = Identity (reference maed.py:35-37; the stage-1 config).  Inference only:
            # BatchNorm runs on its running statistics; num_blocks / num_heads / st_mode are ignored like in the reference
            self.encoder = CNNEncoder()
            if st_mode not in _lib.MODES:
                st_mode = "vanilla"
        elif encoder.lower() == "ste":
            self.encoder = STEncoder(num_blocks, num_heads, st_mode, temp_frames=temp_frames)
        else:
            raise NotImplementedError(encoder)
        # what determine_output_feature_dim() measures (utils.py:185-198): 768 for 'ste', 2048 for 'cnn'
        feat_dim = self.feat_dim = self.encoder.num_features
        if decoder.lower() == "ktd":
            self.decoder = KTD(feat_dim=feat_dim, hidden_dim=hidden_dim)
        elif decoder.lower() == "iterative":
            self.decoder = Regressor(feat_dim=feat_dim, hidden_dim=hidden_dim, mean_params=kwargs.get("mean_params"))
        else:
            raise NotImplementedError(decoder)
        if kwargs.get("smpl_assets") is not None:
            self.decoder.smpl.load_assets(kwargs["smpl_assets"])
        self.precision = precision or _default_precision()
        self._cfg = _lib.MaedConfig(num_blocks, num_heads, _lib.MODES[st_mode], _lib.DECODERS[decoder.lower()],
                                    hidden_dim, 3 if self.precision == "split" else 1, temp_frames,
                                    _lib.ENCODERS[encoder.lower()])
        self._engine = None
        self._packed = None
        self._packed_key = None
        self._pack_gen = 0
        self._workspace = None
        self._param_ptrs = None
        self._training_enabled = True
        self._train_dropout_p = None
        self._train_state = None

    # ------------------------------------------------------------------------------------------ engine
    def _get_engine(self):
        if self._engine is None:
            h = C.c_void_p()
            _lib.call("maed_engine_create", C.byref(self._cfg), C.byref(h))
            self._engine = h
            lib = _lib.load()
            n = lib.maed_engine_num_params(h)
            self._param_names = [lib.maed_engine_param_name(h, i).decode() for i in range(n)]
            self._param_numels = [lib.maed_engine_param_numel(h, i) for i in range(n)]
        return self._engine

    def __del__(self):
        try:
            if getattr(self, "_engine", None) is not None:
                _lib.load().maed_engine_destroy(self._engine)
        except Exception:
            pass

    def _tensor_table(self):
        sd = dict(self.named_parameters())
        sd.update(dict(self.named_buffers()))
        out = []
        for name, numel in zip(self._param_names, self._param_numels):
            t = sd[name]
            if t.numel() != numel:
                raise RuntimeError("parameter %s has %d elements, engine expects %d" % (name, t.numel(), numel))
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("parameter %s must be a contiguous CUDA float32 tensor (got %s on %s); move the "
                                   "model with .to('cuda')" % (name, t.dtype, t.device))
            out.append(t)
        return out

    def _prepare(self, device):
        """(Re)derives the packed tensor-core weights when any parameter changed (version counters)."""
        eng = self._get_engine()
        tensors = self._tensor_table()
        key = (str(device),) + tuple((t.data_ptr(), t._version) for t in tensors)
        arr = (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])
        self._param_ptrs = arr
        if key != self._packed_key:
            lib = _lib.load()
            nbytes = lib.maed_engine_packed_bytes(eng)
            if self._packed is None or self._packed.numel() < nbytes or self._packed.device != device:
                self._packed = torch.empty(nbytes, dtype=torch.uint8, device=device)
            _lib.call("maed_engine_pack", eng, arr, _lib.ptr(self._packed), _lib.stream_ptr())
            self._packed_key = key
            self._pack_gen += 1          # derived caches elsewhere (train.TrainState.tpack) follow this counter
        return eng

    def invalidate_cache(self):
        """Force re-packing of the derived weights at the next forward (e.g. after in-place `.data` edits that do
        not bump tensor version counters)."""
        self._packed_key = None

    def _run(self, x, want_taps=None):
        if x.dim() != 5 or x.shape[2:] != (3, 224, 224):
            raise ValueError("MAED expects (N, T, 3, 224, 224) frames, got %s" % (tuple(x.shape),))
        if not x.is_cuda:
            raise RuntimeError("maed_b200.MAED runs on CUDA (sm_100a) only; got a %s tensor — there is no CPU fallback"
                               % x.device)
        N, T = x.shape[:2]
        x = x.to(torch.float32).contiguous()
        dev = x.device
        with torch.cuda.device(dev):
            eng = self._prepare(dev)
            lib = _lib.load()
            BT = N * T
            wbytes = lib.maed_engine_workspace_bytes(eng, BT)
            if self._workspace is None or self._workspace.numel() < wbytes or self._workspace.device != dev:
                self._workspace = torch.empty(wbytes, dtype=torch.uint8, device=dev)
            f32 = dict(dtype=torch.float32, device=dev)
            nj = self.decoder.smpl.n_joints
            o = {"feat": torch.empty(BT, self.feat_dim, **f32), "pose6d": torch.empty(BT, 144, **f32),
                 "shape": torch.empty(BT, 10, **f32), "cam": torch.empty(BT, 3, **f32),
                 "rotmat": torch.empty(BT, 24, 3, 3, **f32), "theta": torch.empty(BT, 85, **f32),
                 "kp_2d": torch.empty(BT, nj, 2, **f32)}
            outs = _lib.MaedOutputs(_lib.ptr(o["feat"]), _lib.ptr(o["pose6d"]), _lib.ptr(o["shape"]), _lib.ptr(o["cam"]),
                                    _lib.ptr(o["rotmat"]), _lib.ptr(o["theta"]), _lib.ptr(o["kp_2d"]), None, nj)
            taps_arr = None
            taps = {}
            if want_taps:
                shapes = {"stem": (BT, 56, 56, 64), "stage0": (BT, 56, 56, 256), "stage1": (BT, 28, 28, 512),
                          "stage2": (BT, 14, 14, 1024), "embed": (BT, 197, 768)}
                if self.encoder_type.lower() == "cnn":          # layer4's output arrives in the 'embed' slot
                    shapes["embed"] = (BT, 7, 7, 2048)
                for i in range(8):
                    shapes["block%d" % i] = (BT, 197, 768)
                ptrs = []
                for name in _lib.TAP_NAMES:
                    if name in want_taps and (not name.startswith("block") or int(name[5:]) < self._cfg.num_blocks):
                        taps[name] = torch.empty(shapes[name], **f32)
                        ptrs.append(taps[name].data_ptr())
                    else:
                        ptrs.append(None)
                taps_arr = (C.c_void_p * len(ptrs))(*ptrs)
            _lib.call("maed_engine_forward", eng, self._param_ptrs, _lib.ptr(self._packed), _lib.ptr(x), N, T,
                      _lib.ptr(self._workspace), C.c_size_t(self._workspace.numel()), C.byref(outs), taps_arr,
                      _lib.stream_ptr())
        o["taps"] = taps
        return o

    # ------------------------------------------------------------------------------------ public API
    @torch.no_grad()
    def extract_feature(self, x):
        """reference maed.py:43-50: (N,T,3,H,W) -> (N,T,768) ('cnn': 2048)."""
        N, T = x.shape[:2]
        return self._run(x)["feat"].reshape(N, T, -1)

    def enable_training(self, flag=True, dropout_p=None):
        """train()-mode forwards with autograd enabled run the engine's training path (maed_b200/train.py: saved-activation
        tape + CUDA backward) — the default.  `enable_training(False)` turns train()-mode forwards into graph-less inference
        calls (with a warning); `dropout_p` overrides the decoders' nn.Dropout() probability (0.0 for parity runs)."""
        self._training_enabled = bool(flag)
        self._train_dropout_p = dropout_p       # None: nn.Dropout() default 0.5 (ktd.py:54-56); 0.0 for parity runs
        return self

    def forward(self, x, J_regressor=None, **kwargs):
        """reference maed.py:52-66.  `J_regressor` (17x6890) only matters once the SMPL tier exists: with the
        placeholder body model verts are zeros, so J_regressor @ verts is zeros as well."""
        wants_grad = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if wants_grad and kwargs.get("_allow_train_mode") is None:
            if getattr(self, "_training_enabled", True):
                from .. import train as _train
                return _train.train_forward(self, x, J_regressor)
            import warnings
            warnings.warn("maed_b200.MAED: the training path was switched off with enable_training(False): outputs carry no "
                          "autograd graph and train()-mode dropout (reference ktd.py:54-56) is not applied.",
                          RuntimeWarning, stacklevel=2)
        with torch.no_grad():
            return self._forward_inference(x, J_regressor, **kwargs)

    def load_smpl_assets(self, assets):
        """Install a body model (see SMPLHead.load_assets); afterwards verts / kp_3d / kp_2d are computed by csrc/smpl.cu."""
        self.decoder.smpl.load_assets(assets)
        dev = next(self.parameters()).device
        self.decoder.smpl.to(dev)
        return self

    def _smpl(self, o, J_regressor):
        """verts, joints and their projection from the engine's shape / rotmat (reference ktd.py:100-114)."""
        smpl = self.decoder.smpl
        BT = o["shape"].shape[0]
        dev = o["shape"].device
        f32 = dict(dtype=torch.float32, device=dev)
        if smpl.v_template.device != dev:
            smpl.to(dev)
        assets = _lib.MaedSmplAssets(*[_lib.ptr(getattr(smpl, k)) for k in (
            "v_template", "shapedirs", "posedirs", "J_template", "J_shapedirs", "lbs_weights", "J_regressor_extra", "parents",
            "extra_vertex_ids", "joint_map")])
        reg, n_reg = None, 0
        if J_regressor is not None:
            reg = J_regressor.to(dev, torch.float32).contiguous()
            n_reg = reg.shape[0]
        nj = n_reg if reg is not None else smpl.n_joints
        verts = torch.empty(BT, 6890, 3, **f32)
        joints = torch.empty(BT, nj, 3, **f32)
        kp2d = torch.empty(BT, nj, 2, **f32)
        with torch.cuda.device(dev):
            lib = _lib.load()
            nbytes = lib.maed_smpl_scratch_bytes(BT)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("maed_smpl_forward", C.byref(assets), _lib.ptr(o["shape"]), _lib.ptr(o["rotmat"]), BT, _lib.ptr(reg), n_reg,
                      _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(scratch), C.c_size_t(nbytes), _lib.stream_ptr())
            rot, theta = torch.empty_like(o["rotmat"]), torch.empty_like(o["theta"])
            _lib.call("maed_op_decode_outputs", _lib.ptr(o["pose6d"]), _lib.ptr(o["shape"]), _lib.ptr(o["cam"]), BT,
                      _lib.ptr(joints), nj, _lib.ptr(rot), _lib.ptr(theta), _lib.ptr(kp2d), _lib.stream_ptr())
        return verts, joints, kp2d

    def _forward_inference(self, x, J_regressor=None, **kwargs):
        N, T = x.shape[:2]
        o = self._run(x, want_taps=kwargs.get("_taps"))
        nj = 17 if J_regressor is not None else self.decoder.smpl.n_joints
        if self.decoder.smpl.has_assets:
            verts, kp3d, kp2d = self._smpl(o, J_regressor)
            nj = kp3d.shape[1]
        else:
            kp2d = o["kp_2d"] if nj == o["kp_2d"].shape[1] else o["kp_2d"][:, :nj].contiguous()
            verts = torch.zeros(N * T, 6890, 3, dtype=torch.float32, device=x.device)
            kp3d = torch.zeros(N * T, nj, 3, dtype=torch.float32, device=x.device)
        out = {
            "theta": o["theta"].reshape(N, T, -1),
            "verts": verts.reshape(N, T, 6890, 3),
            "kp_2d": kp2d.reshape(N, T, -1, 2),
            "kp_3d": kp3d.reshape(N, T, nj, 3),
            "rotmat": o["rotmat"].reshape(N, T, -1, 3, 3),
        }
        if kwargs.get("_taps") or kwargs.get("_debug"):
            out["_debug"] = {k: o[k] for k in ("feat", "pose6d", "shape", "cam")}
            out["_debug"].update(o["taps"])
        return out

    @torch.no_grad()
    def forward_subclips(self, images, seqlen=16, interp=1, J_regressor=None):
        """The evaluator's inner loop (reference lib/core/evaluate.py:71-100) as ONE forward.

        The reference feeds a long window `images` (N, L0, 3, 224, 224) to the model as `sample_freq = L // seqlen`
        interleaved sub-clips, `images[:, ::interp][:, i::sample_freq]` for i in range(sample_freq) (L = frames left after
        the `interp` stride), copies five outputs to the host after each call and re-interleaves them with
        `merge_sequence` (evaluate.py:127-133).  Here the sub-clips are gathered on the device into one batch of
        N * sample_freq clips (the path shards by clip, so the results are the same), run through one engine call, and
        returned already merged: every value has shape (N, L, ...) in the frame order of `images[:, ::interp]` — what
        `merge_sequence` yields before `interpolate`.  Requires L to be a multiple of seqlen, as the reference's stacking
        does."""
        if images.dim() != 5:
            raise ValueError("forward_subclips expects (N, L, 3, 224, 224) frames, got %s" % (tuple(images.shape),))
        x = images[:, ::interp]
        N, L = x.shape[:2]
        if seqlen < 1 or L < seqlen or L % seqlen != 0:
            raise ValueError("window of %d frames (after interp=%d) is not a multiple of seqlen=%d" % (L, interp, seqlen))
        sf = L // seqlen
        # frame t of sub-clip i is window frame t * sf + i
        clips = x.reshape(N, seqlen, sf, *x.shape[2:]).transpose(1, 2).reshape(N * sf, seqlen, *x.shape[2:])
        out = self._forward_inference(clips.contiguous(), J_regressor)
        return {k: v.reshape(N, sf, seqlen, *v.shape[2:]).transpose(1, 2).reshape(N, L, *v.shape[2:]) for k, v in out.items()}
