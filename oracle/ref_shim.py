"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (`/root/reference/lib/models`) on CPU.

The tree is ``/root/reference`` in the build container and its verbatim, git-ignored copy ``oracle/_ref/``
(``oracle/vendor_ref.py``, made by ``__graft_entry__.build()``) on the GPU box.  Used by
``tests/golden/make_golden*.py`` to generate the committed golden vectors and by ``bench.py --impl reference``
/ the ``cpu_baseline`` leg to time the reference's own code on the host cores.

The reference cannot be imported as-is on torch 2.11 / offline:
  * ``lib/models/vision_transformer.py:19,23`` import modules removed from torch / torchvision;
  * ``lib/models/smpl.py:9-11`` needs ``smplx`` (un-vendored third party, not installable);
  * ``lib/core/config.py:3`` needs ``yacs``;
  * ``lib/models/maed.py:36,39`` download pretrained weights;
  * ``lib/models/smpl.py:90`` / ``lib/models/spin.py:42`` read licensed data files relative to CWD.
Every blocker is shimmed through ``sys.modules`` / a scratch CWD; nothing under /root/reference is
modified.  The ``smplx.SMPL`` stand-in returns zeros, so ``verts/kp_3d/kp_2d`` are NOT pinned by this
harness (parity unpinned for those three tensors; ``theta``/``rotmat`` do not depend on them).
"""
import collections.abc
import os
import sys
import tempfile
import types
from collections import namedtuple

import numpy as np
import torch
import torch.nn as nn

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # oracle/vendor_ref.py (GPU box)


def _find_root():
    for c in (os.environ.get("MAED_REFERENCE_ROOT"), "/root/reference", _VENDORED):
        if c and os.path.isfile(os.path.join(c, "lib", "models", "maed.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "models", "maed.py"))


_loaded = {}


def synthetic_mean_params():
    """Synthetic stand-in for data/smpl_data/smpl_mean_params.npz (licensed, absent).  Shared with the
    B200 package through oracle/synth.py so the 'iterative' decoder starts from the same buffers."""
    from oracle import synth
    return synth.mean_params()


def install_import_shims():
    """sys.modules stand-ins for what the reference imports but this image lacks (torch._six, torchvision.models.utils, yacs,
    smplx); idempotent.  Nothing under the reference tree is modified."""

    m = types.ModuleType("torch._six")
    m.container_abcs = collections.abc
    sys.modules["torch._six"] = m
    m = types.ModuleType("torchvision.models.utils")
    from torch.hub import load_state_dict_from_url
    m.load_state_dict_from_url = load_state_dict_from_url
    sys.modules["torchvision.models.utils"] = m

    class CN(dict):
        __getattr__ = dict.get

        def __setattr__(self, k, v):
            self[k] = v

        def clone(self):
            return self

    y = types.ModuleType("yacs")
    yc = types.ModuleType("yacs.config")
    yc.CfgNode = CN
    sys.modules.update({"yacs": y, "yacs.config": yc})

    ModelOutput = namedtuple("ModelOutput", "vertices joints full_pose betas global_orient body_pose",
                             defaults=[None] * 6)

    def vertices2joints(J, v):
        return torch.einsum("bik,ji->bjk", [v, J])

    class SMPL(nn.Module):  # placeholder body model: zeros (see module docstring)
        def __init__(self, *a, **k):
            super().__init__()
            self.faces = np.zeros((13776, 3), np.int64)

        def forward(self, betas=None, body_pose=None, global_orient=None, pose2rot=True, **kw):
            B = betas.shape[0]
            return ModelOutput(vertices=betas.new_zeros(B, 6890, 3), joints=betas.new_zeros(B, 45, 3),
                               betas=betas, global_orient=global_orient, body_pose=body_pose)

    sx = types.ModuleType("smplx")
    sx.SMPL = SMPL
    bm = types.ModuleType("smplx.body_models")
    bm.ModelOutput = ModelOutput
    lb = types.ModuleType("smplx.lbs")
    lb.vertices2joints = vertices2joints
    sys.modules.update({"smplx": sx, "smplx.body_models": bm, "smplx.lbs": lb})


def load_reference():
    """Returns the reference's ``lib.models`` module (with MAED), importing it once."""
    if "models" in _loaded:
        return _loaded["models"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)

    # scratch CWD with synthetic stand-ins for the licensed SMPL data files
    scratch = tempfile.mkdtemp(prefix="maed_ref_cwd_")
    os.makedirs(os.path.join(scratch, "data", "smpl_data"))
    np.save(os.path.join(scratch, "data", "smpl_data", "J_regressor_extra.npy"),
            np.zeros((9, 6890), np.float32))
    mp = synthetic_mean_params()
    np.savez(os.path.join(scratch, "data", "smpl_data", "smpl_mean_params.npz"),
             pose=mp["pose"], shape=mp["shape"], cam=mp["cam"])
    os.chdir(scratch)
    _loaded["scratch"] = scratch

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    install_import_shims()

    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lib.models.vision_transformer as vt
        import lib.models.maed as maed
        import torchvision
    _orig = vt.vit_custom_resnet50_224_in21k
    maed.vit_custom_resnet50_224_in21k = lambda nb, nh, st, **kw: _orig(nb, nh, st, pretrained=False, **kw)
    maed.resnet50 = lambda pretrained=True: torchvision.models.resnet50(weights=None)
    import lib.models as models
    _loaded["models"] = models
    return models


def build_reference_model(st_mode="parallel", decoder="ktd", num_blocks=6, num_heads=12, encoder="ste",
                          temp_frames=16):
    """Reference MAED on CPU, eval mode.  ``temp_frames`` > 16 replaces the ``temp_embed`` parameter
    DATA by a longer tensor (a data change, not a code change) so T=32 can run."""
    models = load_reference()
    os.chdir(_loaded["scratch"])          # the reference opens data/smpl_data/* relative to CWD
    model = models.MAED(encoder=encoder, num_blocks=num_blocks, num_heads=num_heads, st_mode=st_mode,
                        decoder=decoder, hidden_dim=1024).eval()
    if temp_frames != 16 and hasattr(model.encoder, "temp_embed"):
        model.encoder.temp_embed = nn.Parameter(torch.zeros(1, temp_frames, 1, 768))
    return model
