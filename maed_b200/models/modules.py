"""Parameter containers with the reference's module tree, so ``state_dict()`` keys, shapes and
initialisation match ``lib/models`` exactly.  These modules only HOLD parameters (plus
the reference's init distributions); no forward arithmetic lives here — the hot path is the CUDA engine
(``maed_b200/csrc``) driven by ``maed.py``.  Calling them directly raises.

Reference files mirrored:
  lib/models/resnetv2.py:74-93,35-49,159-274,277-348   (StdConv2dSame / GroupNormAct / Bottleneck / stem / ResNetV2)
  lib/models/vision_transformer.py:96-130,244-261,287-375 (Mlp / Attention / Block / HybridEmbed / VisionTransformer)
  lib/models/ktd.py:37-67, lib/models/spin.py:17-48       (KTD / Regressor)
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

ANCESTOR_INDEX = [
    [], [0], [0], [0], [0, 1], [0, 2], [0, 3], [0, 1, 4], [0, 2, 5], [0, 3, 6], [0, 1, 4, 7],
    [0, 2, 5, 8], [0, 3, 6, 9], [0, 3, 6, 9], [0, 3, 6, 9], [0, 3, 6, 9, 12], [0, 3, 6, 9, 13],
    [0, 3, 6, 9, 14], [0, 3, 6, 9, 13, 16], [0, 3, 6, 9, 14, 17], [0, 3, 6, 9, 13, 16, 18],
    [0, 3, 6, 9, 14, 17, 19], [0, 3, 6, 9, 13, 16, 18, 20], [0, 3, 6, 9, 14, 17, 19, 21],
]


class _Holder(nn.Module):
    """Base class: parameters only; arithmetic happens in libmaed_b200.so."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            "%s holds parameters for the B200 engine and has no PyTorch forward; call MAED.forward "
            "(there is no eager/CPU fallback)" % type(self).__name__)


class StdConv(_Holder):
    """`StdConv2dSame` parameters: weight (Cout,Cin,k,k), no bias; kaiming-normal fan_out (resnetv2.py:334-335)."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.stride, self.kernel_size = stride, k
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")


class Norm(_Holder):
    """GroupNorm / LayerNorm affine parameters (weight=1, bias=0)."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class Lin(_Holder):
    def __init__(self, cin, cout, init="trunc02"):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin))
        self.bias = nn.Parameter(torch.zeros(cout))
        if init == "trunc02":          # vision_transformer.py:366-370
            nn.init.trunc_normal_(self.weight, std=0.02)
        elif init == "xavier001":      # ktd.py:61,66-67 / spin.py:38-40
            nn.init.xavier_uniform_(self.weight, gain=0.01)
            bound = 1.0 / math.sqrt(cin)
            nn.init.uniform_(self.bias, -bound, bound)
        else:                          # nn.Linear default (ktd.py:53,55 / spin.py:31,33)
            nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
            bound = 1.0 / math.sqrt(cin)
            nn.init.uniform_(self.bias, -bound, bound)


class Downsample(_Holder):
    def __init__(self, cin, cout, stride):
        super().__init__()
        self.conv = StdConv(cin, cout, 1, stride)
        self.norm = Norm(cout)


class Bottleneck(_Holder):
    def __init__(self, cin, cout, stride, has_ds):
        super().__init__()
        mid = cout // 4
        if has_ds:
            self.downsample = Downsample(cin, cout, stride)
        self.conv1 = StdConv(cin, mid, 1)
        self.norm1 = Norm(mid)
        self.conv2 = StdConv(mid, mid, 3, stride)
        self.norm2 = Norm(mid)
        self.conv3 = StdConv(mid, cout, 1)
        self.norm3 = Norm(cout)


class Stage(_Holder):
    def __init__(self, cin, cout, stride, depth):
        super().__init__()
        self.blocks = nn.Sequential(OrderedDict(
            (str(i), Bottleneck(cin if i == 0 else cout, cout, stride if i == 0 else 1, i == 0))
            for i in range(depth)))


class ResNetV2(_Holder):
    """layers=(3,4,9), preact=False, stem_type='same' (vision_transformer.py:564-566)."""

    def __init__(self):
        super().__init__()
        self.stem = nn.Sequential(OrderedDict([("conv", StdConv(3, 64, 7, 2)), ("norm", Norm(64))]))
        self.stages = nn.Sequential(OrderedDict([
            ("0", Stage(64, 256, 1, 3)), ("1", Stage(256, 512, 2, 4)), ("2", Stage(512, 1024, 2, 9))]))


class ProjConv(_Holder):
    """HybridEmbed.proj = nn.Conv2d(1024, 768, 1) (vision_transformer.py:304) — default Conv2d init."""

    def __init__(self, cin, cout):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, cin, 1, 1))
        self.bias = nn.Parameter(torch.zeros(cout))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        bound = 1.0 / math.sqrt(cin)
        nn.init.uniform_(self.bias, -bound, bound)


class HybridEmbed(_Holder):
    def __init__(self):
        super().__init__()
        self.backbone = ResNetV2()
        self.proj = ProjConv(1024, 768)
        self.num_patches = 196


class Attention(_Holder):
    def __init__(self, dim, st_mode):
        super().__init__()
        self.proj = Lin(dim, dim)
        if st_mode == "parallel":
            self.ts_attn = Lin(2 * dim, 2 * dim)
        self.qkv = Lin(dim, 3 * dim)
        self.mode = st_mode


class Mlp(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = Lin(dim, 4 * dim)
        self.fc2 = Lin(4 * dim, dim)


class Block(_Holder):
    def __init__(self, dim, st_mode):
        super().__init__()
        self.norm1 = Norm(dim)
        self.attn = Attention(dim, st_mode)
        self.norm2 = Norm(dim)
        self.mlp = Mlp(dim)


class PreLogits(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.fc = Lin(dim, dim)


class STEncoder(_Holder):
    """`vit_custom_resnet50_224_in21k(num_blocks, num_heads, st_mode)` parameters (vision_transformer.py:560-576)."""

    def __init__(self, num_blocks, num_heads, st_mode, temp_frames=16):
        super().__init__()
        if st_mode not in ("series", "parallel", "coupling", "vanilla", "temporal"):
            raise NotImplementedError(st_mode)                     # vision_transformer.py:175
        dim = 768
        self.embed_dim = self.num_features = dim
        self.num_heads, self.st_mode = num_heads, st_mode
        self.patch_embed = HybridEmbed()
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 197, dim))
        self.blocks = nn.ModuleList([Block(dim, st_mode) for _ in range(num_blocks)])
        self.norm = Norm(dim)
        self.pre_logits = PreLogits(dim)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        if st_mode in ("coupling", "parallel", "series"):
            # the reference hard-codes 16 frames (vision_transformer.py:364); temp_frames=32 is the T=32 extension
            self.temp_embed = nn.Parameter(torch.zeros(1, temp_frames, 1, dim))
            nn.init.trunc_normal_(self.temp_embed, std=0.02)


class TVConv(_Holder):
    """Bias-free nn.Conv2d of torchvision's ResNet: kaiming-normal fan_out / relu (torchvision models/resnet.py)."""

    def __init__(self, cin, cout, k, stride=1):
        super().__init__()
        self.stride, self.kernel_size = stride, k
        self.weight = nn.Parameter(torch.empty(cout, cin, k, k))
        nn.init.kaiming_normal_(self.weight, mode="fan_out", nonlinearity="relu")


class TVBatchNorm(_Holder):
    """nn.BatchNorm2d state: affine parameters (1, 0) and the running statistics the inference path folds into the conv."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class TVBottleneck(_Holder):
    """torchvision.models.resnet.Bottleneck (expansion 4, stride on conv2); downsample = Sequential(conv1x1, bn)."""

    def __init__(self, cin, mid, stride, has_ds):
        super().__init__()
        self.conv1, self.bn1 = TVConv(cin, mid, 1), TVBatchNorm(mid)
        self.conv2, self.bn2 = TVConv(mid, mid, 3, stride), TVBatchNorm(mid)
        self.conv3, self.bn3 = TVConv(mid, 4 * mid, 1), TVBatchNorm(4 * mid)
        if has_ds:
            self.downsample = nn.Sequential(TVConv(cin, 4 * mid, 1, stride), TVBatchNorm(4 * mid))


class CNNEncoder(_Holder):
    """`torchvision.models.resnet50()` with `fc = nn.Identity()` (reference lib/models/maed.py:35-37): same module tree, so
    the state_dict keys are torchvision's (conv1, bn1, layer1..4.{i}.{conv,bn}{1,2,3}, downsample.{0,1}); 2048 features."""

    def __init__(self):
        super().__init__()
        self.num_features = 2048
        self.conv1, self.bn1 = TVConv(3, 64, 7, 2), TVBatchNorm(64)
        cin = 64
        for li, (mid, depth) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3))):
            blocks = []
            for b in range(depth):
                blocks.append(TVBottleneck(cin, mid, 2 if (li > 0 and b == 0) else 1, b == 0))
                cin = 4 * mid
            setattr(self, "layer%d" % (li + 1), nn.Sequential(*blocks))
        self.fc = nn.Identity()


# SMPL kinematic tree (parent of joint j = last entry of ANCESTOR_INDEX[j]); vertex ids of the 21 vertex-selected joints
# (smplx VertexJointSelector: face, feet, finger tips); indices of the 49 output joints in the 54-joint list
# [24 SMPL | 21 selected | 9 regressed] (= JOINT_MAP[name] for name in JOINT_NAMES, reference lib/models/smpl.py:15-55).
SMPL_PARENTS = [-1] + [a[-1] for a in ANCESTOR_INDEX[1:]]
SMPL_EXTRA_VERTEX_IDS = [332, 6260, 2800, 4071, 583, 3216, 3226, 3387, 6617, 6624, 6787,
                         2746, 2319, 2445, 2556, 2673, 6191, 5782, 5905, 6016, 6133]
SMPL_JOINT_MAP = [24, 12, 17, 19, 21, 16, 18, 20, 0, 2, 5, 8, 1, 4, 7, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                  8, 5, 45, 46, 4, 7, 21, 19, 17, 16, 18, 20, 47, 48, 49, 50, 51, 52, 53, 24, 26, 25, 28, 27]


class SMPLHead(_Holder):
    """Stand-in for `lib/models/smpl.py` (smplx.SMPL subclass).  smplx==0.1.13 and the licensed SMPL assets are absent
    here, so by default verts / kp_3d are zeros and kp_2d is the projection of zero joints — exactly what the shimmed
    reference returns in this environment.  `load_assets()` installs a body model (the arrays of SMPL_NEUTRAL.pkl +
    J_regressor_extra.npy, or the seeded synthetic pack used by the tests) as NON-persistent buffers: `state_dict()`
    keys stay those of the reference minus `decoder.smpl.*`, which every reference loader drops anyway (train.py:101,
    eval.py:29).  The forward itself runs in libmaed_b200.so (csrc/smpl.cu)."""
    n_joints = 49

    def __init__(self):
        super().__init__()
        self.has_assets = False

    def load_assets(self, assets):
        """assets: mapping with v_template [6890,3], shapedirs [6890,3,10], posedirs [207,20670], J_regressor [24,6890],
        lbs_weights [6890,24], J_regressor_extra [9,6890]; optional parents [24], extra_vertex_ids [21], joint_map [49]."""
        def f(name, shape):
            t = torch.as_tensor(assets[name], dtype=torch.float32).reshape(shape).contiguous()
            return t
        v_template, shapedirs = f("v_template", (6890, 3)), f("shapedirs", (6890, 3, 10))
        J_regressor = f("J_regressor", (24, 6890))
        bufs = {
            "v_template": v_template, "shapedirs": shapedirs.reshape(6890 * 3, 10), "posedirs": f("posedirs", (207, 6890 * 3)),
            "J_template": J_regressor @ v_template,
            "J_shapedirs": torch.einsum("jv,vkl->jkl", J_regressor, shapedirs).reshape(72, 10).contiguous(),
            "lbs_weights": f("lbs_weights", (6890, 24)), "J_regressor_extra": f("J_regressor_extra", (9, 6890)),
        }
        ints = {"parents": assets.get("parents", SMPL_PARENTS), "extra_vertex_ids": assets.get("extra_vertex_ids", SMPL_EXTRA_VERTEX_IDS),
                "joint_map": assets.get("joint_map", SMPL_JOINT_MAP)}
        for k, v in bufs.items():
            self.register_buffer(k, v, persistent=False)
        for k, v in ints.items():
            self.register_buffer(k, torch.as_tensor(v, dtype=torch.int32).reshape(-1).contiguous(), persistent=False)
        self.has_assets = True
        return self


class KTD(_Holder):
    def __init__(self, feat_dim=768, hidden_dim=1024):
        super().__init__()
        self.feat_dim = feat_dim
        self.smpl = SMPLHead()
        self.fc1 = Lin(feat_dim, hidden_dim, "default")
        self.fc2 = Lin(hidden_dim, hidden_dim, "default")
        self.joint_regs = nn.ModuleList(
            [Lin(hidden_dim + 6 * len(a), 6, "xavier001") for a in ANCESTOR_INDEX])
        self.decshape = Lin(hidden_dim, 10, "xavier001")
        self.deccam = Lin(hidden_dim, 3, "xavier001")


class Regressor(_Holder):
    def __init__(self, feat_dim=768, hidden_dim=1024, mean_params=None):
        super().__init__()
        self.smpl = SMPLHead()
        self.fc1 = Lin(feat_dim + 144 + 10 + 3, hidden_dim, "default")
        self.fc2 = Lin(hidden_dim, hidden_dim, "default")
        self.decpose = Lin(hidden_dim, 144, "xavier001")
        self.decshape = Lin(hidden_dim, 10, "xavier001")
        self.deccam = Lin(hidden_dim, 3, "xavier001")
        if mean_params is None:     # 6-D identity pose, zero shape, unit scale (stand-in for smpl_mean_params.npz)
            mean_params = {"pose": torch.tensor([1., 0., 0., 1., 0., 0.]).repeat(24).numpy(),
                           "shape": torch.zeros(10).numpy(), "cam": torch.tensor([0.9, 0., 0.]).numpy()}
        self.register_buffer("init_pose", torch.as_tensor(mean_params["pose"], dtype=torch.float32).reshape(1, 144))
        self.register_buffer("init_shape", torch.as_tensor(mean_params["shape"], dtype=torch.float32).reshape(1, 10))
        self.register_buffer("init_cam", torch.as_tensor(mean_params["cam"], dtype=torch.float32).reshape(1, 3))
