"""`Loss`, `LossVideo`, `LossImage` — drop-ins for the reference's `lib/core/loss.py` (:159-210, :214-283, :285-326) backed by
the fused CUDA loss of libmaed_b200.so (`csrc/loss.cu`, `maed_loss_forward_backward`).

Same constructor arguments, same `forward(preds, ...)` signatures, same `(total_loss, loss_dict)` return with the reference's
keys in the reference's order.  One autograd node computes every term and the gradient of the total with respect to
`preds['kp_2d']`, `preds['kp_3d']`, `preds['theta']` in three kernel launches with no host synchronisation; the entries of
`loss_dict` are detached (the reference only logs them: lib/core/trainer.py:205-226), `total_loss` carries the graph.

Observable differences: tensors must be CUDA float32 (no CPU fallback); the 49-joint layout is assumed for `kp_3d`
(pelvis = joints 27 / 28, loss.py:55-58).  Like the reference, LossImage ignores `w_smpl` (loss.py:75 masks video input only).
"""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib

TERMS = ("loss_kp_2d", "loss_kp_3d", "loss_shape", "loss_pose", "loss_norm", "loss_accl")


def _f32c(t):
    return t.to(torch.float32).contiguous()


class _FusedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kp2d, kp3d, theta, gt2, gt3, gt_theta, valid, T, weights):
        for t in (kp2d, kp3d, theta):
            if not t.is_cuda:
                raise RuntimeError("maed_b200.loss runs on CUDA only; got a %s tensor — there is no CPU fallback" % t.device)
        dev = theta.device
        kp2d, kp3d, theta = _f32c(kp2d), _f32c(kp3d), _f32c(theta)
        M2, J2 = (kp2d.shape[0], kp2d.shape[1]) if gt2 is not None else (0, 49)
        M3, J3 = theta.shape[0], kp3d.shape[1]
        gt2 = _f32c(gt2) if gt2 is not None else None
        gt3 = _f32c(gt3) if gt3 is not None else None
        gt_theta = _f32c(gt_theta)
        valid = valid.to(torch.uint8).contiguous() if valid is not None else None
        d2, d3, dt = torch.empty_like(kp2d), torch.empty_like(kp3d), torch.empty_like(theta)
        if gt2 is None:
            d2.zero_()
        losses = torch.empty(8, dtype=torch.float32, device=dev)
        w = _lib.MaedLossWeights(*weights)
        with torch.cuda.device(dev):
            lib = _lib.load()
            nbytes = lib.maed_loss_scratch_bytes(M2, M3)
            scratch = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("maed_loss_forward_backward", _lib.ptr(kp2d), _lib.ptr(gt2), M2, J2, _lib.ptr(kp3d), _lib.ptr(gt3), M3, J3,
                      _lib.ptr(theta), _lib.ptr(gt_theta), _lib.ptr(valid), T, C.byref(w), _lib.ptr(losses), _lib.ptr(d2),
                      _lib.ptr(d3), _lib.ptr(dt), _lib.ptr(scratch), C.c_size_t(nbytes), _lib.stream_ptr())
        ctx.save_for_backward(d2, d3, dt)
        terms = losses[:6].clone()
        ctx.mark_non_differentiable(terms)
        return losses[6].clone(), terms

    @staticmethod
    def backward(ctx, g_total, _g_terms):
        d2, d3, dt = ctx.saved_tensors
        return d2 * g_total, d3 * g_total, dt * g_total, None, None, None, None, None, None


def _flat(t, tail):
    return t.reshape(-1, *t.shape[-tail:])


class _LossBase(nn.Module):
    def __init__(self, device="cuda"):
        super().__init__()
        self.device = device

    def _run(self, kp2d, kp3d, theta, gt2, gt3, gt_theta, valid, T, with_smpl, with_norm, with_accl):
        w = (self.e_loss_weight, self.e_3d_loss_weight, self.e_pose_loss_weight if with_smpl else 0.0,
             self.e_shape_loss_weight if with_smpl else 0.0, self.e_smpl_norm_loss if with_norm else 0.0,
             getattr(self, "e_smpl_accl_loss", 0.0) if with_accl else 0.0)
        total, terms = _FusedLoss.apply(kp2d, kp3d, theta, gt2, gt3, gt_theta, valid, T, w)
        d = {"loss_kp_2d": terms[0], "loss_kp_3d": terms[1]}
        if with_smpl:
            d["loss_shape"], d["loss_pose"] = terms[2], terms[3]
        if with_norm:
            d["loss_norm"] = terms[4]
        if with_accl:
            d["loss_accl"] = terms[5]
        return total, d


class LossVideo(_LossBase):
    """reference loss.py:121-210."""

    def __init__(self, e_loss_weight=60., e_3d_loss_weight=30., e_pose_loss_weight=1., e_shape_loss_weight=0.001,
                 e_smpl_norm_loss=1., e_smpl_accl_loss=0., device="cuda"):
        super().__init__(device)
        self.e_loss_weight, self.e_3d_loss_weight = e_loss_weight, e_3d_loss_weight
        self.e_pose_loss_weight, self.e_shape_loss_weight = e_pose_loss_weight, e_shape_loss_weight
        self.e_smpl_norm_loss, self.e_smpl_accl_loss = e_smpl_norm_loss, e_smpl_accl_loss

    def forward(self, preds, data_3d, data_2d):
        if data_2d:
            n2 = data_2d["kp_2d"].shape[0]
            gt2 = torch.cat((data_2d["kp_2d"], data_3d["kp_2d"]), 0)
        else:
            n2, gt2 = 0, data_3d["kp_2d"]
        p2, p3, th = preds["kp_2d"], preds["kp_3d"][n2:], preds["theta"][n2:]
        T = th.shape[1]
        gt3 = data_3d["kp_3d"]
        return self._run(_flat(p2, 2), _flat(p3, 2), _flat(th, 1), _flat(gt2, 2) if len(gt2) > 0 else None,
                         _flat(gt3, 2) if len(gt3) > 0 else None, _flat(data_3d["theta"], 1), data_3d["w_smpl"].reshape(-1) != 0, T,
                         self.e_shape_loss_weight > 0 and self.e_pose_loss_weight > 0, self.e_smpl_norm_loss > 0,
                         self.e_smpl_accl_loss > 0)


class LossImage(_LossBase):
    """reference loss.py:214-283 (T = 1 predictions; w_smpl not applied, see the module docstring)."""

    def __init__(self, e_loss_weight=60., e_3d_loss_weight=600., e_pose_loss_weight=1., e_shape_loss_weight=0.001,
                 e_smpl_norm_loss=1., device="cuda"):
        super().__init__(device)
        self.e_loss_weight, self.e_3d_loss_weight = e_loss_weight, e_3d_loss_weight
        self.e_pose_loss_weight, self.e_shape_loss_weight = e_pose_loss_weight, e_shape_loss_weight
        self.e_smpl_norm_loss = e_smpl_norm_loss

    def forward(self, preds, target):
        p2, p3, th = preds["kp_2d"].squeeze(1), preds["kp_3d"].squeeze(1), preds["theta"].squeeze(1)
        gt2, gt3 = target["kp_2d"], target.get("kp_3d")
        return self._run(p2, p3, th, gt2 if len(gt2) > 0 else None, gt3 if gt3 is not None and len(gt3) > 0 else None,
                         target["theta"], None, 1, self.e_shape_loss_weight > 0 and self.e_pose_loss_weight > 0,
                         self.e_smpl_norm_loss > 0, False)


class Loss(nn.Module):
    """reference loss.py:285-326: dispatch on the keyword the trainer passes (trainer.py:188,195)."""

    def __init__(self, e_loss_weight=60., e_3d_loss_weight=30., e_pose_loss_weight=1., e_shape_loss_weight=0.001,
                 e_smpl_norm_loss=1., e_smpl_accl_loss=0., device="cuda"):
        super().__init__()
        self.loss_video = LossVideo(e_loss_weight, e_3d_loss_weight, e_pose_loss_weight, e_shape_loss_weight, e_smpl_norm_loss,
                                    e_smpl_accl_loss, device)
        self.loss_image = LossImage(e_loss_weight, e_3d_loss_weight, e_pose_loss_weight, e_shape_loss_weight, e_smpl_norm_loss,
                                    device)

    def forward(self, preds, **kwargs):
        if "target_2d" in kwargs:
            return self.loss_video(preds, kwargs["target_3d"], kwargs["target_2d"])
        if "target_img" in kwargs:
            return self.loss_image(preds, kwargs["target_img"])
        return 0, {}

    def merge_loss(self, loss_vid, loss_vid_dict, loss_img, loss_img_dict, vid_w=1.0, img_w=1.0):
        """reference loss.py:332-345: weighted merge of the video and image losses of one iteration (trainer.py:200-205)."""
        merged = {}
        for k in set(list(loss_vid_dict.keys()) + list(loss_img_dict.keys())):
            v = 0
            if k in loss_vid_dict:
                v = v + loss_vid_dict[k] * vid_w
            if k in loss_img_dict:
                v = v + loss_img_dict[k] * img_w
            merged[k] = v
        return loss_vid * vid_w + loss_img * img_w, merged
