"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch ops, autograd for the gradients) of the reference's training
loss, `lib/core/loss.py` (LossVideo :159-210, LossImage :214-283, the shared terms :21-117) with `batch_rodrigues` /
`quat2mat` from `lib/utils/geometry.py:12-58`.

Only tests/ may import this file.  Pinned: tests/golden/loss_*.npz hold the loss_dict values and the gradients with respect to
the predictions produced by the UNMODIFIED reference classes on seeded inputs (tests/golden/make_golden_loss.py);
tests/test_loss.py checks this restatement against them.
"""
import torch


def batch_rodrigues(aa):
    """geometry.py:12-24 + quat2mat :27-58: (n,3) angle-axis -> (n,9) rotation matrices, row-major."""
    n = torch.norm(aa + 1e-8, p=2, dim=1, keepdim=True)
    axis = aa / n
    half = 0.5 * n
    q = torch.cat([torch.cos(half), torch.sin(half) * axis], dim=1)
    q = q / q.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1)


def keypoint_loss(pred, gt):
    """loss.py:21-38: confidence-weighted squared error, mean over every element."""
    if len(gt) == 0:
        return pred.new_zeros(())
    pred, gt = pred.reshape(-1, *pred.shape[-2:]), gt.reshape(-1, *gt.shape[-2:])
    return (gt[:, :, -1:] * (pred - gt[:, :, :-1]) ** 2).mean()


def keypoint_3d_loss(pred, gt):
    """loss.py:40-62: both skeletons centred on the mean of joints 27 and 28."""
    if len(gt) == 0:
        return pred.new_zeros(())
    pred, gt = pred.reshape(-1, *pred.shape[-2:]), gt.reshape(-1, *gt.shape[-2:])
    conf, g = gt[:, :, -1:], gt[:, :, :-1]
    g = g - ((g[:, 27] + g[:, 28]) / 2)[:, None]
    p = pred - ((pred[:, 27] + pred[:, 28]) / 2)[:, None]
    return (conf * (p - g) ** 2).mean()


def smpl_losses(pred_pose, pred_shape, gt_pose, gt_shape, w_smpl):
    """loss.py:64-93.  The validity mask is applied to VIDEO inputs only (3-D tensors); image inputs use every row."""
    if pred_pose.dim() > 2:
        m = w_smpl.reshape(-1)
        pred_pose, pred_shape = pred_pose.reshape(-1, 72)[m], pred_shape.reshape(-1, 10)[m]
        gt_pose, gt_shape = gt_pose.reshape(-1, 72)[m], gt_shape.reshape(-1, 10)[m]
    if len(pred_pose) == 0:
        return pred_pose.new_zeros(()), pred_pose.new_zeros(())
    rp, rg = batch_rodrigues(pred_pose.reshape(-1, 3)), batch_rodrigues(gt_pose.reshape(-1, 3))
    return ((rp - rg) ** 2).mean(), ((pred_shape - gt_shape) ** 2).mean()


def accl_loss(pred, gt):
    """loss.py:95-117: second differences over time, weighted by conf[t+2]^4 on both sides."""
    conf = gt[..., -1:]
    cv = conf[:, 1:] * conf[:, 1:]
    ca = cv[:, 1:] * cv[:, 1:]
    acc = lambda k: (k[:, 2:] - k[:, 1:-1]) - (k[:, 1:-1] - k[:, :-2])  # noqa: E731
    return ((acc(pred) * ca - acc(gt[..., :3]) * ca) ** 2).mean()


def loss_video(preds, data_3d, data_2d, w_kp2d=60., w_kp3d=30., w_pose=1., w_shape=0.001, w_norm=1., w_accl=0.):
    """LossVideo.forward (loss.py:159-210).  Returns (total, dict in the reference's key order)."""
    n2 = data_2d["kp_2d"].shape[0] if data_2d else 0
    gt2 = torch.cat((data_2d["kp_2d"], data_3d["kp_2d"]), 0) if data_2d else data_3d["kp_2d"]
    p3, th = preds["kp_3d"][n2:], preds["theta"][n2:]
    gt_th = data_3d["theta"]
    d = {"loss_kp_2d": w_kp2d * keypoint_loss(preds["kp_2d"], gt2), "loss_kp_3d": w_kp3d * keypoint_3d_loss(p3, data_3d["kp_3d"])}
    if w_shape > 0 and w_pose > 0:
        lp, ls = smpl_losses(th[:, :, 3:75], th[:, :, 75:], gt_th[:, :, 3:75], gt_th[:, :, 75:], data_3d["w_smpl"].bool())
        d["loss_shape"], d["loss_pose"] = ls * w_shape, lp * w_pose
    if w_norm > 0:
        d["loss_norm"] = w_norm * torch.norm(th.reshape(-1, 85)[:, 3:], p=2, dim=(0, 1)) / (th.shape[0] * th.shape[1])
    if w_accl > 0:
        d["loss_accl"] = w_accl * accl_loss(p3, data_3d["kp_3d"])
    return torch.stack(list(d.values())).sum(), d


def loss_image(preds, target, w_kp2d=60., w_kp3d=600., w_pose=1., w_shape=0.001, w_norm=1.):
    """LossImage.forward (loss.py:214-283): T = 1 predictions squeezed; w_smpl is NOT applied (2-D inputs, loss.py:75)."""
    p2, p3, th = preds["kp_2d"].squeeze(1), preds["kp_3d"].squeeze(1), preds["theta"].squeeze(1)
    gt_th = target["theta"]
    d = {"loss_kp_2d": w_kp2d * keypoint_loss(p2, target["kp_2d"]),
         "loss_kp_3d": w_kp3d * keypoint_3d_loss(p3, target["kp_3d"]) if "kp_3d" in target else th.new_zeros(())}
    if w_shape > 0 and w_pose > 0:
        lp, ls = smpl_losses(th[:, 3:75], th[:, 75:], gt_th[:, 3:75], gt_th[:, 75:], target["w_smpl"].bool())
        d["loss_shape"], d["loss_pose"] = ls * w_shape, lp * w_pose
    if w_norm > 0:
        d["loss_norm"] = w_norm * torch.norm(th[:, 3:], p=2, dim=(0, 1)) / th.shape[0]
    return torch.stack(list(d.values())).sum(), d


def synth_loss_case(n2, n3, T, seed, image=False):
    """Seeded predictions / targets with the trainer's shapes: kp_2d in [-1,1] with 0/1
    confidences, kp_3d ~ N(0,0.3), theta ~ N(0,0.2) with cam = [1,0,0], some frames without SMPL labels."""
    from oracle import synth
    g = synth._gen("loss_case_%d_%d_%d_%d" % (n2, n3, T, int(image)), seed)
    r = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    u = lambda *s: torch.rand(*s, generator=g)   # noqa: E731
    n = n2 + n3
    preds = {"kp_2d": 0.5 * r(n, T, 49, 2), "kp_3d": 0.3 * r(n, T, 49, 3), "theta": 0.2 * r(n, T, 85)}
    preds["theta"][..., 3:6] += torch.tensor([2.0, 0.5, -0.3])              # a large root rotation like real data
    preds["theta"][0, 0, 6:9] = 0.0                                          # |aa| = 0: batch_rodrigues' singular point
    def target(k):
        t = {"kp_2d": torch.cat([2 * u(k, T, 49, 2) - 1, (u(k, T, 49, 1) > 0.3).float()], -1),
             "kp_3d": torch.cat([0.3 * r(k, T, 49, 3), (u(k, T, 49, 1) > 0.2).float()], -1),
             "theta": 0.2 * r(k, T, 85), "w_smpl": (u(k, T) > 0.25).float()}
        t["theta"][..., :3] = torch.tensor([1.0, 0.0, 0.0])
        return t
    data_3d = target(n3)
    data_2d = {"kp_2d": target(n2)["kp_2d"]} if n2 else None
    if image:
        data_3d = {k: v.squeeze(1) for k, v in data_3d.items()}
    return preds, data_3d, data_2d
