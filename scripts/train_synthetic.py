#!/usr/bin/env python
"""The reference's training iteration (lib/core/trainer.py:177-255) on synthetic data with the maed_b200 modules only —
model, criterion and optimiser are the drop-ins, the loop is the reference's:

    preds = model(inp); loss, loss_dict = criterion(preds, target_img=... | target_3d=..., target_2d=...)
    optimizer.zero_grad(); loss.backward(); optimizer.step()

    python scripts/train_synthetic.py --stage 1 --iters 20                         # 'cnn' + KTD, 128 images / GPU (config_stage1.yaml)
    python scripts/train_synthetic.py --stage 2 --iters 20                         # 'ste' parallel + KTD, 8 clips x T=16 (config_stage2.yaml)
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_synthetic.py --stage 1 --ddp

--ddp wraps the model in torch DistributedDataParallel exactly like reference train.py:113 (stage 1: the SyncBatchNorm exchange of
maed_b200 replaces convert_sync_batchnorm, train.py:95); without it the flat-buffer path (FusedAdam + allreduce_gradients) is used.
Needs a B200.
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synthetic_targets(n, T, dev, image):
    g = torch.Generator().manual_seed(5)
    ones = torch.ones(n, T, 49, 1)
    th = 0.2 * torch.randn(n, T, 85, generator=g)
    th[..., :3] = torch.tensor([1.0, 0.0, 0.0])
    t = {"kp_2d": torch.cat([2 * torch.rand(n, T, 49, 2, generator=g) - 1, ones], -1), "kp_3d": torch.cat([0.3 * torch.randn(n, T, 49, 3, generator=g), ones], -1),
         "theta": th, "w_smpl": torch.ones(n, T)}
    if image:
        t = {k: v.squeeze(1) for k, v in t.items()}
    return {k: v.to(dev) for k, v in t.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", type=int, default=2, choices=[1, 2])
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--ddp", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("train_synthetic.py: no CUDA device (the maed_b200 path has no CPU fallback)")
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from maed_b200 import build, train
    from maed_b200.loss import Loss
    from maed_b200.models import MAED
    from maed_b200 import synth                      # deterministic random-init weights / frames
    build.build()
    if args.stage == 1:
        model, n, T = MAED("cnn", 6, 12, "vanilla", "ktd", 1024), 128, 1
    else:
        model, n, T = MAED("ste", 6, 12, "parallel", "ktd", 1024), 8, 16
    synth.fill_module_(model, 0)
    model = model.to(dev).train().enable_training(True)
    criterion = Loss(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06, device=dev)
    if args.ddp and world > 1:
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], broadcast_buffers=False)
        opt = torch.optim.Adam([{"params": p, "name": k} for k, p in model.named_parameters()], lr=1e-4, weight_decay=1e-5)
    else:
        net, opt = model, train.FusedAdam.for_model(model, lr=1e-4, weight_decay=1e-5)
    x = synth.synth_frames(n, T, 100 + rank).to(dev)
    target = synthetic_targets(n, T, dev, image=(args.stage == 1))
    t0 = None
    for it in range(args.iters):
        if it == 3:
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
        preds = net(x)
        loss, loss_dict = criterion(preds, target_img=target) if args.stage == 1 else criterion(preds, target_3d=target, target_2d=None)
        opt.zero_grad()
        loss.backward()
        if net is model and world > 1:
            train.allreduce_gradients(model, world)
        opt.step()
        if rank == 0 and (it % 5 == 0 or it == args.iters - 1):
            print("iter %3d  loss %.5f  %s" % (it, loss.item(), "  ".join("%s %.4f" % (k[5:], float(v)) for k, v in loss_dict.items())))
    torch.cuda.synchronize(dev)
    if rank == 0 and t0 is not None and args.iters > 3:
        dt = (time.perf_counter() - t0) / (args.iters - 3)
        print("%.1f ms / iteration, %.1f %s/s over %d GPU(s)" % (1e3 * dt, world * n / dt, "images" if args.stage == 1 else "clips", world))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
