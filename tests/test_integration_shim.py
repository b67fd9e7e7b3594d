"""The drop-in boundary end to end: a copy of the reference tree whose `lib/models/__init__.py` is replaced by
`integration/lib_models__init__.py` runs the reference's OWN training-side code around the B200 module —

  * `from lib.models import MAED` (train.py:27) with the constructor call of train.py:87-94,
  * `SyncBatchNorm.convert_sync_batchnorm(model).to(device)` (train.py:95), `load_state_dict(strict=False)` (train.py:101),
  * the reference's `get_optimizer` (lib/utils/utils.py:120-135: one param group per named parameter) and its `Loss`
    (lib/core/loss.py), both imported from the copied tree, unmodified,
  * one iteration of lib/core/trainer.py:177-245: video forward, image forward (`unsqueeze(1)`, T = 1), ONE
    `loss.backward()` through both graphs, `optimizer.step()`.

CPU suite: the product modules run on the CUDA-on-CPU test build (tests/emu); needs the reference tree (/root/reference in
the build container, its verbatim copy oracle/_ref elsewhere) and is skipped without it."""
import os
import shutil
import sys

import pytest
import torch

from oracle import ref_shim, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree not present")
def test_reference_training_iteration_runs_on_the_shimmed_tree(tmp_path):
    import harness
    from oracle.vendor_ref import FILES
    src = ref_shim.REFERENCE_ROOT
    for rel in FILES:                                           # the reference's files, verbatim ...
        if os.path.isfile(os.path.join(src, rel)):
            os.makedirs(os.path.dirname(tmp_path / rel), exist_ok=True)
            shutil.copyfile(os.path.join(src, rel), tmp_path / rel)
    shutil.copyfile(os.path.join(ROOT, "integration", "lib_models__init__.py"), tmp_path / "lib" / "models" / "__init__.py")   # ... but one
    ref_shim.install_import_shims()
    saved_mods = {k: v for k, v in sys.modules.items() if k == "lib" or k.startswith("lib.")}
    for k in saved_mods:
        del sys.modules[k]
    sys.path.insert(0, str(tmp_path))
    try:
        from lib.core.loss import Loss                          # the reference's loss, unmodified
        from lib.models import MAED                             # train.py:27 — resolves to the B200 module through the shim
        from lib.utils.utils import get_optimizer               # the reference's optimiser factory, unmodified
        import maed_b200.models
        assert MAED is maed_b200.models.MAED
        import lib.models.ops as ops                            # the unchanged re-export still works
        assert hasattr(ops, "DropPath") or hasattr(ops, "drop_path")
        with harness.product_on_cpu():
            device = "cpu"
            model = MAED(encoder="ste", num_blocks=1, num_heads=12, st_mode="parallel", decoder="ktd", hidden_dim=1024)
            synth.fill_module_(model, 3)
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).to(device)              # train.py:95
            ck = {"module." + k: v.clone() for k, v in model.state_dict().items()}               # a DDP-saved checkpoint
            ck = {k[len("module."):]: w for k, w in ck.items() if k.startswith("module.") and "smpl" not in k and "decoder" not in k}
            missing = model.load_state_dict(ck, strict=False)                                    # train.py:100-101
            assert all(k.startswith("decoder") for k in missing.missing_keys) and not missing.unexpected_keys
            optimizer = get_optimizer(model=model, optim_type="Adam", lr=1e-4, weight_decay=1e-5, momentum=0.9)
            assert len(optimizer.param_groups) == len(list(model.named_parameters()))
            criterion = Loss(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06,
                             e_smpl_norm_loss=1., e_smpl_accl_loss=0., device=device)
            model.train()
            model.enable_training(True, dropout_p=0.0)
            g = torch.Generator().manual_seed(5)
            def targets(n, T, image):
                th = 0.2 * torch.randn(n, T, 85, generator=g)
                th[..., :3] = torch.tensor([1.0, 0.0, 0.0])
                t = {"kp_2d": torch.cat([2 * torch.rand(n, T, 49, 2, generator=g) - 1, torch.ones(n, T, 49, 1)], -1),
                     "kp_3d": torch.cat([0.3 * torch.randn(n, T, 49, 3, generator=g), torch.ones(n, T, 49, 1)], -1),
                     "theta": th, "w_smpl": torch.ones(n, T), "w_3d": torch.ones(n, T)}
                return {k: v.squeeze(1) for k, v in t.items()} if image else t
            inp_vid = synth.synth_frames(1, 2, 3)
            inp_img = synth.synth_frames(1, 1, 4)[:, 0].unsqueeze(1)                             # trainer.py:193-195
            before = {k: p.detach().clone() for k, p in model.named_parameters()}
            # trainer.py:198-212: two forwards, the weighted sum, ONE backward
            preds = model(inp_vid)
            loss_vid, d_vid = criterion(preds=preds, target_3d=targets(1, 2, False), target_2d=None)
            preds = model(inp_img)
            loss_img, d_img = criterion(preds=preds, target_img=targets(1, 1, True))
            nt_vid, nt_img = 2, 1
            w_vid = nt_vid / (nt_img + nt_vid)
            optimizer.zero_grad()
            (loss_img * (1 - w_vid) + loss_vid * w_vid).backward()
            optimizer.step()
            total, merged = criterion.merge_loss(loss_vid, d_vid, loss_img, d_img, vid_w=w_vid, img_w=1 - w_vid)
            assert torch.isfinite(total) and set(merged) >= {"loss_kp_2d", "loss_kp_3d", "loss_norm"}
            assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
            moved = sum(int(not torch.equal(before[k], p.detach())) for k, p in model.named_parameters())
            assert moved == len(before), "%d of %d parameters updated" % (moved, len(before))
    finally:
        sys.path.remove(str(tmp_path))
        for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
            del sys.modules[k]
        sys.modules.update(saved_mods)
