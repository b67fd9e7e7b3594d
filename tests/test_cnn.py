"""encoder='cnn' (torchvision ResNet-50 + decoder; reference lib/models/maed.py:35-37): inference and training.

CPU (default suite):
  * oracle/maed_oracle.py's ResNet-50 restatement against golden vectors of the UNMODIFIED reference (tests/golden/cnn_*.npz,
    made by tests/golden/make_golden.py through torchvision's own resnet50) -> the oracle is pinned;
  * the host module's state_dict keys / shapes against the reference's (stored in the same files);
  * the real cnn_engine.cu / cnn_kernels.cu sources on the CUDA-on-CPU test build (tests/emu): whole forward against the
    golden vectors, the new kernels against torch, the training step (train()-mode BatchNorm) against gradients and running
    buffers of the unmodified reference (tests/golden/grads_cnn_ktd.npz) through the C ABI and through the product modules, and
    a 2-rank gloo run of the SyncBatchNorm exchange against the reference's full-batch step.
GPU (`-m gpu`): the product library against the golden vectors and the oracle.  Green on hardware since round 2.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from helpers import build_model, load_golden, rel_err, state_dict_of
from oracle import maed_oracle as O
from oracle import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
CASES = ["cnn_ktd", "cnn_iterative"]


@pytest.fixture(scope="module")
def harness():
    import harness as h
    h.load()
    return h


def _check_outputs(o, g, tol):
    for k, gk in (("feat", "tap_feat"), ("pose6d", "tap_pose6d"), ("shape", "tap_shape"), ("cam", "tap_cam"),
                  ("rotmat", "out_rotmat"), ("theta", "out_theta"), ("kp_2d", "out_kp_2d")):
        assert rel_err(o[k].reshape(g[gk].shape), g[gk]) < tol, k


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    g, meta = load_golden(name)
    assert meta["encoder"] == "cnn"
    sd = state_dict_of(build_model(meta))
    taps = {}
    with torch.no_grad():
        out = O.maed_forward(synth.synth_frames(meta["N"], meta["T"], meta["seed"]), sd, meta["mode"], meta["decoder"], taps=taps,
                             encoder="cnn")
    for k in ("feat", "pose6d", "shape", "cam"):
        assert rel_err(taps[k], g["tap_" + k]) < 1e-5, k
    for k in ("theta", "rotmat", "kp_2d"):
        assert rel_err(out[k], g["out_" + k]) < 1e-4, k
    for k in ("stem", "stage0", "stage1", "stage2", "stage3"):
        assert rel_err(synth.tap_digest(taps[k])[0], g["dig_%s_sub" % k]) < 1e-5, k


@pytest.mark.parametrize("name", CASES)
def test_state_dict_keys_match_reference(name):
    g, meta = load_golden(name)
    ref = {str(k): [int(d) for d in str(s).split(",") if d] for k, s in zip(g["state_dict_keys"], g["state_dict_shapes"])}
    mine = {k: list(v.shape) for k, v in build_model(meta).state_dict().items()}
    assert set(mine) == set(ref), sorted(set(mine) ^ set(ref))[:10]
    for k in ref:
        assert mine[k] == ref[k], k


def test_cnn_module_surface():
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
    assert m.encoder_type == "cnn" and m.feat_dim == 2048 and m.decoder.fc1.weight.shape == (1024, 2048)
    assert sum(p.numel() for p in m.encoder.parameters()) == 23508032        # torchvision resnet50 minus fc
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.eval()(torch.zeros(1, 1, 3, 224, 224))


@pytest.mark.parametrize("name", CASES)
def test_emulated_engine_matches_reference_golden(harness, name):
    g, meta = load_golden(name)
    em = harness.EmuModel(build_model(meta))
    o = em.forward(synth.synth_frames(meta["N"], meta["T"], meta["seed"]), taps=("stem", "stage0", "stage1", "stage2", "embed"))
    _check_outputs(o, g, 2e-4)
    for mine, ref in (("stem", "stem"), ("stage0", "stage0"), ("stage1", "stage1"), ("stage2", "stage2"), ("embed", "stage3")):
        nchw = o["taps"][mine].permute(0, 3, 1, 2).contiguous()              # engine taps are NHWC
        assert rel_err(synth.tap_digest(nchw)[0], g["dig_%s_sub" % ref]) < 1e-4, mine


def test_emulated_engine_fp16_fast_mode(harness):
    """precision='fp16' (single MMA, hi planes only): the lo planes of the workspace stay unwritten (NaN-poisoned by the
    harness) and must never be read — residual planes included."""
    g, meta = load_golden("cnn_ktd")
    em = harness.EmuModel(build_model(meta, precision="fp16"))
    o = em.forward(synth.synth_frames(meta["N"], meta["T"], meta["seed"]))
    assert not torch.isnan(o["feat"]).any() and rel_err(o["feat"], g["tap_feat"]) < 2e-3


# ------------------------------------------------------------------------------------------------ training path
GRADS = "grads_cnn_ktd"        # reference in train() mode (BatchNorm on batch statistics), dropout modules in eval mode
GRADS_ITER = "grads_cnn_iterative"   # same with the iterative regressor (spin.py:51-74), 3 images


def _digest(g, nsamp=8):
    g = g.detach().double().reshape(-1).cpu()
    idx = np.unique(np.linspace(0, g.numel() - 1, nsamp).round().astype(np.int64))
    return g.norm().item(), g[torch.from_numpy(idx)].numpy()


def _check_grads(z, grads_by_name, buffers_by_name):
    """Every parameter gradient against the digests of the unmodified reference; running statistics after the step."""
    worst = ("", 0.0)
    for k in [str(s) for s in z["names"]]:
        g = grads_by_name[k]
        assert g is not None and not torch.isnan(g).any(), "gradient of %s not (fully) written" % k
        norm, samp = _digest(g)
        ref_norm = float(z["g_stats/" + k][0])
        err = abs(norm - ref_norm) / max(ref_norm, 1e-12)
        serr = np.abs(samp - z["g_samp/" + k]).max() / max(ref_norm / np.sqrt(g.numel()), 1e-20)
        if max(err, serr / 25) > worst[1]:
            worst = (k, max(err, serr / 25))
        assert err < 2e-2, "%s: |g| %.6e vs reference %.6e" % (k, norm, ref_norm)     # fp32 noise floor, see test_emu_model.py
        assert serr < 0.25, "%s: sampled entries off by %.3f rms" % (k, serr)
    for k in [str(s) for s in z["buffers"]]:
        assert rel_err(buffers_by_name[k], z["buf/" + k]) < 1e-4, k
    print("%s: worst %s %.2e" % (GRADS, worst[0], worst[1]))


def _grad_case(name=GRADS):
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    N, T, seed = [int(v) for v in z["meta"]]
    A, B, C = [synth.synth_tensor("grad_probe.%s" % k, (N * T, n), seed) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]
    return z, N, T, seed, A, B, C


@pytest.mark.parametrize("name", [GRADS, GRADS_ITER])
def test_oracle_training_gradients_match_reference(name):
    """oracle autograd with train()-mode BatchNorm == the reference's gradients and running-buffer updates."""
    z, N, T, seed, A, B, C = _grad_case(name)
    dec = str(z["decoder"])
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", dec, 1024, mean_params=synth.mean_params())
    synth.fill_module_(m, seed)
    L, g, outs = O.maed_param_grads(synth.synth_frames(N, T, seed), state_dict_of(m), A, B, C, "vanilla", dec, encoder="cnn")
    assert abs(float(L) - float(z["loss"])) < 1e-6
    for k in [str(s) for s in z["names"]]:
        assert abs(_digest(g[k])[0] - float(z["g_stats/" + k][0])) <= 1e-5 * float(z["g_stats/" + k][0]), k
    for k in [str(s) for s in z["buffers"]]:
        assert rel_err(outs["buffers"][k], z["buf/" + k]) < 1e-6, k


@pytest.mark.parametrize("name", [GRADS, GRADS_ITER])
def test_emulated_training_matches_reference_gradients(harness, name):
    """cnn_train_forward / backward (real sources on the CUDA-on-CPU build) through the C ABI: outputs, every parameter
    gradient (161 encoder + decoder tensors) and the 106 running buffers against the unmodified reference in train() mode."""
    z, N, T, seed, A, B, C = _grad_case(name)
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", str(z["decoder"]), 1024, mean_params=synth.mean_params())
    synth.fill_module_(m, seed)
    em = harness.EmuModel(m)
    out = em.train_forward(synth.synth_frames(N, T, seed))
    for k, kk in (("pose", "pose6d"), ("shape", "shape"), ("cam", "cam")):
        assert rel_err(out[kk], z["out_" + k]) < 2e-4, k
    loss = (out["pose6d"] * A).sum() + (out["shape"] * B).sum() + (out["cam"] * C).sum()
    assert abs(loss.item() - float(z["loss"])) < 1e-3 * max(1.0, abs(float(z["loss"])))
    grads = em.train_backward(A, B, C, loss_scale=4096.0)
    _check_grads(z, grads, dict(zip(em.names, em.tensors)))


def _product_training_step(device):
    """The PRODUCT modules: MAED('cnn').train().enable_training() -> loss.backward(); returns what _check_grads needs."""
    z, N, T, seed, A, B, C = _grad_case()
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
    synth.fill_module_(m, seed)
    m = m.to(device).train().enable_training(True, dropout_p=0.0)
    out = m(synth.synth_frames(N, T, seed).to(device))
    d = out["_debug"]
    loss = (d["pose6d"] * A.to(device)).sum() + (d["shape"] * B.to(device)).sum() + (d["cam"] * C.to(device)).sum()
    loss.backward()
    assert abs(loss.item() - float(z["loss"])) < 1e-3 * max(1.0, abs(float(z["loss"])))
    assert all(int(b) == 1 for n, b in m.named_buffers() if n.endswith("num_batches_tracked"))
    assert all(p.grad is not None for p in m.parameters())                  # DDP needs a gradient for every parameter
    _check_grads(z, {k: p.grad for k, p in m.named_parameters()}, dict(m.named_buffers()))
    return m


def test_product_training_step_on_the_emulator(harness):
    with harness.product_on_cpu():
        m = _product_training_step("cpu")
        from maed_b200.train import FusedAdam
        opt = FusedAdam.for_model(m, lr=1e-3, weight_decay=1e-2)
        assert opt._flat is not None and opt._flat["p"].numel() == m._train_state.flat_grad.numel()   # one launch, parameters only
        before = {n: b.clone() for n, b in m.named_buffers()}
        w0 = m.encoder.conv1.weight.detach().clone()
        opt.step()
        assert all(torch.equal(before[n], b) for n, b in m.named_buffers())
        assert not torch.equal(w0, m.encoder.conv1.weight)


def test_stage1_training_loop_on_the_emulator(harness):
    """The stage-1 loop of the reference (trainer.py:195,240-245: preds = model(images); loss = criterion(preds, target_img=...);
    loss.backward(); optimizer.step()) with the product modules only: 'cnn' encoder, fused LossImage, FusedAdam."""
    from oracle import loss_oracle as LO
    with harness.product_on_cpu():
        from maed_b200.loss import Loss
        from maed_b200.models import MAED
        from maed_b200.train import FusedAdam
        m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
        synth.fill_module_(m, 3)
        m = m.train().enable_training(True, dropout_p=0.0)
        crit = Loss(e_loss_weight=300., e_3d_loss_weight=600., e_pose_loss_weight=60., e_shape_loss_weight=0.06, device="cpu")
        opt = FusedAdam.for_model(m, lr=2e-4, weight_decay=1e-5)
        images = synth.synth_frames(3, 1, 3)                                   # image batch: T = 1 (trainer.py:177-179)
        _, target, _ = LO.synth_loss_case(0, 3, 1, 5, image=True)
        losses = []
        for _ in range(3):
            opt.zero_grad()
            loss, terms = crit(m(images), target_img=target)
            loss.backward()
            opt.step()
            losses.append(float(loss.detach()))
            assert set(terms) == {"loss_kp_2d", "loss_kp_3d", "loss_shape", "loss_pose", "loss_norm"}
        assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
        assert int(m.encoder.bn1.num_batches_tracked) == 3


def test_iterative_decoder_dropout_masks_are_replayed(harness):
    """train()-mode dropout of the iterative regressor (spin.py:66-69, six masks per step): the backward must replay the masks
    of the forward — checked by linearity: with the same seed, gradients for probes (A, B, C) and (2A, 2B, 2C) differ by
    exactly 2, and a different seed gives different outputs."""
    z, N, T, seed, A, B, C = _grad_case(GRADS_ITER)
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", "iterative", 1024, mean_params=synth.mean_params())
    synth.fill_module_(m, seed)
    x = synth.synth_frames(N, T, seed)
    key = "decoder.fc1.weight"

    def run(scale, sd, backward=True):
        em = harness.EmuModel(m)                          # (train()-mode BatchNorm normalises with batch statistics: no carry-over)
        out = em.train_forward(x, dropout_p=0.5, seed=sd)
        return out, (em.train_backward(scale * A, scale * B, scale * C, loss_scale=1024.0, dropout_p=0.5)[key] if backward else None)

    o1, g1 = run(1.0, 7)
    _, g2 = run(2.0, 7)
    o3, _ = run(1.0, 8, backward=False)
    assert torch.isfinite(g1).all() and rel_err(g2, 2.0 * g1) < 1e-5
    assert rel_err(o3["pose6d"], o1["pose6d"]) > 1e-3                     # another seed, other masks


def _syncbn_worker(rank, world, port, q):
    """One data-parallel rank: its clip of the golden batch, SyncBatchNorm exchange over gloo, gradients summed over ranks."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    os.environ.setdefault("MAED_EMU_THREADS", "4")
    torch.set_num_threads(4)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import harness
        from maed_b200.models import MAED
        z, N, T, seed, A, B, C = _grad_case()
        assert N == world
        with harness.product_on_cpu():
            m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
            synth.fill_module_(m, seed)
            m = m.train().enable_training(True, dropout_p=0.0)
            x = synth.synth_frames(N, T, seed)[rank:rank + 1]                    # this rank's clip
            rows = slice(rank * T, (rank + 1) * T)
            d = m(x)["_debug"]
            out_err = max(rel_err(d[kk], z["out_" + k][rows]) for k, kk in (("pose", "pose6d"), ("shape", "shape"), ("cam", "cam")))
            loss = (d["pose6d"] * A[rows]).sum() + (d["shape"] * B[rows]).sum() + (d["cam"] * C[rows]).sum()
            loss.backward()
            for p in m.parameters():                                              # sum (not mean): the golden loss is a sum over frames
                dist.all_reduce(p.grad)
            if rank == 0:
                _check_grads(z, {k: p.grad for k, p in m.named_parameters()}, dict(m.named_buffers()))
            q.put((rank, out_err, None))
    except Exception as e:                                                        # noqa: BLE001
        import traceback
        q.put((rank, None, traceback.format_exc()[-2000:] + repr(e)))
    finally:
        dist.destroy_process_group()


def test_syncbn_two_ranks_gloo_match_the_full_batch_reference(harness):
    """2 ranks x 1 clip with the SyncBatchNorm exchange == the reference's single-process step on the 2-clip batch: outputs of
    each rank's frames, running buffers, and the rank-summed gradients of all 161 parameters."""
    import socket

    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_syncbn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, out_err, err in res:
        assert err is None, "rank %d: %s" % (rank, err)
        assert out_err < 2e-4, (rank, out_err)


# ------------------------------------------------------------------------------------------------ per-kernel checks
def _planes(n, plane):
    return torch.zeros(2 * plane if plane else n, dtype=torch.float16)


def _run_ops(lib, stream, dev):
    """fold_bn / maxpool3x3s2 / the bottleneck GEMM epilogue against torch, on `dev` through `lib`."""
    from maed_b200 import _lib
    g = torch.Generator().manual_seed(5)
    # fold_bn
    cout, cin, k = 24, 8, 3
    w = torch.randn(cout, cin, k, k, generator=g)
    gamma, beta, mean = [torch.randn(cout, generator=g) for _ in range(3)]
    var = torch.rand(cout, generator=g) + 0.1
    t = [v.to(dev) for v in (w, gamma, beta, mean, var)]
    w_out, b_out = torch.empty_like(t[0]), torch.empty(cout, device=dev)
    assert lib.maed_op_fold_bn(_lib.ptr(t[0]), cout, cin * k * k, _lib.ptr(t[1]), _lib.ptr(t[2]), _lib.ptr(t[3]), _lib.ptr(t[4]),
                               1e-5, _lib.ptr(w_out), _lib.ptr(b_out), stream) == 0
    x = torch.randn(2, cin, 9, 9, generator=g)
    ref = F.batch_norm(F.conv2d(x, w, None, 1, 1), mean, var, gamma, beta, False, 0.0, 1e-5)
    got = F.conv2d(x, w_out.cpu(), b_out.cpu(), 1, 1)
    assert rel_err(got, ref) < 1e-5
    # maxpool: odd and even sizes, negative values so that zero padding would be wrong
    for (n, H, W, C) in ((2, 12, 10, 8), (1, 7, 9, 4)):
        x = (torch.randn(n, H, W, C, generator=g) - 2.0).to(dev)
        OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        plane = n * OH * OW * C
        o32, op = torch.empty(n, OH, OW, C, device=dev), torch.zeros(2 * plane, dtype=torch.float16, device=dev)
        assert lib.maed_op_maxpool3x3s2(_lib.ptr(x), n, H, W, C, _lib.ptr(o32), _lib.ptr(op), plane, stream) == 0
        ref = F.max_pool2d(x.cpu().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
        assert torch.equal(o32.cpu(), ref)
        rec = op[:plane].float().cpu() + op[plane:].float().cpu()
        assert rel_err(rec.reshape(ref.shape), ref) < 1e-6
    # bottleneck tail: relu(A B^T + bias + identity planes) in one GEMM launch, plane and fp32 outputs, hi-only residual
    M, N, K = 200, 128, 64
    a, b = torch.randn(M, K, generator=g), 0.1 * torch.randn(N, K, generator=g)
    bias, idt = torch.randn(N, generator=g), torch.randn(M, N, generator=g)

    def planes(t):
        hi = t.half()
        return torch.stack([hi, (t - hi.float()).half()]).contiguous().to(dev)

    pa, pb, pi = planes(a), planes(b), planes(idt)
    join = lambda p: p[0].float().cpu() + p[1].float().cpu()  # noqa: E731
    bias_d = bias.to(dev)
    ref = torch.relu(join(pa).double() @ join(pb).double().t() + bias.double() + join(pi).double())
    out = torch.zeros(2, M, N, dtype=torch.float16, device=dev)
    assert lib.maed_op_gemm_bottleneck(_lib.ptr(pa), M * K, K, _lib.ptr(pb), N * K, K, M, N, K, 3, _lib.ptr(bias_d), _lib.ptr(pi),
                                       M * N, 2, 2, _lib.ptr(out), M * N, N, stream) == 0
    assert rel_err(join(out), ref) < 3e-6
    out32 = torch.zeros(M, N, device=dev)
    assert lib.maed_op_gemm_bottleneck(_lib.ptr(pa), M * K, K, _lib.ptr(pb), N * K, K, M, N, K, 3, _lib.ptr(bias_d), _lib.ptr(pi),
                                       0, 0, 0, _lib.ptr(out32), 0, N, stream) == 0          # res_plane = 0: hi plane only, no ReLU
    ref_hi = join(pa).double() @ join(pb).double().t() + bias.double() + pi[0].float().cpu().double()
    assert rel_err(out32, ref_hi) < 3e-6


def test_cnn_kernels_emulated(harness):
    _run_ops(harness.load(), None, "cpu")


# ---------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cnn_kernels_gpu(lib):
    from maed_b200 import _lib
    _run_ops(lib, _lib.stream_ptr(), "cuda")
    torch.cuda.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cnn_forward_matches_reference_golden_gpu(name):
    g, meta = load_golden(name)
    m = build_model(meta, "cuda").eval()
    x = synth.synth_frames(meta["N"], meta["T"], meta["seed"]).cuda()
    o = m._run(x, want_taps=("stem", "stage0", "stage1", "stage2", "embed"))
    torch.cuda.synchronize()
    _check_outputs(o, g, 2e-4)                       # north_star gate: 1e-3 on pose / shape / cam; asserted tighter
    for mine, ref in (("stem", "stem"), ("stage0", "stage0"), ("stage1", "stage1"), ("stage2", "stage2"), ("embed", "stage3")):
        nchw = o["taps"][mine].permute(0, 3, 1, 2).contiguous()
        assert rel_err(synth.tap_digest(nchw)[0], g["dig_%s_sub" % ref]) < 1e-4, mine
    out = m(x)
    assert out["theta"].shape == (meta["N"], meta["T"], 85) and out["kp_3d"].shape == (meta["N"], meta["T"], 49, 3)
    assert rel_err(out["theta"].reshape(g["out_theta"].shape), g["out_theta"]) < 1e-3


@pytest.mark.gpu
def test_cnn_forward_matches_oracle_on_fresh_input_gpu():
    """bs = 2 x T = 4 (a different batch than the golden files), KTD decoder, against the pinned oracle."""
    from maed_b200.models import MAED
    m = MAED("cnn", 6, 12, "vanilla", "ktd", 1024)
    synth.fill_module_(m, 21)
    sd = state_dict_of(m)
    x = synth.synth_frames(2, 4, 21)
    taps = {}
    with torch.no_grad():
        ref = O.maed_forward(x, sd, "vanilla", "ktd", taps=taps, encoder="cnn")
    o = m.cuda().eval()(x.cuda())
    for k in ("theta", "rotmat"):
        assert rel_err(o[k], ref[k]) < 1e-3, k
    assert rel_err(m.extract_feature(x.cuda()), taps["feat"].reshape(2, 4, -1)) < 2e-4


@pytest.mark.gpu
def test_cnn_training_step_matches_reference_gradients_gpu(lib):
    _product_training_step("cuda")
