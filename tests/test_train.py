"""Model-level tests of the training path through the PRODUCT modules (maed_b200.models.MAED + maed_b200.train): tape
forward == inference forward, parameter gradients against the reference's gradient digests (tests/golden/grads_*.npz) and
against oracle autograd on fresh inputs, loss-scale invariance, FusedAdam vs torch Adam, a short fit.  Two backends (see
tests/test_bwd_ops.py):

  * ``emu``  — CPU, default suite: product Python code + real CUDA-core kernel sources + engine/train orchestration on the
    CUDA-on-CPU shim (tests/emu/harness.py::product_on_cpu); the tcgen05 kernels are contract stubs there.  The long cases
    run only with MAED_EMU_FULL=1 (the digest check of all three modes is in tests/test_emu_model.py either way);
  * ``cuda`` — `-m gpu`, the product library on a B200 (green on hardware since round 2).
"""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN_DIR, rel_err, state_dict_of
from oracle import maed_oracle as O
from oracle import synth

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))

DEV = "cuda"
_CUDA_MARKS = [pytest.mark.gpu]
_EMU_FULL = bool(os.environ.get("MAED_EMU_FULL"))


@pytest.fixture(params=["emu", pytest.param("cuda", marks=_CUDA_MARKS)])
def lib(request):
    global DEV
    if request.param == "emu":
        import harness
        DEV = "cpu"
        with harness.product_on_cpu() as l:
            yield l
    else:
        from maed_b200 import _lib, build
        build.build()
        DEV = "cuda"
        yield _lib.load()


def _slow_on_emu():
    if DEV == "cpu" and not _EMU_FULL:
        pytest.skip("long on the emulator: set MAED_EMU_FULL=1")


GRAD_CASES = ["grads_vanilla_ktd", "grads_series_ktd", "grads_parallel_ktd", "grads_vanilla_iterative", "grads_temporal_ktd",
              "grads_coupling_ktd"]


def _model(mode, seed, lib, decoder="ktd", num_blocks=6):
    from maed_b200.models import MAED
    m = MAED("ste", num_blocks, 12, mode, decoder, 1024, mean_params=synth.mean_params())
    synth.fill_module_(m, seed)
    return m.to(DEV).train().enable_training(True, dropout_p=0.0)


def _probes(nt, seed):
    return [synth.synth_tensor("grad_probe.%s" % k, (nt, n), seed).to(DEV) for k, n in (("pose", 144), ("shape", 10), ("cam", 3))]


def _loss(out, A, B, C_):
    d = out["_debug"]
    return (d["pose6d"] * A).sum() + (d["shape"] * B).sum() + (d["cam"] * C_).sum()


def _digest(g, nsamp=8):
    g = g.detach().double().reshape(-1).cpu()
    idx = np.unique(np.linspace(0, g.numel() - 1, nsamp).round().astype(np.int64))
    return g.norm().item(), g[torch.from_numpy(idx)].numpy()


@pytest.mark.parametrize("mode", ["vanilla", "series", "parallel"])
def test_tape_forward_matches_inference(lib, mode):
    if mode != "parallel":
        _slow_on_emu()
    m = _model(mode, 21, lib)
    x = synth.synth_frames(1, 3 if DEV == "cuda" else 2, 21).to(DEV)
    out = m(x)
    with torch.no_grad():
        ref = m.eval()(x, _debug=True)
    for k in ("pose6d", "shape", "cam"):
        assert rel_err(out["_debug"][k], ref["_debug"][k]) < 1e-5, k
    # the tape forward runs GroupNorm unfused (fp32 conv output kept for the backward): different summation order than inference
    assert rel_err(out["theta"], ref["theta"]) < 1e-4 and rel_err(out["rotmat"], ref["rotmat"]) < 5e-5
    assert out["theta"].requires_grad


@pytest.mark.parametrize("name", GRAD_CASES)
def test_gradients_match_reference_digests(lib, name):
    if name != "grads_series_ktd":
        _slow_on_emu()
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    N, T, seed = [int(v) for v in z["meta"]]
    m = _model(str(z["mode"]), seed, lib, str(z["decoder"]))
    A, B, C_ = _probes(N * T, seed)
    loss = _loss(m(synth.synth_frames(N, T, seed).to(DEV)), A, B, C_)
    assert abs(loss.item() - float(z["loss"])) < 1e-3 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    grads = {k: p.grad for k, p in m.named_parameters()}
    worst = ("", 0.0)
    for k in [str(s) for s in z["names"]]:
        assert grads.get(k) is not None, "no gradient for %s" % k
        norm, samp = _digest(grads[k])
        ref_norm, ref_samp = float(z["g_stats/" + k][0]), z["g_samp/" + k]
        err = abs(norm - ref_norm) / max(ref_norm, 1e-12)
        rms = ref_norm / np.sqrt(grads[k].numel())
        serr = np.abs(samp - ref_samp).max() / max(rms, 1e-20)
        if max(err, serr / 50) > worst[1]:
            worst = (k, max(err, serr / 50))
        # fp32 noise floor of these gradients: the oracle's own fp32 vs fp64 autograd differ by up to 1.7e-2 per parameter
        # (median 2e-3) on this random weight-standardised network (ReLU / arg-max flips, x80 backbone amplification)
        assert err < 2e-2, "%s: |g| %.6e vs reference %.6e" % (k, norm, ref_norm)
        assert serr < 0.25, "%s: sampled entries off by %.3f rms" % (k, serr)
    print("%s: worst %s %.2e" % (name, worst[0], worst[1]))


def _group(k):
    if "backbone" in k:
        return "backbone." + k.split("backbone.")[1].split(".")[0] + ("." + k.split("stages.")[1].split(".")[0] if "stages." in k else "")
    if ".blocks." in k:
        return "ste.blocks"
    return "decoder" if k.startswith("decoder") else "ste.other"


def test_gradients_match_oracle_autograd(lib):
    """Every entry of every parameter gradient against FLOAT64 autograd over the CPU oracle (parallel mode, 2 clips x 2
    frames).  On this random weight-standardised network the gradients of the early layers are ill-conditioned (ReLU /
    arg-max flips, x80 backbone amplification): the reference's own arithmetic in fp32 (the oracle run in float32) differs
    from float64 by ~2e-2 in stage 0.  The yardstick is therefore measured in the same test: the CUDA path must be as close
    to float64 as an fp32 run of the reference is — per layer group median <= 2 x the fp32 oracle's + 1e-3 — and, absolutely,
    overall median <= 5e-3 with no parameter above 5e-2."""
    _slow_on_emu()
    seed, N, T = 33, 2, 2
    m = _model("parallel", seed, lib)
    A, B, C_ = _probes(N * T, seed)
    x = synth.synth_frames(N, T, seed)
    _loss(m(x.to(DEV)), A, B, C_).backward()
    cpu = lambda t: t.detach().cpu()  # noqa: E731
    dbl = lambda t: cpu(t).double() if t.dtype.is_floating_point else cpu(t)  # noqa: E731
    sd32 = {k: cpu(v) for k, v in state_dict_of(m).items()}
    _, ref, _ = O.maed_param_grads(x.double(), {k: dbl(v) for k, v in sd32.items()}, dbl(A), dbl(B), dbl(C_), "parallel", "ktd")
    _, g32, _ = O.maed_param_grads(x, sd32, cpu(A), cpu(B), cpu(C_), "parallel", "ktd")
    assert next(iter(ref.values())).dtype == torch.float64
    worst, errs, groups = ("", 0.0), [], {}
    for k, p in m.named_parameters():
        e = rel_err(p.grad, ref[k])
        errs.append(e)
        groups.setdefault(_group(k), ([], []))
        groups[_group(k)][0].append(e)
        groups[_group(k)][1].append(rel_err(g32[k], ref[k]))
        if e > worst[1]:
            worst = (k, e)
        assert e < 5e-2, "%s: relative gradient error %.3e" % (k, e)
    med = float(np.median(errs))
    print("parameter gradients vs float64 oracle autograd: worst %s %.2e, median %.2e" % (worst + (med,)))
    for g, (ours, f32) in sorted(groups.items()):
        mo, mf = float(np.median(ours)), float(np.median(f32))
        print("  %-24s n=%3d median %.2e max %.2e   | fp32 oracle: median %.2e max %.2e" % (g, len(ours), mo, max(ours), mf, max(f32)))
        assert mo < 2.0 * mf + 1e-3, "%s: median relative gradient error %.3e (fp32 oracle: %.3e)" % (g, mo, mf)
    assert med < 5e-3, "median relative gradient error %.3e" % med


def test_loss_scale_invariance_and_determinism(lib):
    _slow_on_emu()          # (the emulator's copy of this check: tests/test_emu_model.py, grads_vanilla_ktd, MAED_EMU_FULL=1)
    m = _model("vanilla", 5, lib)
    nt = 2 if DEV == "cuda" else 1
    x = synth.synth_frames(1, nt, 5).to(DEV)
    A, B, C_ = _probes(nt, 5)
    gs = []
    for scale in (4096.0, 4096.0, 256.0):
        m.zero_grad(set_to_none=True)
        out = m(x)
        m._train_state.loss_scale = scale
        _loss(out, A, B, C_).backward()
        gs.append(torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone())
    assert torch.equal(gs[0], gs[1])                       # bit-reproducible
    assert rel_err(gs[2], gs[0]) < 1e-3


def test_fused_adam_matches_torch_adam(lib):
    from maed_b200.train import FusedAdam
    nb = 6 if DEV == "cuda" else 1                                             # (one STE block on the emulator: CPU suite time)
    m1, m2 = _model("vanilla", 6, lib, num_blocks=nb), _model("vanilla", 6, lib, num_blocks=nb)
    nt = 2 if DEV == "cuda" else 1
    x = synth.synth_frames(1, nt, 6).to(DEV)
    A, B, C_ = _probes(nt, 6)
    o1 = FusedAdam.for_model(m1, lr=1e-4, weight_decay=1e-5)
    o2 = torch.optim.Adam([{"params": p, "name": n} for n, p in m2.named_parameters()], lr=1e-4, weight_decay=1e-5)
    # step 1 sees identical gradients: the two optimisers must agree to rounding.  The 1e-9 parameter differences that
    # leaves are amplified to ~1e-2 in the step-2 gradients (random weight-standardised backbone, ReLU / arg-max flips), so
    # the second comparison only guards against gross errors such as stale derived weights.
    steps = ((1, 1e-7), (2, 2e-4)) if (DEV == "cuda" or _EMU_FULL) else ((1, 1e-7),)     # step 2 on the emulator: MAED_EMU_FULL=1
    for step, tol in steps:
        for m, o in ((m1, o1), (m2, o2)):
            o.zero_grad(set_to_none=True)
            _loss(m(x), A, B, C_).backward()
            o.step()
        for (k, a), (_, b) in zip(m1.named_parameters(), m2.named_parameters()):
            assert rel_err(a, b) < tol, (step, k)
        if step == 1:
            _check_adam_state_dicts(m1, o1, m2, o2)


def _check_adam_state_dicts(m1, o1, m2, o2):
    """FusedAdam.state_dict() has torch.optim.Adam's layout and content (the reference checkpoints and resumes its optimiser,
    lib/core/trainer.py:335,359), and load_state_dict() restores the flat moments and the step count."""
    from maed_b200.train import FusedAdam
    sd1, sd2 = o1.state_dict(), o2.state_dict()
    assert len(sd1["state"]) == len(sd2["state"]) == len(list(m1.parameters()))
    for i in sd2["state"]:
        a, b = sd1["state"][i], sd2["state"][i]
        assert float(a["step"]) == float(b["step"]) == 1.0
        assert rel_err(a["exp_avg"], b["exp_avg"]) < 1e-6 and rel_err(a["exp_avg_sq"], b["exp_avg_sq"]) < 1e-6
    # torch Adam's checkpoint -> a fresh flat FusedAdam: moments land in the flat buffers, the next step is step 2
    m3 = _model("vanilla", 6, None, num_blocks=len(m1.encoder.blocks))
    o3 = FusedAdam.for_model(m3, lr=1e-4, weight_decay=1e-5)
    o3.load_state_dict(sd2)
    assert o3._flat["step"] == 1
    assert rel_err(o3._flat["m"], o1._flat["m"]) < 1e-6 and rel_err(o3._flat["v"], o1._flat["v"]) < 1e-6
    p0 = o3._flat["params"][0]
    assert o3.state[p0]["exp_avg"].data_ptr() == o3._flat["m"].data_ptr()            # still views of the flat buffers
    # parameters re-allocated after for_model (model.to() / .cuda()): the one-launch step must refuse, not silently detach
    p0.data = p0.data.clone()
    with pytest.raises(RuntimeError, match="no longer live in the optimiser's flat buffer"):
        o3.step()


def test_two_forwards_one_backward_and_gradient_accumulation(lib):
    """The reference's stage-2 iteration (lib/core/trainer.py:186-202): model(video batch), model(image batch), ONE
    loss.backward() — every forward keeps its own tape and the second node accumulates.  Then the two other ways a
    gradient can already be present at backward time: zero_grad(set_to_none=False) and micro-batch accumulation."""
    m = _model("vanilla", 9, lib, num_blocks=1 if DEV == "cpu" else 6)      # (one STE block on the emulator: CPU suite time)
    x1, x2 = synth.synth_frames(1, 1, 9).to(DEV), synth.synth_frames(1, 1, 10).to(DEV)
    P1, P2 = _probes(1, 9), _probes(1, 10)
    flat = lambda: torch.cat([p.grad.reshape(-1) for p in m.parameters()]).clone()  # noqa: E731
    gs = []
    for x, P in ((x1, P1), (x2, P2)):
        m.zero_grad(set_to_none=True)
        _loss(m(x), *P).backward()
        gs.append(flat())
    st = m._train_state
    params = [p for _, p in m._train_param_order]
    # two outstanding forwards, one backward
    m.zero_grad(set_to_none=True)
    o1, o2 = m(x1), m(x2)
    (_loss(o1, *P1) + _loss(o2, *P2)).backward()
    assert torch.equal(flat(), gs[0] + gs[1])
    assert st.grads_are_flat(params)                          # p.grad still aliases the flat buffer after the accumulation
    # zero_grad(set_to_none=False): gradients present (zeros) -> the backward must ADD, not alias-and-double
    m.zero_grad(set_to_none=False)
    _loss(m(x1), *P1).backward()
    assert torch.equal(flat(), gs[0])
    # micro-batch accumulation: no zero_grad in between
    _loss(m(x2), *P2).backward()
    assert torch.equal(flat(), gs[0] + gs[1])
    assert st.grads_are_flat(params)
    # a dropped graph returns its tape to the pool; a second backward through a used graph says so
    o = m(x1)
    loss = _loss(o, *P1)
    loss.backward()
    with pytest.raises(RuntimeError):
        loss.backward()
    del o, loss
    assert len(st.tapes.free) >= 1


def test_short_fit_reduces_loss(lib):
    """A few Adam steps on one synthetic batch with the reference's parameter-space losses (theta MSE) must go down."""
    from maed_b200.train import FusedAdam
    _slow_on_emu()
    m = _model("parallel", 7, lib)
    x = synth.synth_frames(1, 4, 7).to(DEV)
    target = torch.zeros(1, 4, 85, device=DEV)
    target[..., 0] = 1.0
    opt = FusedAdam.for_model(m, lr=1e-4)
    losses = []
    for _ in range(5):
        opt.zero_grad(set_to_none=True)
        loss = ((m(x)["theta"] - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses
