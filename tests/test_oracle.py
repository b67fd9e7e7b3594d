"""CPU: pins oracle/maed_oracle.py against the golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py), and the host module's state_dict against the reference's key list."""
import json
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, GOLDEN_DIR, build_model, load_golden, rel_err, state_dict_of
from oracle import maed_oracle as O
from oracle import synth

FAST_CASES = ["vanilla_ktd", "coupling_ktd", "temporal_ktd", "parallel_iterative", "series_iterative", "parallel_ktd_T1",
              "series_ktd", "c1_parallel_ktd"]


@pytest.mark.parametrize("name", FAST_CASES)
def test_oracle_matches_reference_golden(name):
    g, meta = load_golden(name)
    model = build_model(meta)
    sd = state_dict_of(model)
    x = synth.synth_frames(meta["N"], meta["T"], meta["seed"])
    taps = {}
    with torch.no_grad():
        out = O.maed_forward(x, sd, meta["mode"], meta["decoder"], taps=taps)
    # the oracle restates the same fp32 ops in the same order: agreement is at rounding level
    assert rel_err(taps["feat"], g["tap_feat"]) < 1e-5
    assert rel_err(taps["pose6d"], g["tap_pose6d"]) < 1e-5
    assert rel_err(taps["shape"], g["tap_shape"]) < 1e-5
    assert rel_err(taps["cam"], g["tap_cam"]) < 1e-5
    assert rel_err(out["theta"], g["out_theta"]) < 1e-4
    assert rel_err(out["rotmat"], g["out_rotmat"]) < 1e-4
    assert rel_err(out["kp_2d"], g["out_kp_2d"]) < 1e-4
    for k in ("stem", "stage0", "stage1", "stage2", "block0", "block5"):
        sub, stats = synth.tap_digest(taps[k])
        assert rel_err(sub, g["dig_%s_sub" % k]) < 1e-5, k
    assert float(g["out_verts_absmax"]) == 0.0 and float(g["out_kp_3d_absmax"]) == 0.0   # placeholder body model


def test_state_dict_keys_match_reference():
    """Key names and shapes of every parameter/buffer the reference's MAED has (minus decoder.smpl.*)."""
    spec = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))
    for cfg_name, ref in spec.items():
        mode, dec = cfg_name.split("/")
        from maed_b200.models import MAED
        m = MAED("ste", 6, 12, mode, dec, 1024)
        mine = {k: list(v.shape) for k, v in m.state_dict().items()}
        ref = {k: v for k, v in ref.items() if "smpl" not in k}
        assert set(mine) == set(ref), (cfg_name, sorted(set(mine) ^ set(ref))[:10])
        for k in ref:
            assert mine[k] == ref[k], (cfg_name, k, mine[k], ref[k])


def test_rotation_helpers_known_answers():
    """Identity 6-D -> identity rotation -> zero angle-axis; 90 degrees about z."""
    six = torch.tensor([[1., 0., 0., 1., 0., 0.]])
    R = O.rot6d_to_rotmat(six)
    assert torch.allclose(R[0], torch.eye(3), atol=1e-7)
    assert torch.allclose(O.rotmat_to_angle_axis(R), torch.zeros(1, 3), atol=1e-7)
    Rz = torch.tensor([[[0., -1., 0.], [1., 0., 0.], [0., 0., 1.]]])
    aa = O.rotmat_to_angle_axis(Rz)
    assert torch.allclose(aa, torch.tensor([[0., 0., np.pi / 2]]), atol=1e-6)


def test_unknown_mode_and_decoder_raise():
    from maed_b200.models import MAED
    with pytest.raises(NotImplementedError):
        MAED("ste", 6, 12, "series-parallel", "ktd")
    with pytest.raises(NotImplementedError):
        MAED("ste", 6, 12, "parallel", "gru")
    with pytest.raises(NotImplementedError):
        MAED("vit", 6, 12, "parallel", "ktd")


def test_cpu_input_fails_loudly():
    from maed_b200.models import MAED
    m = MAED("ste", 1, 12, "vanilla", "ktd")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 1, 3, 224, 224))


def test_train_mode_runs_the_training_path_by_default():
    """train()-mode forwards go to the CUDA training path (no opt-in, like the reference module); on a CPU tensor that path
    fails loudly as well, and enable_training(False) turns the call into a warned, graph-less inference call."""
    import warnings
    from maed_b200.models import MAED
    m = MAED("ste", 1, 12, "vanilla", "ktd").train()
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.zeros(1, 1, 3, 224, 224))
    assert not any("switched off" in str(x.message) for x in w)
    m.enable_training(False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(torch.zeros(1, 1, 3, 224, 224))
    assert any("switched off" in str(x.message) for x in w)
