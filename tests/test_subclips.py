"""CPU: index logic of MAED.forward_subclips against the reference evaluator's loop + merge_sequence
(lib/core/evaluate.py:71-100,127-133), with the engine replaced by a frame-wise stand-in."""
import numpy as np
import pytest
import torch

from maed_b200.models import MAED


def _reference_loop(model_fn, images, seqlen, interp):
    """evaluate.py:71-100 restated with numpy stacking (merge_sequence)."""
    interp_len = images[:, ::interp].shape[1]
    sample_freq = interp_len // seqlen
    seqs = []
    for i in range(sample_freq):
        inp = images[:, ::interp][:, i::sample_freq]
        seqs.append(model_fn(inp)["theta"].numpy())
    seq = np.stack(seqs, axis=2)
    return seq.reshape((-1,) + seq.shape[3:])


@pytest.mark.parametrize("N,L0,seqlen,interp", [(2, 32, 16, 1), (1, 64, 16, 2), (3, 16, 16, 1), (2, 24, 4, 2)])
def test_forward_subclips_equals_reference_loop(N, L0, seqlen, interp):
    m = MAED("ste", 1, 12, "vanilla", "ktd")
    calls = []

    def fake(x, J_regressor=None, **kw):
        # a per-frame function of the pixels that also depends on the position inside the clip
        calls.append(tuple(x.shape))
        n, t = x.shape[:2]
        pos = torch.arange(t, dtype=torch.float32).reshape(1, t, 1)
        theta = x.reshape(n, t, -1)[:, :, :85] * 2.0 + pos
        return {"theta": theta, "kp_2d": theta[:, :, :98].reshape(n, t, 49, 2) if theta.shape[-1] >= 98 else theta.reshape(n, t, -1, 1)}

    m._forward_inference = fake
    images = torch.randn(N, L0, 3, 8, 8)                     # the stand-in does not care about 224x224
    got = m.forward_subclips(images, seqlen=seqlen, interp=interp)
    assert len(calls) == 1 and calls[0][0] == N * (images[:, ::interp].shape[1] // seqlen) and calls[0][1] == seqlen
    ref = _reference_loop(lambda inp: fake(inp), images, seqlen, interp)
    L = images[:, ::interp].shape[1]
    assert got["theta"].shape == (N, L, 85)
    np.testing.assert_array_equal(got["theta"].reshape(N * L, 85).numpy(), ref)


def test_forward_subclips_rejects_ragged_windows():
    m = MAED("ste", 1, 12, "vanilla", "ktd")
    with pytest.raises(ValueError, match="multiple of seqlen"):
        m.forward_subclips(torch.zeros(1, 20, 3, 8, 8), seqlen=16)
