"""GPU: the CTA-pair (tcgen05 cta_group::2) GEMM of csrc/gemm2_sm100.cuh against float64 matmul and against the validated
single-CTA kernel.  The kernel is selected per process (MAED_B200_GEMM_2CTA=1, read once), so the
test body runs in a subprocess with the variable set."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BODY = r'''
import sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
from helpers import rel_err
from maed_b200 import build, ops
build.build()
import torch.nn.functional as F
def rnd(*s, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).cuda()
worst = 0.0
# (M, N, K, block_n): full pair tiles, ragged M (peer half partly / fully out of range), BLOCK_N 128 and 256, K tail
for M, N, K, bn in [(256, 256, 64, 0), (512, 128, 192, 0), (25216, 768, 768, 0), (25216, 3072, 768, 0), (25216, 768, 3072, 0),
                    (300, 256, 256, 0), (1000, 2304, 152, 0), (4096, 2304, 768, 128), (130, 128, 64, 0)]:
    a, b = rnd(M, K, seed=M + K), rnd(N, K, scale=0.05, seed=N)
    pa, pb = ops.split(a), ops.split(b)
    ref = a.double() @ b.double().t()
    for nsplit in (3, 1):
        out = ops.gemm(pa, pb, nsplit=nsplit, block_n=bn)
        r = ref if nsplit == 3 else pa[0].double() @ pb[0].double().t()
        e = rel_err(out, r)
        worst = max(worst, e)
        assert e < 2e-5, (M, N, K, bn, nsplit, e)
    bias, res = rnd(N, seed=6), rnd(M, N, seed=7)
    out = ops.gemm(pa, pb, bias=bias, act=ops.ACT_GELU, out_mode=ops.OUT_F16_SPLIT)
    tol = 3e-6 if K <= 256 else 2e-5          # fp32 tensor-core accumulation over K up to 3072, as in test_ops_gpu.test_gemm_plain
    e = rel_err(ops.join(out), F.gelu(ref + bias.double()))
    assert e < tol, ("gelu/split", M, N, K, e)
    out = ops.gemm(pa, pb, bias=bias, residual=res)
    e = rel_err(out, ref + bias.double() + res.double())
    assert e < tol, ("residual", M, N, K, e)
# implicit-GEMM convs (5-D TMA A operand): even / odd numbers of 128-row tiles, the three map sizes of the backbones
for n, H, Cin, Cout, k in [(4, 56, 64, 128, 3), (3, 28, 128, 128, 3), (5, 14, 256, 256, 3), (2, 14, 64, 128, 1)]:
    x = rnd(n, Cin, H, H, seed=8 + n)
    w = rnd(Cout, Cin, k, k, scale=0.1, seed=9)
    a = ops.split(x.permute(0, 2, 3, 1).contiguous())
    wp = ops.split(w.permute(0, 2, 3, 1).reshape(Cout, -1).contiguous())
    out = ops.conv_gemm(a, wp, k, k, k // 2, k // 2)
    ref = F.conv2d(x.double(), w.double(), padding=k // 2).permute(0, 2, 3, 1).reshape(-1, Cout)
    e = rel_err(out, ref)
    worst = max(worst, e)
    assert e < 2e-5, ("conv", n, H, Cin, Cout, k, e)
torch.cuda.synchronize()
print("PAIR_GEMM_OK worst %%.2e" %% worst)
'''


@pytest.mark.gpu
@pytest.mark.timeout(200)
def test_pair_gemm_matches_float64_in_a_subprocess():
    env = dict(os.environ, MAED_B200_GEMM_2CTA="1")
    r = subprocess.run([sys.executable, "-c", BODY % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and "PAIR_GEMM_OK" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])
