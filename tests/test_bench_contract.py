"""bench.py contract pieces that can be checked without a GPU: both arms share one metric string, the product arm refuses to
run without a CUDA device (no CPU fallback), the clock sampler degrades to an empty record when neither NVML nor nvidia-smi
can be used."""
import importlib.util
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_single_metric_string_for_both_arms():
    b = _bench()
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert isinstance(b.METRIC, str) and "clips/sec" in b.METRIC
    # the literal appears once (the constant); both JSON lines are built from it
    assert src.count(b.METRIC) == 1 and src.count('"metric": METRIC') >= 2


def test_product_arm_fails_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and not r.stdout.strip()


def test_clock_sampler_without_nvml_or_gpu():
    if torch.cuda.is_available():
        return
    b = _bench()
    s = b.ClockSampler(0)
    s.start()
    out = s.stop()
    assert out["samples"] == 0 and out["reasons"] == [] and out["sm_mhz"] is None
