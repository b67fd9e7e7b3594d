"""Non-gating GPU canary for the parts that have not yet run on hardware (training path, SMPL tier, 'cnn' encoder, fused loss).

Their GPU tests (the ``cuda`` parametrisations of tests/test_bwd_ops.py, tests/test_smpl.py, tests/test_cnn.py,
tests/test_loss.py, tests/test_train.py) stay
behind MAED_B200_TRAIN_TESTS=1 so that a first-contact failure cannot turn the validated suite red or poison its CUDA
context.  This file runs them ONCE, last (file name), in a SUBPROCESS with a hard timeout:

  * every selected test passes  -> this test passes: the training path is confirmed on the GPU it ran on;
  * anything else (failure, crash, timeout) -> ``xfail`` with the summary line as the reason — expected-failure status for
    code whose status in DESIGN.md section 9 is "not yet validated on hardware", never an error of the suite.

The full log is written to gpurun_out/training_canary.log when that directory is writable.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_unvalidated_gpu_paths_in_a_subprocess():
    if os.environ.get("MAED_B200_TRAIN_TESTS") or os.environ.get("MAED_B200_NO_CANARY"):
        pytest.skip("the gated tests run in-process (MAED_B200_TRAIN_TESTS) or the canary is disabled")
    env = dict(os.environ, MAED_B200_TRAIN_TESTS="1", MAED_B200_NO_CANARY="1")
    # per-test limit (pytest-timeout, thread method: dumps the stacks and ends the subprocess, which is what a hung kernel
    # needs) + an overall limit well inside any sensible budget for the whole GPU suite
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", "90",
           "--timeout-method", "thread", "tests/test_bwd_ops.py", "tests/test_smpl.py", "tests/test_cnn.py", "tests/test_loss.py",
           "tests/test_train.py", "tests/test_gemm_pair.py"]          # the least certain kernel last: a hang ends the subprocess
    try:
        r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=420)
        out, code = r.stdout + r.stderr, r.returncode
    except subprocess.TimeoutExpired as e:
        out = ((e.stdout or b"").decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\nTIMEOUT after 420 s"
        code = -1
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "training_canary.log"), "w") as f:
            f.write(out)
    except OSError:
        pass
    lines = [l for l in out.strip().splitlines() if l.strip()]
    summary = lines[-1] if lines else "no output"
    failed = [l for l in lines if l.startswith(("FAILED", "ERROR"))][:8]
    print("training canary (exit %d): %s" % (code, summary))
    if code != 0:
        pytest.xfail("not-yet-validated GPU paths: %s | %s" % (summary, "; ".join(failed)))
