"""Non-gating GPU canary for the parts that have not yet run on hardware (training path, SMPL tier, 'cnn' encoder, fused loss,
CTA-pair GEMM).

Their GPU tests (the ``cuda`` parametrisations of tests/test_bwd_ops.py, tests/test_smpl.py, tests/test_cnn.py,
tests/test_loss.py, tests/test_geometry_tail.py, tests/test_train.py, tests/test_gemm_pair.py) stay behind MAED_B200_TRAIN_TESTS=1 so that a first-contact
failure cannot turn the validated suite red or poison its CUDA context.  This file runs them ONCE, last (file name), each test
file in its OWN subprocess (a trap or an illegal address in one kernel leaves a sticky error in that process only) with a hard
time limit per file and overall:

  * every selected test passes  -> this test passes: those paths are confirmed on the GPU it ran on;
  * anything else (failure, crash, timeout) -> ``xfail`` with the per-file summary lines as the reason — expected-failure status
    for code whose status in DESIGN.md is "not yet validated on hardware", never an error of the suite.

The full logs are written to gpurun_out/canary_<file>.log when that directory is writable.
"""
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# most certain first; the CTA-pair GEMM (never run, hand-written cluster protocol) last
FILES = ["test_loss.py", "test_geometry_tail.py", "test_cnn.py", "test_smpl.py", "test_bwd_ops.py", "test_train.py", "test_gemm_pair.py"]
TOTAL_BUDGET_S, PER_FILE_S, PER_TEST_S = 660, 330, 200    # test_train.py runs oracle autograd on the host CPU


@pytest.mark.gpu
def test_unvalidated_gpu_paths_in_subprocesses():
    if os.environ.get("MAED_B200_TRAIN_TESTS") or os.environ.get("MAED_B200_NO_CANARY"):
        pytest.skip("the gated tests run in-process (MAED_B200_TRAIN_TESTS) or the canary is disabled")
    env = dict(os.environ, MAED_B200_TRAIN_TESTS="1", MAED_B200_NO_CANARY="1")
    t_end = time.time() + TOTAL_BUDGET_S
    report, all_ok = [], True
    for name in FILES:
        left = t_end - time.time()
        if left < 20:
            report.append("%s: not run (canary budget of %d s spent)" % (name, TOTAL_BUDGET_S))
            all_ok = False
            continue
        # per-test limit: pytest-timeout, thread method (dumps the stacks and ends that subprocess — what a hung kernel needs)
        cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", str(PER_TEST_S),
               "--timeout-method", "thread", os.path.join("tests", name)]
        try:
            r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=min(PER_FILE_S, left))
            out, code = r.stdout + r.stderr, r.returncode
        except subprocess.TimeoutExpired as e:
            out = (e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")) + "\nTIMEOUT"
            code = -1
        try:
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "canary_%s.log" % name[:-3]), "w") as f:
                f.write(out)
        except OSError:
            pass
        lines = [l for l in out.strip().splitlines() if l.strip()]
        summary = lines[-1] if lines else "no output"
        failed = [l for l in lines if l.startswith(("FAILED", "ERROR"))][:4]
        report.append("%s (exit %d): %s%s" % (name, code, summary, (" | " + "; ".join(failed)) if failed else ""))
        all_ok = all_ok and code == 0
    print("training canary:\n  " + "\n  ".join(report))
    if not all_ok:
        pytest.xfail("not-yet-validated GPU paths: " + " || ".join(report))
