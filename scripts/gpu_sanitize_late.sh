#!/bin/bash
# compute-sanitizer memcheck over the kernels added late in round 2: cluster-per-image GroupNorm (gn_cluster.cu), the fused conv+GN
# kernel with its pair mode / parallel statistics push, the rewritten GroupNorm backward statistics pass.  Short timeouts.
mkdir -p gpurun_out
echo "=== memcheck: GroupNorm forward / backward (cluster kernels)"
timeout -k 5 150 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_late_gn.log \
  python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_bwd_ops.py -k "groupnorm and cuda" > gpurun_out/sanitize_late_gn_pytest.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/sanitize_late_gn_pytest.log; grep -E "ERROR SUMMARY|Invalid|Race" gpurun_out/sanitize_late_gn.log | tail -n 4
echo "=== memcheck: fused conv+GN (pair mode, many-image cases)"
timeout -k 5 150 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_late_convgn.log \
  python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_ops_gpu.py -k "conv_gn and (7-14 or 9-14 or 40-14-1024 or 5-56 or 4-28)" > gpurun_out/sanitize_late_convgn_pytest.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/sanitize_late_convgn_pytest.log; grep -E "ERROR SUMMARY|Invalid|Race" gpurun_out/sanitize_late_convgn.log | tail -n 4
