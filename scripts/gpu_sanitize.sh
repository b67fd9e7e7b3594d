#!/bin/bash
# Round-2 triage helper: the not-yet-validated CUDA-core kernels under compute-sanitizer (memcheck, then racecheck — the one class
# of bug the CUDA-on-CPU test build cannot see: its fibers run one at a time).  Slow (10-50x): per-kernel tests only.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out

SEL="layernorm or groupnorm or batchnorm or maxpool or colsum or transpose or gelu or dropout or ktd or blend or wstd or adam or dilate or scatter"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool (backward kernels)"
  timeout 700 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_bwd_ops.py -k "$SEL" > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "exit $?"; tail -n 3 gpurun_out/sanitize_${tool}_pytest.log; grep -E "ERROR SUMMARY|Race reported|Invalid" gpurun_out/sanitize_$tool.log | tail -n 8
done
echo "=== memcheck: fused loss, geometry tail, attention backward"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck2.log \
  python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_loss.py tests/test_geometry_tail.py tests/test_bwd_ops.py -k "loss or tail or projection or attention_bwd" \
  > gpurun_out/sanitize_memcheck2_pytest.log 2>&1
echo "exit $?"; tail -n 3 gpurun_out/sanitize_memcheck2_pytest.log; grep -E "ERROR SUMMARY|Invalid" gpurun_out/sanitize_memcheck2.log | tail -n 5
