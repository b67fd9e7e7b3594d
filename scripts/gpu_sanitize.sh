#!/bin/bash
# compute-sanitizer pass over the kernels written without hardware access (memcheck, then racecheck — the one class of bug the
# CUDA-on-CPU test build cannot see: its fibers run one at a time).  Slow (10-50x): per-kernel tests only.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_sanitize.sh'
mkdir -p gpurun_out
SEL="layernorm or groupnorm or batchnorm or maxpool or colsum or gelu or dropout or ktd or blend or wstd or adam or dilate or scatter"
for tool in memcheck racecheck; do
  echo "=== compute-sanitizer --tool $tool (HBM-bound backward kernels)"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --log-file gpurun_out/sanitize_$tool.log \
    python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_bwd_ops.py -k "$SEL" > gpurun_out/sanitize_${tool}_pytest.log 2>&1
  echo "exit $?"; tail -n 2 gpurun_out/sanitize_${tool}_pytest.log; grep -E "ERROR SUMMARY|Race reported|Invalid" gpurun_out/sanitize_$tool.log | tail -n 6
done
echo "=== memcheck: fused loss, geometry tail, SMPL forward / backward"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck2.log \
  python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_loss.py tests/test_geometry_tail.py tests/test_smpl.py \
  > gpurun_out/sanitize_memcheck2_pytest.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/sanitize_memcheck2_pytest.log; grep -E "ERROR SUMMARY|Invalid" gpurun_out/sanitize_memcheck2.log | tail -n 4
echo "=== memcheck: tcgen05 kernels of round 2 (attention backward, MN-major / implicit weight gradients, temporal attention)"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck3.log \
  python -m pytest -q -m gpu -x -p no:cacheprovider tests/test_bwd_ops.py tests/test_ops_gpu.py -k "(attention_bwd and (2-2-197 or 1-3-60 or 3-8-50 or 2-4-33)) or (wgrad and (64-64 or 192-96 or 1-14-64)) or (attention and temporal and 3-8-50)" \
  > gpurun_out/sanitize_memcheck3_pytest.log 2>&1
echo "exit $?"; tail -n 2 gpurun_out/sanitize_memcheck3_pytest.log; grep -E "ERROR SUMMARY|Invalid" gpurun_out/sanitize_memcheck3.log | tail -n 4
