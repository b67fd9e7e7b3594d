#!/bin/bash
# Round 2, GPU call 1: run every not-yet-validated test file un-gated, one process per file, full logs kept.
mkdir -p gpurun_out

nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
for f in test_bwd_ops test_geometry_tail test_loss test_smpl test_cnn test_gemm_pair test_train; do
  echo "=== $f"
  timeout 900 python -m pytest -q -m gpu --timeout 400 --timeout-method thread -p no:cacheprovider -rfE tests/$f.py > gpurun_out/t_$f.log 2>&1
  echo "exit $?"
  grep -E "passed|failed|^FAILED|^ERROR|Timeout" gpurun_out/t_$f.log | cut -c1-260 | tail -n 40
done
