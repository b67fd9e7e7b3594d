#!/bin/bash
# Round 2, GPU call 26: after reverting the input/output slot split (parity failures at 32+ images): model parity repeated, conv_gn
# unit tests at large image counts, forward bench.
mkdir -p gpurun_out
for i in 1 2 3 4 5; do
  timeout -k 5 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_model_gpu.py tests/test_ops_gpu.py -k "golden or conv_gn" > gpurun_out/c26_model_$i.log 2>&1; echo "run $i exit $?: $(grep -E 'passed|failed' gpurun_out/c26_model_$i.log | tail -n 1)"
done
echo "=== forward bench"; timeout -k 5 600 python bench.py --no-cpu-baseline --no-train --steps 30 --warmup 5 > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.err
echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c26_bench.json | head -n 3 | tr '\n' ' '; echo
