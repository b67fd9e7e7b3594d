#!/bin/bash
# Round 2, GPU call 3: suite with the TMA-store epilogue as default + new train.py semantics; default bench (nested train step); reference arm
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE -x tests > gpurun_out/c3_tests.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c3_tests.log | cut -c1-300 | tail -n 10
echo "=== default bench"; ( time timeout 900 python bench.py > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err ) 2>&1 | grep real; echo "exit $?"; cut -c1-2500 gpurun_out/c3_bench.json; tail -n 5 gpurun_out/c3_bench.err
echo "=== reference arm"; ( time timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c3_ref.json 2> gpurun_out/c3_ref.err ) 2>&1 | grep real; cut -c1-1500 gpurun_out/c3_ref.json; tail -n 5 gpurun_out/c3_ref.err
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
