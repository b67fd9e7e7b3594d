#!/bin/bash
# Round 2, GPU call 31: weight packs spread over side streams (fork / join); train + model parity, default bench, train launch summary
mkdir -p gpurun_out
echo "=== tests"; timeout -k 5 1200 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py tests/test_train.py tests/test_cnn.py tests/test_parallel.py > gpurun_out/c31_tests.log 2>&1; echo "exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c31_tests.log | cut -c1-200 | tail -n 8
echo "=== default bench"; timeout -k 5 900 python bench.py --no-cpu-baseline > gpurun_out/c31_bench.json 2> gpurun_out/c31_bench.err; echo "exit $?"; grep -o "\"value\": [0-9.]*\|\"ms_per_step\": [0-9.]*" gpurun_out/c31_bench.json | tr "\n" " "; echo; tail -n 2 gpurun_out/c31_bench.err
echo "=== train bench, 40 steps"; timeout -k 5 900 python bench.py --mode train --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/c31_bench_train.json 2> gpurun_out/c31_bench_train.err; echo "exit $?"; grep -o "\"value\": [0-9.]*\|\"ms_per_step\": [0-9.]*" gpurun_out/c31_bench_train.json | tr "\n" " "; echo
