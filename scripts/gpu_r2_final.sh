#!/bin/bash
# Round 2 final-state validation: smoke(), the whole -m gpu suite, the default bench line
mkdir -p gpurun_out
echo "=== smoke"; timeout -k 5 600 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -n 3
echo "=== full gpu suite"; timeout -k 5 1500 python -m pytest -q -m gpu --timeout 400 -rfE tests > gpurun_out/final_tests.log 2>&1; echo "exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/final_tests.log | cut -c1-200 | tail -n 8
echo "=== default bench"; timeout -k 5 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "exit $?"; cut -c1-330 gpurun_out/final_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/final_bench.json | head -n 2; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/final_bench.json | cut -c1-260; grep -o '"clocks": {[^}]*}' gpurun_out/final_bench.json | head -n 1; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/final_bench.json | cut -c1-330; tail -n 3 gpurun_out/final_bench.err
