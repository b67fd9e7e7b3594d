#!/bin/bash
# Round 2 final-state validation: smoke(), the whole -m gpu suite, the default bench line and the reference arm
mkdir -p gpurun_out
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -n 4
echo "=== full gpu suite"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE tests > gpurun_out/final_tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/final_tests.log | cut -c1-200
echo "=== default bench"; timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "exit $?"; cut -c1-330 gpurun_out/final_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/final_bench.json | head -1; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/final_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/final_bench.json | cut -c1-330; tail -n 3 gpurun_out/final_bench.err
