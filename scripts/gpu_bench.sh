#!/bin/bash
# model parity tests, then the bench line, then the ncu launch list of the same command and full captures of the top kernels
mkdir -p gpurun_out
echo "=== model tests"; timeout 900 python -m pytest -q -m gpu --timeout 300 tests/test_model_gpu.py -s > gpurun_out/model.log 2>&1; echo "exit $?"; grep -E "^[a-z0-9_]+ \{|passed|failed|rel err" gpurun_out/model.log | tail -n 20
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cat gpurun_out/bench.json; tail -n 5 gpurun_out/bench.err
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "=== ncu full: gemm + attention"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_spatial" -s 120 -c 16 -o gpurun_out/prof_top -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "exit $?"
ls -la gpurun_out/
