#!/bin/bash
# Round-end evidence: parity tests, smoke, bench line, launch list, ncu --set full of the hot kernels.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
echo "=== tests"; timeout 900 python -m pytest -q -m gpu --timeout 300 tests > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tests.log
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-400 gpurun_out/bench.json
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 100 -c 12 -o gpurun_out/final_gemm -f $CMD > gpurun_out/ncu_final_gemm.log 2>&1; echo "exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"attn_|stem_conv|ts_blend|layernorm|gn_apply_maxpool|im2col" -s 30 -c 12 -o gpurun_out/final_misc -f $CMD > gpurun_out/ncu_final_misc.log 2>&1; echo "exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"gemm_gn" -c 53 -o gpurun_out/final_gn -f $CMD > gpurun_out/ncu_final_gn.log 2>&1; echo "exit $?"
