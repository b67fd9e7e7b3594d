#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel<256>" -s 56 -c 10 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
timeout 1200 ncu --set full --clock-control none -k regex:"attn_|stem_conv|ts_blend|layernorm|gn_apply_maxpool|im2col" -s 30 -c 12 -o gpurun_out/prof_misc -f $CMD > gpurun_out/ncu_misc.log 2>&1; echo "exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-300 gpurun_out/bench.json
