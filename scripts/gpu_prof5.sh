#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 100 -c 12 -o gpurun_out/prof_gemm -f $CMD > gpurun_out/ncu_gemm.log 2>&1; echo "exit $?"
ls -la gpurun_out/prof_gemm.ncu-rep
