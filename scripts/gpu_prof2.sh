#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_gn" -c 34 -o gpurun_out/prof_gn4 -f $CMD > gpurun_out/ncu_gn4.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
