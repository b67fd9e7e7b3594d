#!/bin/bash
# ncu --set full captures of the non-GEMM hot kernels (one GPU, short command)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"stem_conv|attn_spatial|attn_temporal" -c 5 -o gpurun_out/prof_attn -f $CMD > gpurun_out/ncu_attn.log 2>&1; echo "exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_gn" -c 14 -o gpurun_out/prof_gn -f $CMD > gpurun_out/ncu_gn.log 2>&1; echo "exit $?"
ls -la gpurun_out/*.ncu-rep
