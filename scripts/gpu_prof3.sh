#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"attn_spatial" -c 2 -o gpurun_out/prof_sp2 -f $CMD > gpurun_out/ncu_sp2.log 2>&1; echo "exit $?"
