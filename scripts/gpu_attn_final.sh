#!/bin/bash
# Evidence for the spatial-attention kernel after the TMA-store epilogue: ncu --set full in the bench step and in the
# stand-alone launch loop, plus the refreshed launch list and bench line.  .ncu-rep exported to CSV on the box.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
cap() { name=$1; shift; timeout 400 ncu --set full --clock-control none -k regex:attn_spatial "$@" > gpurun_out/ncu_$name.log 2>&1; echo "$name exit $?"; ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; rm -f gpurun_out/$name.ncu-rep; }
cap attn_inbench -s 12 -c 3 -o gpurun_out/attn_inbench -f $CMD
cap attn_standalone -s 3 -c 3 -o gpurun_out/attn_standalone -f python scripts/dbg_attn_timeline.py time
echo "=== launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-400 gpurun_out/bench.json
du -sh gpurun_out
