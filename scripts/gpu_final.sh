#!/bin/bash
# Round-end evidence: (optional) parity tests + smoke, bench line, launch list, ncu --set full of the hot kernels.
# The .ncu-rep files are exported to CSV on the box and deleted (gpurun copies back at most 64 MiB).
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
if [ "$1" == "tests" ]; then
  echo "=== tests"; timeout 900 python -m pytest -q -m gpu --timeout 300 tests > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tests.log
  echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
fi
echo "=== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; cut -c1-300 gpurun_out/bench.json
echo "=== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
echo "=== ncu full"
cap() { name=$1; shift; timeout 900 ncu --set full --clock-control none "$@" -o gpurun_out/$name -f $CMD > gpurun_out/ncu_$name.log 2>&1; echo "exit $?"; ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null; rm -f gpurun_out/$name.ncu-rep; }
cap final_gemm -k regex:"gemm_tc_kernel" -s 100 -c 12
cap final_misc -k regex:"attn_|stem_conv|ts_blend|layernorm|gn_apply_maxpool|im2col" -s 30 -c 12
cap final_gn -k regex:"gemm_gn" -c 53
du -sh gpurun_out
