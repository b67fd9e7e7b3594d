#!/bin/bash
# Round 2, GPU call 27: pair mode (128-wide blocks, 3 staging slots) on the 14 x 14 shortcut layers; many-image conv_gn cases
mkdir -p gpurun_out
echo "=== conv_gn unit"; timeout -k 5 300 python -m pytest -q -m gpu --timeout 120 -x -rfE tests/test_ops_gpu.py -k "conv_gn" > gpurun_out/c27_unit.log 2>&1; rc=$?; echo "exit $rc"
grep -E "passed|failed|^FAILED|^ERROR|assert |Timeout" gpurun_out/c27_unit.log | cut -c1-250 | tail -n 8
if [ $rc -ne 0 ]; then echo "failed: stopping"; exit 0; fi
for i in 1 2 3; do
  timeout -k 5 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_model_gpu.py > gpurun_out/c27_model_$i.log 2>&1; echo "model run $i exit $?: $(grep -E 'passed|failed' gpurun_out/c27_model_$i.log | tail -n 1)"
done
for r in 0 1; do
echo "=== forward bench PAIR_RES=$r"; MAED_B200_GN_PAIR_RES=$r timeout -k 5 600 python bench.py --no-cpu-baseline --no-train --steps 30 --warmup 5 > gpurun_out/c27_bench_$r.json 2> gpurun_out/c27_bench_$r.err
echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c27_bench_$r.json | head -n 3 | tr '\n' ' '; echo
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout -k 5 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
  --log-file gpurun_out/c27_launches_fwd.csv python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c27_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/launch_table.py gpurun_out/c27_launches_fwd.csv > gpurun_out/c27_fwd_per_launch.txt 2>&1
python scripts/summarize_launches.py gpurun_out/c27_launches_fwd.csv > gpurun_out/c27_fwd_summary.txt 2>&1; head -n 14 gpurun_out/c27_fwd_summary.txt
