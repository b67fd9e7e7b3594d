#!/bin/bash
# Round 2, GPU call 25: conv+GN separate input and output slots in pass 2; unit tests, timeline, forward bench, launch list
mkdir -p gpurun_out
echo "=== conv_gn unit"; timeout -k 5 240 python -m pytest -q -m gpu --timeout 120 -x -rfE tests/test_ops_gpu.py -k "conv_gn" > gpurun_out/c25_unit.log 2>&1; rc=$?; echo "exit $rc"
grep -E "passed|failed|^FAILED|^ERROR|assert |Timeout" gpurun_out/c25_unit.log | cut -c1-250 | tail -n 8
if [ $rc -ne 0 ]; then echo "failed: stopping"; exit 0; fi
echo "=== timeline"; timeout -k 5 300 python scripts/dbg_gn_timeline.py 2>&1 | grep -E "kernel|item [12]:|avg period" | cut -c1-200
echo "=== model parity"; timeout -k 5 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_model_gpu.py > gpurun_out/c25_model.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c25_model.log | cut -c1-250 | tail -n 8
for r in 0 1; do
echo "=== forward bench RES_BN128=$r"; MAED_B200_GN_RES_BN128=$r timeout -k 5 600 python bench.py --no-cpu-baseline --no-train --steps 30 --warmup 5 > gpurun_out/c25_bench_$r.json 2> gpurun_out/c25_bench_$r.err
echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c25_bench_$r.json | head -n 3 | tr '\n' ' '; echo
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout -k 5 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
  --log-file gpurun_out/c25_launches_fwd.csv python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c25_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/launch_table.py gpurun_out/c25_launches_fwd.csv > gpurun_out/c25_fwd_per_launch.txt 2>&1
python scripts/summarize_launches.py gpurun_out/c25_launches_fwd.csv > gpurun_out/c25_fwd_summary.txt 2>&1; head -n 14 gpurun_out/c25_fwd_summary.txt
