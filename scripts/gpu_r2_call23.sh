#!/bin/bash
# Round 2, GPU call 23: division-free producers in gemm_tc / gemm_tc2 / split-K; whole -m gpu suite; default bench; train launch list
mkdir -p gpurun_out
echo "=== smoke"; timeout -k 5 600 python -c "import __graft_entry__ as g; g.smoke(); print('__SMOKE_OK__')" 2>&1 | tail -n 3
echo "=== full gpu suite"; timeout -k 5 1500 python -m pytest -q -m gpu --timeout 400 -rfE tests > gpurun_out/c23_tests.log 2>&1; echo "exit $?"; grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c23_tests.log | cut -c1-200 | tail -n 8
echo "=== default bench"; timeout -k 5 900 python bench.py > gpurun_out/c23_bench.json 2> gpurun_out/c23_bench.err; echo "exit $?"; cut -c1-330 gpurun_out/c23_bench.json; grep -o '"e2e": {[^}]*}' gpurun_out/c23_bench.json | head -n 2; grep -o '"cpu_baseline": {[^}]*}' gpurun_out/c23_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c23_bench.json | cut -c1-330; tail -n 3 gpurun_out/c23_bench.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout -k 5 1200 ncu --profile-from-start off --cache-control none --metrics $M --clock-control none --csv --log-file gpurun_out/c23_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c23_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c23_launches_train.csv > gpurun_out/c23_launches_train_summary.txt 2>&1; head -n 26 gpurun_out/c23_launches_train_summary.txt
python scripts/launch_table.py gpurun_out/c23_launches_train.csv > gpurun_out/c23_train_per_launch.txt 2>&1
