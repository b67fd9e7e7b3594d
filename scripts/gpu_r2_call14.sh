#!/bin/bash
# Round 2, GPU call 14: one-pass attention backward (forward row statistics + D precompute): unit tests, train parity, bench, launch list
mkdir -p gpurun_out
echo "=== unit"; timeout 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_bwd_ops.py tests/test_ops_gpu.py -k "attention" > gpurun_out/c14_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c14_unit.log | cut -c1-250 | tail -n 12
echo "=== model + train tests"; timeout 1200 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py tests/test_train.py tests/test_loss.py > gpurun_out/c14_train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c14_train.log | cut -c1-250 | tail -n 12
echo "=== default bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; echo "exit $?"; cut -c1-250 gpurun_out/c14_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c14_bench.json | cut -c1-400; tail -n 3 gpurun_out/c14_bench.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c14_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c14_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c14_launches_train.csv > gpurun_out/c14_launches_train_summary.txt 2>&1; head -n 3 gpurun_out/c14_launches_train_summary.txt; grep -E "attn|rowdot|im2col" gpurun_out/c14_launches_train_summary.txt
