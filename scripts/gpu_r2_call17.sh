#!/bin/bash
# Round 2, GPU call 17: ReLU mask folded into the GroupNorm backward, zig-zag image order of the GN passes (L2 reuse), merged
# gamma/beta column sums; conv+GN operand rings.  Unit tests, train parity, default bench (forward + nested train step), launch list.
mkdir -p gpurun_out
echo "=== unit"; timeout 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_bwd_ops.py tests/test_ops_gpu.py > gpurun_out/c17_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c17_unit.log | cut -c1-250 | tail -n 12
echo "=== model + train tests"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py tests/test_train.py tests/test_cnn.py > gpurun_out/c17_train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c17_train.log | cut -c1-250 | tail -n 12
echo "=== default bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c17_bench.json 2> gpurun_out/c17_bench.err; echo "exit $?"; cut -c1-250 gpurun_out/c17_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c17_bench.json | cut -c1-400; tail -n 3 gpurun_out/c17_bench.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --cache-control none --metrics $M --clock-control none --csv --log-file gpurun_out/c17_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c17_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c17_launches_train.csv > gpurun_out/c17_launches_train_summary.txt 2>&1; head -n 30 gpurun_out/c17_launches_train_summary.txt
