#!/bin/bash
# Round 2, GPU call 4: MN-major weight gradients + tcgen05 attention backward: unit tests, train parity, train bench, launch list
mkdir -p gpurun_out
echo "=== new kernels"; timeout 600 python -m pytest -q -m gpu --timeout 200 -rfE tests/test_bwd_ops.py -k "wgrad or attention_bwd" > gpurun_out/c4_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|Error|rel_err|assert" gpurun_out/c4_unit.log | cut -c1-250 | tail -n 30
echo "=== train tests"; timeout 900 python -m pytest -q -m gpu --timeout 400 -rfE -s tests/test_train.py > gpurun_out/c4_train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|worst" gpurun_out/c4_train.log | cut -c1-250 | tail -n 20
echo "=== train bench"; timeout 600 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/c4_train_bench.json 2> gpurun_out/c4_train_bench.err; echo "exit $?"; cut -c1-330 gpurun_out/c4_train_bench.json; tail -n 3 gpurun_out/c4_train_bench.err
echo "=== launch list: train step"
MAED_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c4_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c4_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c4_launches_train.csv > gpurun_out/c4_launches_train_summary.txt 2>&1; head -n 45 gpurun_out/c4_launches_train_summary.txt
