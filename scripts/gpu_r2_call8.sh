#!/bin/bash
# Round 2, GPU call 8: implicit-im2col weight gradient: unit test, train parity, bench, launch list; ncu --set full of the round-2 kernels
mkdir -p gpurun_out
echo "=== unit"; timeout 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_bwd_ops.py -k "wgrad" > gpurun_out/c8_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c8_unit.log | cut -c1-250 | tail -n 12
echo "=== train + cnn tests"; timeout 1200 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_train.py tests/test_cnn.py > gpurun_out/c8_train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c8_train.log | cut -c1-250 | tail -n 12
echo "=== default bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c8_bench.json 2> gpurun_out/c8_bench.err; echo "exit $?"; cut -c1-250 gpurun_out/c8_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c8_bench.json | cut -c1-400; tail -n 3 gpurun_out/c8_bench.err
for mode in series vanilla; do
  echo "=== train bench $mode"; timeout 600 python bench.py --mode train --st-mode $mode --steps 5 --warmup 3 > gpurun_out/c8_train_$mode.json 2> gpurun_out/c8_train_$mode.err; cut -c1-330 gpurun_out/c8_train_$mode.json; tail -n 2 gpurun_out/c8_train_$mode.err
done
echo "=== cnn train bench"; timeout 600 python bench.py --mode train --encoder cnn --steps 5 --warmup 3 > gpurun_out/c8_train_cnn.json 2> gpurun_out/c8_train_cnn.err; cut -c1-330 gpurun_out/c8_train_cnn.json; tail -n 2 gpurun_out/c8_train_cnn.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
echo "=== launch list: train step"
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c8_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c8_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c8_launches_train.csv > gpurun_out/c8_launches_train_summary.txt 2>&1; head -n 30 gpurun_out/c8_launches_train_summary.txt
echo "=== ncu --set full: round-2 tensor-core kernels inside the train step"
MAED_BENCH_PROFILE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"attn_bwd_tc_kernel|gemm_splitk_kernel<256|attn_temporal_tc_kernel" -c 9 -o gpurun_out/c8_full_train \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c8_full_train.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/c8_full_train.ncu-rep --page raw --csv > gpurun_out/c8_full_train_raw.csv 2>/dev/null; ls -la gpurun_out/c8_full_train* | cut -c20-
echo "=== ncu --set full: fc1 GEMM (TMA-store epilogue) in the forward step"
MAED_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"gemm_tc_kernel" -s 8 -c 4 -o gpurun_out/c8_full_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/c8_full_gemm.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/c8_full_gemm.ncu-rep --page raw --csv > gpurun_out/c8_full_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/c8_full_gemm* | cut -c20-
