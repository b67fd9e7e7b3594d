#!/bin/bash
# Round 2, GPU call 6: full suite on the current tree, benches, launch lists with tensor-pipe / DRAM metrics, ncu --set full
# captures of the round-2 kernels inside the bench step.
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE -s tests > gpurun_out/c6_tests.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|median|fp32 oracle" gpurun_out/c6_tests.log | cut -c1-200 | tail -n 24
echo "=== default bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; echo "exit $?"; cut -c1-250 gpurun_out/c6_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c6_bench.json | cut -c1-400; tail -n 3 gpurun_out/c6_bench.err
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
echo "=== launch list: train step"
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c6_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c6_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c6_launches_train.csv > gpurun_out/c6_launches_train_summary.txt 2>&1; head -n 40 gpurun_out/c6_launches_train_summary.txt
echo "=== launch list: fwd step"
MAED_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c6_launches_fwd.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/c6_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c6_launches_fwd.csv > gpurun_out/c6_launches_fwd_summary.txt 2>&1; head -n 14 gpurun_out/c6_launches_fwd_summary.txt
echo "=== ncu --set full: round-2 kernels inside the train step"
MAED_BENCH_PROFILE=1 timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"attn_spatial_bwd_tc|gemm_splitk_kernel<256|attn_temporal_bwd|attn_temporal_kernel|attn_spatial_tc" -c 10 -o gpurun_out/c6_full_train \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c6_full_train.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/c6_full_train.ncu-rep --page raw --csv > gpurun_out/c6_full_train_raw.csv 2>/dev/null; ls -la gpurun_out/c6_full_train*
echo "=== ncu --set full: fc1 GEMM (TMA-store epilogue) in the forward step"
MAED_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on \
  -k regex:"gemm_tc_kernel<256" -s 12 -c 3 -o gpurun_out/c6_full_gemm \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/c6_full_gemm.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/c6_full_gemm.ncu-rep --page raw --csv > gpurun_out/c6_full_gemm_raw.csv 2>/dev/null; ls -la gpurun_out/c6_full_gemm*
