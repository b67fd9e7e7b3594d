#!/bin/bash
# Round 2, GPU call 2: the whole -m gpu suite un-gated, then first measurements of everything round 1 never measured.
mkdir -p gpurun_out
echo "=== full gpu suite"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE -s tests > gpurun_out/c2_tests.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|worst|PAIR" gpurun_out/c2_tests.log | cut -c1-300 | tail -n 30
B="--steps 10 --warmup 3 --no-cpu-baseline"
echo "=== fwd bench"; timeout 600 python bench.py $B > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; echo "exit $?"; cut -c1-200 gpurun_out/c2_bench.json
echo "=== fwd bench TMA epi"; MAED_B200_GEMM_TMA_EPI=1 timeout 600 python bench.py $B > gpurun_out/c2_bench_tmaepi.json 2> gpurun_out/c2_bench_tmaepi.err; echo "exit $?"; cut -c1-200 gpurun_out/c2_bench_tmaepi.json
echo "=== fwd bench 2CTA"; MAED_B200_GEMM_2CTA=1 timeout 600 python bench.py $B > gpurun_out/c2_bench_pair.json 2> gpurun_out/c2_bench_pair.err; echo "exit $?"; cut -c1-200 gpurun_out/c2_bench_pair.json
grep -o '"roofline": {[^}]*}' gpurun_out/c2_bench.json gpurun_out/c2_bench_tmaepi.json gpurun_out/c2_bench_pair.json | cut -c1-400
echo "=== 2CTA model parity"; MAED_B200_GEMM_2CTA=1 timeout 600 python -m pytest -q -m gpu --timeout 300 tests/test_model_gpu.py 2>&1 | tail -n 3
for mode in parallel series; do
  echo "=== train bench $mode"; timeout 900 python bench.py --mode train --st-mode $mode --steps 5 --warmup 3 > gpurun_out/c2_train_$mode.json 2> gpurun_out/c2_train_$mode.err; echo "exit $?"; cut -c1-330 gpurun_out/c2_train_$mode.json; tail -n 3 gpurun_out/c2_train_$mode.err
done
echo "=== train bench fused loss"; timeout 900 python bench.py --mode train --loss fused --steps 5 --warmup 3 > gpurun_out/c2_train_fused.json 2> gpurun_out/c2_train_fused.err; echo "exit $?"; cut -c1-330 gpurun_out/c2_train_fused.json; tail -n 3 gpurun_out/c2_train_fused.err
echo "=== cnn fwd bench"; timeout 600 python scripts/bench_cnn.py --steps 10 --warmup 3 > gpurun_out/c2_cnn.json 2> gpurun_out/c2_cnn.err; echo "exit $?"; cut -c1-330 gpurun_out/c2_cnn.json; tail -n 3 gpurun_out/c2_cnn.err
echo "=== cnn train bench"; timeout 900 python bench.py --mode train --encoder cnn --steps 5 --warmup 3 > gpurun_out/c2_train_cnn.json 2> gpurun_out/c2_train_cnn.err; echo "exit $?"; cut -c1-330 gpurun_out/c2_train_cnn.json; tail -n 3 gpurun_out/c2_train_cnn.err
echo "=== config5"; timeout 600 python scripts/bench_config5.py --steps 10 --warmup 3 > gpurun_out/c2_config5.json 2> gpurun_out/c2_config5.err; echo "exit $?"; cut -c1-330 gpurun_out/c2_config5.json; tail -n 3 gpurun_out/c2_config5.err
echo "=== train kernels"; timeout 600 python scripts/bench_train_kernels.py > gpurun_out/c2_train_kernels.log 2>&1; echo "exit $?"; tail -n 30 gpurun_out/c2_train_kernels.log | cut -c1-250
echo "=== launch list: train step"
MAED_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c2_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c2_launches_train.csv > gpurun_out/c2_launches_train_summary.txt 2>&1; head -n 50 gpurun_out/c2_launches_train_summary.txt
echo "=== launch list: fwd step"
MAED_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c2_launches_fwd.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/c2_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c2_launches_fwd.csv > gpurun_out/c2_launches_fwd_summary.txt 2>&1; head -n 40 gpurun_out/c2_launches_fwd_summary.txt
