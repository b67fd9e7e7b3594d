#!/bin/bash
# Round 2, GPU call 19: cluster-per-image GroupNorm forward / backward of the training path (gn_cluster.cu): unit tests both
# ways, train parity, train-step A/B against the multi-kernel path, launch list.
mkdir -p gpurun_out
echo "=== unit (cluster)"; timeout 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_bwd_ops.py > gpurun_out/c19_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c19_unit.log | cut -c1-250 | tail -n 12
echo "=== unit (multi-kernel)"; MAED_B200_GN_CLUSTER=0 timeout 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_bwd_ops.py -k groupnorm > gpurun_out/c19_unit_mk.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c19_unit_mk.log | cut -c1-250 | tail -n 6
echo "=== model + train tests"; timeout 1500 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py tests/test_train.py tests/test_cnn.py tests/test_loss.py > gpurun_out/c19_train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c19_train.log | cut -c1-250 | tail -n 12
for cl in 0 1; do
  echo "=== train bench GN_CLUSTER=$cl"; MAED_B200_GN_CLUSTER=$cl timeout 900 python bench.py --mode train --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/c19_bench_train_$cl.json 2> gpurun_out/c19_bench_train_$cl.err; echo "exit $?"
  grep -o '"ms_per_step": [0-9.]*' gpurun_out/c19_bench_train_$cl.json | head -n 2 | tr '\n' ' '; echo; tail -n 2 gpurun_out/c19_bench_train_$cl.err
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --cache-control none --metrics $M --clock-control none --csv --log-file gpurun_out/c19_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c19_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c19_launches_train.csv > gpurun_out/c19_launches_train_summary.txt 2>&1; head -n 24 gpurun_out/c19_launches_train_summary.txt
python scripts/launch_table.py gpurun_out/c19_launches_train.csv > gpurun_out/c19_train_per_launch.txt 2>&1
grep -E "gn_cluster|gn_stats|gn_apply_kernel|gn_bwd" gpurun_out/c19_train_per_launch.txt | awk '{printf "%s:%s:%.0f  ", $1, substr($2, 1, 32), $(NF-5)} END {print ""}' | fold -w 240 | head -n 30
