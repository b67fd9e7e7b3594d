#!/bin/bash
# Round 2, GPU call 7: tcgen05 temporal attention (forward + backward), SMPL backward, colsum stage 2: tests, benches, launch lists
mkdir -p gpurun_out
echo "=== attention + smpl + bwd ops"; timeout 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_ops_gpu.py tests/test_bwd_ops.py tests/test_smpl.py > gpurun_out/c7_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c7_unit.log | cut -c1-250 | tail -n 30
echo "=== model + train tests"; timeout 1200 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py tests/test_train.py tests/test_loss.py tests/test_cnn.py tests/test_geometry_tail.py tests/test_subclips.py > gpurun_out/c7_model.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c7_model.log | cut -c1-250 | tail -n 20
echo "=== default bench"; timeout 900 python bench.py --no-cpu-baseline > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.err; echo "exit $?"; cut -c1-250 gpurun_out/c7_bench.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c7_bench.json | cut -c1-400; tail -n 3 gpurun_out/c7_bench.err
echo "=== A/B: CUDA-core temporal"; MAED_B200_TEMPORAL_TC=0 timeout 900 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/c7_bench_notc.json 2> gpurun_out/c7_bench_notc.err; cut -c1-250 gpurun_out/c7_bench_notc.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c7_bench_notc.json | cut -c1-400
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
echo "=== launch list: train step"
MAED_BENCH_PROFILE=1 timeout 1200 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c7_launches_train.csv \
  python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/c7_launches_train.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c7_launches_train.csv > gpurun_out/c7_launches_train_summary.txt 2>&1; head -n 36 gpurun_out/c7_launches_train_summary.txt
echo "=== launch list: fwd step"
MAED_BENCH_PROFILE=1 timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/c7_launches_fwd.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/c7_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/summarize_launches.py gpurun_out/c7_launches_fwd.csv > gpurun_out/c7_launches_fwd_summary.txt 2>&1; head -n 14 gpurun_out/c7_launches_fwd_summary.txt
