#!/bin/bash
# Round 2, GPU call 22: cta_group::2 pair mode of the fused conv+GN kernel on the 14 x 14 layers (256-wide blocks).  Every step
# under its own short timeout: a protocol error in a persistent cluster kernel is a hang, not a crash.
mkdir -p gpurun_out
echo "=== conv_gn unit (pair)"; timeout -k 5 240 python -m pytest -q -m gpu --timeout 120 -x -rfE tests/test_ops_gpu.py -k "conv_gn" > gpurun_out/c22_unit.log 2>&1; rc=$?; echo "exit $rc"
grep -E "passed|failed|^FAILED|^ERROR|assert |Timeout" gpurun_out/c22_unit.log | cut -c1-250 | tail -n 8
if [ $rc -ne 0 ]; then nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv; echo "pair mode failed: stopping"; exit 0; fi
echo "=== model parity"; timeout -k 5 900 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_model_gpu.py tests/test_ops_gpu.py > gpurun_out/c22_model.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c22_model.log | cut -c1-250 | tail -n 8
for pr in 0 1; do
  echo "=== forward bench PAIR=$pr"
  MAED_B200_GN_PAIR=$pr timeout -k 5 600 python bench.py --no-cpu-baseline --no-train --steps 30 --warmup 5 > gpurun_out/c22_bench_$pr.json 2> gpurun_out/c22_bench_$pr.err
  echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c22_bench_$pr.json | head -n 3 | tr '\n' ' '; echo
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed
MAED_BENCH_PROFILE=1 timeout -k 5 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
  --log-file gpurun_out/c22_launches_fwd.csv python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c22_launches_fwd.log 2>&1; echo "ncu exit $?"
python scripts/launch_table.py gpurun_out/c22_launches_fwd.csv > gpurun_out/c22_fwd_per_launch.txt 2>&1
sed -n 1,62p gpurun_out/c22_fwd_per_launch.txt | cut -c1-130
python scripts/summarize_launches.py gpurun_out/c22_launches_fwd.csv > gpurun_out/c22_fwd_summary.txt 2>&1; head -n 16 gpurun_out/c22_fwd_summary.txt
echo "=== default bench (forward + nested train step)"; timeout -k 5 900 python bench.py > gpurun_out/c22_bench_default.json 2> gpurun_out/c22_bench_default.err; echo "exit $?"; cut -c1-400 gpurun_out/c22_bench_default.json; grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' gpurun_out/c22_bench_default.json | cut -c1-400; grep -o '"e2e": {[^}]*}' gpurun_out/c22_bench_default.json | head -n 2
