#!/bin/bash
# Round 2, GPU call 15: fused conv+GN with separate A / B operand rings (one B load per K block for all tiles of an item) and
# 128-wide channel blocks on the stage-2 shortcut layers.  Parity, forward A/B, per-launch list with the L2 -> SM bytes.
mkdir -p gpurun_out
echo "=== conv_gn unit (new plan)"; timeout 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_ops_gpu.py -k "conv_gn or groupnorm or stem" > gpurun_out/c15_unit.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c15_unit.log | cut -c1-250 | tail -n 8
echo "=== conv_gn unit (old plan)"; MAED_B200_GN_BSHARED=0 MAED_B200_GN_RES_BN128=0 timeout 600 python -m pytest -q -m gpu --timeout 300 -rfE tests/test_ops_gpu.py -k "conv_gn" > gpurun_out/c15_unit_old.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|assert " gpurun_out/c15_unit_old.log | cut -c1-250 | tail -n 8
echo "=== model parity (new plan)"; timeout 900 python -m pytest -q -m gpu --timeout 400 -rfE tests/test_model_gpu.py > gpurun_out/c15_model.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/c15_model.log | cut -c1-250 | tail -n 8
for cfg in "0 0" "1 0" "1 1"; do
  set -- $cfg
  echo "=== forward bench BSHARED=$1 RES_BN128=$2"
  MAED_B200_GN_BSHARED=$1 MAED_B200_GN_RES_BN128=$2 timeout 600 python bench.py --no-cpu-baseline --no-train --steps 30 --warmup 5 > gpurun_out/c15_bench_$1$2.json 2> gpurun_out/c15_bench_$1$2.err
  echo "exit $?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/c15_bench_$1$2.json | head -n 3 | tr '\n' ' '; echo
done
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum,lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed
for cfg in "0 0" "1 1"; do
  set -- $cfg
  MAED_B200_GN_BSHARED=$1 MAED_B200_GN_RES_BN128=$2 MAED_BENCH_PROFILE=1 timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv \
    --log-file gpurun_out/c15_launches_fwd_$1$2.csv python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c15_launches_fwd_$1$2.log 2>&1; echo "ncu exit $?"
  python scripts/launch_table.py gpurun_out/c15_launches_fwd_$1$2.csv > gpurun_out/c15_fwd_per_launch_$1$2.txt 2>&1
  sed -n 1,60p gpurun_out/c15_fwd_per_launch_$1$2.txt | cut -c1-130
done
