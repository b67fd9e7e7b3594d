#!/bin/bash
# Spatial-attention iteration loop: parity tests, tensor-pipe % of the kernel inside a bench step, short bench.
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu --timeout 300 tests/test_ops_gpu.py -k "attention" 2>&1 | tail -n 8
timeout 600 python -m pytest -q -m gpu --timeout 300 tests/test_model_gpu.py -x 2>&1 | tail -n 4
timeout 600 ncu --metrics sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:attn_spatial -s 12 -c 3 --csv --log-file gpurun_out/attn_ncu.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/attn_ncu.log 2>&1; echo "ncu exit $?"
grep -E "attn_spatial" gpurun_out/attn_ncu.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -n 12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['clocks'])"
