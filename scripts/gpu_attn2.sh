#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu --timeout 300 tests/test_ops_gpu.py tests/test_model_gpu.py -x 2>&1 | tail -n 4
timeout 100 python scripts/dbg_attn_timeline.py time 2>&1 | grep ATTN_TIME
MAED_B200_ATTN_DIRECT=1 timeout 100 python scripts/dbg_attn_timeline.py time 2>&1 | grep ATTN_TIME
MAED_B200_ATTN_DBG=1 timeout 100 python scripts/dbg_attn_timeline.py timeline 2>&1 | grep ATTN_DBG | tail -16 > gpurun_out/attn_dbg.txt
MAED_B200_ATTN_DIRECT=1 MAED_B200_ATTN_DBG=1 timeout 100 python scripts/dbg_attn_timeline.py timeline 2>&1 | grep ATTN_DBG | tail -16 > gpurun_out/attn_dbg_direct.txt
M=sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
for cc in all none; do
timeout 300 ncu --metrics $M --clock-control none --cache-control $cc -k regex:attn_spatial -s 3 -c 3 --csv --log-file gpurun_out/attn_ncu_$cc.csv \
  python scripts/dbg_attn_timeline.py time > gpurun_out/attn_ncu_$cc.log 2>&1; echo "ncu cache-control=$cc exit $?"
grep -E "attn_spatial" gpurun_out/attn_ncu_$cc.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -n 6
done
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['clocks'])"
