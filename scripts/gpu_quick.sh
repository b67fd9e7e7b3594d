#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q -m gpu --timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; grep -E "^[a-z0-9_]+ \{|passed|failed|rel err|^FAILED|^E  " gpurun_out/$name.log | tail -n 25; }
run ops tests/test_ops_gpu.py
run model tests/test_model_gpu.py -s
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ['value','ms_per_step','gpu_launches']}, d['e2e']['value'], d['roofline']['achieved'], d['clocks'])"; tail -n 5 gpurun_out/bench.err
echo "=== launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "exit $?"
python scripts/summarize_launches.py gpurun_out/launches.csv | head -30
