#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its CUDA context) and keeps logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 python -m pytest -q -m gpu --timeout 300 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run gemm tests/test_ops_gpu.py -k "split_roundtrip or gemm_plain or gemm_epilogues"
run conv tests/test_ops_gpu.py -k "conv_implicit or prep_conv or im2col"
run norm tests/test_ops_gpu.py -k "groupnorm or layernorm or linear_f32 or decode_outputs"
run attn_cc tests/test_ops_gpu.py -k "test_attention and (generic or temporal)"
run attn_tc tests/test_ops_gpu.py -k "test_attention and spatial or attention_spatial_plain"
run model tests/test_model_gpu.py -s
