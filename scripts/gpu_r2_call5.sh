#!/bin/bash
# Round 2, GPU call 5 (2 GPUs): the driver's launch line at N=2 — forward replicas + nested train step with the NCCL gradient
# all-reduce (overlapped), then the A/B without overlap; plus the single-GPU suite on the new kernels (rank 0's GPU).
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
echo "=== unit + train tests (1 GPU)"; CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest -q -m gpu --timeout 400 -rfE -s tests/test_bwd_ops.py tests/test_train.py > gpurun_out/c5_tests.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|worst|median" gpurun_out/c5_tests.log | cut -c1-250 | tail -n 24
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "=== N=2 default bench"; timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c5_bench2.json 2> gpurun_out/c5_bench2.err; echo "exit $?"; tail -n 1 gpurun_out/c5_bench2.json | cut -c1-400; tail -n 1 gpurun_out/c5_bench2.json | grep -o '"train": {.*' | cut -c1-1800; tail -n 4 gpurun_out/c5_bench2.err
echo "=== N=2 train, no overlap"; timeout 900 $TR bench.py --gpus 2 --mode train --no-overlap --steps 10 --warmup 3 > gpurun_out/c5_train2_noov.json 2> gpurun_out/c5_train2_noov.err; echo "exit $?"; tail -n 1 gpurun_out/c5_train2_noov.json | cut -c1-300; tail -n 1 gpurun_out/c5_train2_noov.json | grep -o '"collective": {[^}]*}'; tail -n 3 gpurun_out/c5_train2_noov.err
echo "=== N=1 train"; CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --mode train --steps 10 --warmup 3 > gpurun_out/c5_train1.json 2> gpurun_out/c5_train1.err; echo "exit $?"; cut -c1-300 gpurun_out/c5_train1.json
echo "=== N=2 reference arm"; timeout 600 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -n 2 | cut -c1-300
