#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): the driver's launch line at N=8 (forward replicas + nested train step with the overlapped NCCL
# gradient all-reduce), the configs[4] shape (bs=4, T=32) and the literal stage-1 shape ('cnn', 128 images/GPU, SyncBatchNorm exchange)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
echo "=== N=8 default bench"; timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/c9_bench8.json 2> gpurun_out/c9_bench8.err; echo "exit $?"; tail -n 1 gpurun_out/c9_bench8.json | cut -c1-330; tail -n 1 gpurun_out/c9_bench8.json | grep -o '"train": {"metric[^}]*"ms_per_step": [0-9.]*' | cut -c1-330; tail -n 1 gpurun_out/c9_bench8.json | grep -o '"collective": {[^}]*}' | cut -c1-400; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/c9_bench8.err | tail -n 4
echo "=== N=8 configs[4]: bs=4 T=32 train"; timeout 600 $TR bench.py --gpus 8 --mode train --clips 4 --seq-len 32 --steps 10 --warmup 3 > gpurun_out/c9_train8_t32.json 2> gpurun_out/c9_train8_t32.err; echo "exit $?"; tail -n 1 gpurun_out/c9_train8_t32.json | cut -c1-330; tail -n 1 gpurun_out/c9_train8_t32.json | grep -o '"collective": {[^}]*}' | cut -c1-300; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/c9_train8_t32.err | tail -n 4
echo "=== N=8 stage-1 shape: cnn train"; timeout 600 $TR bench.py --gpus 8 --mode train --encoder cnn --steps 10 --warmup 3 > gpurun_out/c9_train8_cnn.json 2> gpurun_out/c9_train8_cnn.err; echo "exit $?"; tail -n 1 gpurun_out/c9_train8_cnn.json | cut -c1-330; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/c9_train8_cnn.err | tail -n 4
