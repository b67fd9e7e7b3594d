#!/bin/bash
# Round 2, GPU call 13 (2 GPUs): torch DistributedDataParallel around the product module on NCCL (reference train.py:113) for both
# stages, the reference-style loop of scripts/train_synthetic.py; then the single-GPU train bench after the im2col_stem rewrite
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531"
echo "=== stage 2, torch DDP + torch.optim.Adam"; timeout 600 $TR scripts/train_synthetic.py --stage 2 --ddp --iters 8 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -n 6
echo "=== stage 1 ('cnn' + SyncBatchNorm exchange), torch DDP"; timeout 600 $TR scripts/train_synthetic.py --stage 1 --ddp --iters 8 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -n 6
echo "=== stage 2, flat path (FusedAdam + allreduce_gradients)"; timeout 600 $TR scripts/train_synthetic.py --stage 2 --iters 8 2>&1 | grep -v "OMP_NUM\|^\*\*\*\|^$" | tail -n 4
echo "=== N=1 train benches"; CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --mode train --steps 10 --warmup 3 2>/dev/null | cut -c1-300
CUDA_VISIBLE_DEVICES=1 timeout 600 python bench.py --mode train --encoder cnn --steps 10 --warmup 3 2>/dev/null | cut -c1-300
CUDA_VISIBLE_DEVICES=0 timeout 300 python -m pytest -q -m gpu tests/test_ops_gpu.py -k "im2col" 2>&1 | tail -n 2
