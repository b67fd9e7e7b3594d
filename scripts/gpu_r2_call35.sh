#!/bin/bash
# Round 2, GPU call 35: small ncu --set full captures (6 launches each) of the stage-2 fused conv+GN kernels (pair variant included)
# inside the forward step and of the cluster-per-image GroupNorm kernels inside the train step.
mkdir -p gpurun_out
MAED_BENCH_PROFILE=1 timeout -k 5 400 ncu --profile-from-start off --set full --clock-control none --kernel-name regex:gemm_gn_kernel --launch-skip 29 --launch-count 6 \
  -o gpurun_out/c35_full_gemm_gn -f python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c35_full_gemm_gn.log 2>&1; echo "ncu exit $?"
MAED_BENCH_PROFILE=1 timeout -k 5 400 ncu --profile-from-start off --set full --clock-control none --kernel-name regex:gn_cluster_kernel --launch-skip 20 --launch-count 6 \
  -o gpurun_out/c35_full_gn_cluster -f python bench.py --mode train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/c35_full_gn_cluster.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/c35_full_*.ncu-rep
