#!/bin/bash
# Round 2, GPU call 33: ncu --set full of the fused conv+GN kernels (incl. the cta_group::2 pair variant) inside the forward step and
# of the cluster-per-image GroupNorm kernels inside the train step.  Numbers printed under ncu are not bench values.
# NOTE (result of the one run): 13 GPU-minutes and two reports of 102 + 62 MB — more than gpurun copies back (64 MiB), so nothing
# was kept.  Do not rerun as is: use --launch-count <= 8 per kernel and --section selections instead of --set full.
mkdir -p gpurun_out
MAED_BENCH_PROFILE=1 timeout -k 5 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_gn_kernel --launch-count 56 \
  -o gpurun_out/c33_full_gemm_gn -f python bench.py --no-cpu-baseline --no-train --steps 1 --warmup 3 > gpurun_out/c33_full_gemm_gn.log 2>&1; echo "ncu exit $?"
MAED_BENCH_PROFILE=1 timeout -k 5 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gn_cluster_kernel --launch-count 40 \
  -o gpurun_out/c33_full_gn_cluster -f python bench.py --mode train --no-cpu-baseline --steps 1 --warmup 3 > gpurun_out/c33_full_gn_cluster.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out/c33_full_*.ncu-rep
