#!/bin/bash
# First GPU call of round 2: (1) the validated suite must still be green, (2) run the unvalidated training-path tests and
# keep their full logs, (3) re-capture the in-step ncu numbers of the spatial attention kernel (fp32 TMA-store epilogue).
mkdir -p gpurun_out
echo "=== validated suite"; MAED_B200_NO_CANARY=1 timeout 900 python -m pytest -q -m gpu --timeout 300 tests > gpurun_out/tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tests.log
export MAED_B200_TRAIN_TESTS=1
echo "=== backward kernels"; timeout 900 python -m pytest -q -m gpu --timeout 300 tests/test_bwd_ops.py > gpurun_out/bwd_ops.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/bwd_ops.log | tail -n 45
echo "=== SMPL tier"; timeout 600 python -m pytest -q -m gpu --timeout 300 tests/test_smpl.py > gpurun_out/smpl.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/smpl.log
echo "=== 'cnn' encoder"; timeout 600 python -m pytest -q -m gpu --timeout 300 -s tests/test_cnn.py > gpurun_out/cnn.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/cnn.log
echo "=== geometry tail"; timeout 300 python -m pytest -q -m gpu --timeout 120 tests/test_geometry_tail.py > gpurun_out/geometry.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/geometry.log
echo "=== fused loss"; timeout 300 python -m pytest -q -m gpu --timeout 120 tests/test_loss.py > gpurun_out/loss.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/loss.log
echo "=== training path"; timeout 1200 python -m pytest -q -m gpu --timeout 600 -s tests/test_train.py > gpurun_out/train.log 2>&1; echo "exit $?"
grep -E "passed|failed|^FAILED|^ERROR|worst" gpurun_out/train.log | tail -n 30
echo "=== CTA-pair GEMM (cta_group::2)"; timeout 400 python -m pytest -q -m gpu --timeout 300 tests/test_gemm_pair.py > gpurun_out/gemm_pair.log 2>&1; echo "exit $?"; tail -n 12 gpurun_out/gemm_pair.log
unset MAED_B200_TRAIN_TESTS
M=sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum
timeout 400 ncu --metrics $M --clock-control none -k regex:attn_spatial -s 12 -c 3 --csv --log-file gpurun_out/attn_inbench.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/attn_inbench.log 2>&1; echo "ncu exit $?"
grep -E "attn_spatial" gpurun_out/attn_inbench.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -n 6
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench.json
# (4) opt-in TMA-store epilogue of gemm_tc_kernel: parity of every GEMM / model test, then A/B bench
echo "=== GEMM TMA-store epilogue"
MAED_B200_GEMM_TMA_EPI=1 timeout 900 python -m pytest -q -m gpu --timeout 300 tests/test_ops_gpu.py tests/test_model_gpu.py > gpurun_out/tma_epi_tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/tma_epi_tests.log
MAED_B200_GEMM_TMA_EPI=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tma_epi.json 2> gpurun_out/bench_tma_epi.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_tma_epi.json
# (5) CTA-pair GEMM: only if its parity test passed -> whole-model parity, then A/B bench
if grep -q "1 passed" gpurun_out/gemm_pair.log; then
  MAED_B200_GEMM_2CTA=1 timeout 900 python -m pytest -q -m gpu --timeout 300 tests/test_model_gpu.py > gpurun_out/pair_model_tests.log 2>&1; echo "exit $?"; tail -n 3 gpurun_out/pair_model_tests.log
  MAED_B200_GEMM_2CTA=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_pair.json
fi
# (6) train step (configs[2]): MSE loss first, then the reference loss through the fused kernel
timeout 900 python bench.py --mode train --steps 5 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; echo "train bench exit $?"; cut -c1-400 gpurun_out/bench_train.json
timeout 900 python bench.py --mode train --loss fused --steps 5 --warmup 3 > gpurun_out/bench_train_fused.json 2> gpurun_out/bench_train_fused.err; echo "train bench (fused loss) exit $?"; cut -c1-400 gpurun_out/bench_train_fused.json
# (7) encoder='cnn': inference at the stage-1 shape, then its train step
timeout 600 python scripts/bench_cnn.py --steps 10 --warmup 3 > gpurun_out/bench_cnn.json 2> gpurun_out/bench_cnn.err; echo "cnn bench exit $?"; cut -c1-400 gpurun_out/bench_cnn.json
timeout 900 python bench.py --mode train --encoder cnn --steps 5 --warmup 3 > gpurun_out/bench_train_cnn.json 2> gpurun_out/bench_train_cnn.err; echo "cnn train bench exit $?"; cut -c1-400 gpurun_out/bench_train_cnn.json
# (8) configs[4] shape (bs=4, T=32) on one GPU
timeout 600 python scripts/bench_config5.py --steps 10 --warmup 3 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 bench exit $?"; cut -c1-300 gpurun_out/bench_config5.json
# (9) per-kernel timing of the training kernels at the bench shapes
timeout 600 python scripts/bench_train_kernels.py > gpurun_out/train_kernels.log 2>&1; echo "train kernels exit $?"; tail -n 12 gpurun_out/train_kernels.log
