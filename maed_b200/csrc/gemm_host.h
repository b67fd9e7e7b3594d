// Host-side descriptor of one tcgen05 GEMM launch (see gemm_sm100.cuh for the kernel).
#pragma once
#include "common.h"

namespace maed {

struct GemmArgs {
  // operands: fp16, K-major.  nsplit==3: two planes (hi, lo) `*_plane` elements apart; nsplit==1: one plane.
  const __half* A = nullptr; long long a_plane = 0; int lda = 0;   // plain mode: [M, lda]
  const __half* B = nullptr; long long b_plane = 0; int ldb = 0;   // [N, ldb]
  int M = 0, N = 0, K = 0;
  int nsplit = 3;
  // epilogue: out = act(acc + bias) + residual
  const float* bias = nullptr;
  const float* residual = nullptr;       // fp32 [M, ldc]
  int act = 0;                           // ACT_*
  int out_mode = 0;                      // OUT_*
  void* out = nullptr;
  long long out_plane = 0;
  int ldc = 0;
  // implicit-GEMM conv mode (stride 1): A is NHWC [n_img, H, W, Cin] (x planes), K = KH*KW*Cin
  int conv = 0, n_img = 0, H = 0, W = 0, Cin = 0, KH = 0, KW = 0, pad_h = 0, pad_w = 0;
  int force_block_n = 0;
};

int launch_gemm(const GemmArgs& g, cudaStream_t st);
void conv_tile_shape(int H, int W, int* tile_h, int* tile_w);

}  // namespace maed
