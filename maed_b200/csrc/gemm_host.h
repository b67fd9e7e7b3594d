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
  // epilogue: out = act_post(act(acc + bias) + residual + residual planes)
  const float* bias = nullptr;
  const float* residual = nullptr;       // fp32 [M, ldc]
  const __half* res_hi = nullptr;        // residual as planes [M, ldc]: hi, and lo at + res_plane when res_plane != 0
  long long res_plane = 0;
  int act_post = 0;                      // ACT_NONE / ACT_RELU after the residuals
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

// Fused conv (GEMM) + GroupNorm(32) + optional shortcut + optional ReLU -> planes (gemm_gn_sm100.cu).
struct ConvGnArgs {
  const __half* A = nullptr; long long a_plane = 0;      // plain: [n_img*HW, K]; conv: NHWC [n_img, H, W, Cin]
  const __half* B = nullptr; long long b_plane = 0;      // [C, K]
  int n_img = 0, H_out = 0, W_out = 0, C = 0, K = 0, nsplit = 3;
  int conv = 0, Cin = 0, KH = 0, KW = 0, pad_h = 0, pad_w = 0;   // conv: stride-1 tap mode, H_out x W_out == input size
  const float* gamma = nullptr; const float* beta = nullptr; float eps = 1e-5f; int relu = 0;
  const __half* res = nullptr; long long res_plane = 0;
  __half* out = nullptr; long long out_plane = 0;
  long long* dbg = nullptr;                               // optional in-kernel timeline (see gemm_gn_sm100.cu)
};
// MAED_ERR_UNSUPPORTED (nothing launched) when the shape does not fit the fused kernel.
int conv_gn_fused(const ConvGnArgs& a, cudaStream_t st);
void conv_tile_shape(int H, int W, int* tile_h, int* tile_w);

}  // namespace maed
