// Launchers of the non-GEMM kernels of the MAED hot path (kernels.cu, attention.cu, decoder.cu).
// Layout conventions:
//   * activations are NHWC / token-major; "planes" = split-precision fp16 pair (hi at p, lo at p + plane);
//   * fp32 tensors: conv outputs before GroupNorm, the STE residual stream, everything in the tail.
#pragma once
#include "common.h"

namespace maed {

long long launch_count();
void count_launch(int n = 1);

// ---- precision helpers
int split_f32(const float* in, __half* out_hi, long long plane, long long n, cudaStream_t st);

// ---- weight preparation (derived caches of the fp32 parameters)
// StdConv2dSame.get_weight (reference resnetv2.py:86-89) + OIHW -> [Cout][kh][kw][Cin] + pad K to k_pad + split.
int prep_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int k_pad, int standardize, __half* out_hi,
                     long long plane, cudaStream_t st);
// nn.Linear weight [N,K] fp32 -> planes (same layout)
inline int prep_linear_weight(const float* w, long long n, __half* out_hi, long long plane, cudaStream_t st) {
  return split_f32(w, out_hi, plane, n, st);
}

// ---- gathers feeding explicit-im2col GEMMs
// stem: x fp32 NCHW [n,3,224,224] -> A planes [n*OH*OW, k_pad], k = (r*KW + s)*Cin + c, SAME pad (top/left = pad/2)
int im2col_stem(const float* x, int n_img, int Cin, int H, int W, int KH, int KW, int stride, int pad_t, int pad_l,
                int OH, int OW, int k_pad, __half* out_hi, long long plane, cudaStream_t st);
// NHWC planes -> A planes [n*OH*OW, KH*KW*C]   (stride-2 3x3 and 1x1 convs)
int im2col_nhwc(const __half* in_hi, long long in_plane, int n_img, int H, int W, int C, int KH, int KW, int stride,
                int pad_t, int pad_l, int OH, int OW, __half* out_hi, long long out_plane, cudaStream_t st);

// ---- stem conv 7x7/2 (3 -> 64) as a tcgen05 implicit GEMM with in-kernel im2col (stem_sm100.cu); also accumulates
// the GroupNorm statistics of its output into `stats` (pre-zeroed [n_img][32][2] doubles)
int stem_conv(const float* x, int n_img, const __half* w_hi, long long w_plane, int k_pad, int nsplit, float* out,
              double* stats, cudaStream_t st);

// ---- GroupNorm(32 groups, eps) over NHWC fp32 conv outputs (reference resnetv2.py:45-49)
// stats: double [n_img][32][2] = (sum, sumsq), must be zeroed by the caller (gn_stats accumulates).
// reverse (here and in gn_apply): walk the images downwards — the producer of x wrote upwards, its last images are still in L2
int gn_stats(const float* x, int n_img, int HW, int C, double* stats, cudaStream_t st, int reverse = 0);
// y = relu?( (x-mean)*rstd*gamma + beta (+ residual) ) -> planes
int gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int HW, int C,
             float eps, int relu, const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane,
             cudaStream_t st, int reverse = 0);
// stem: GN + ReLU + MaxPool2dSame(3, stride 2) fused (reference resnetv2.py:61-72)
int gn_apply_maxpool(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int H,
                     int W, int C, float eps, __half* out_hi, long long out_plane, cudaStream_t st);

// ---- 'cnn' encoder (torchvision ResNet-50; cnn_kernels.cu)
// eval-mode BatchNorm folded into the conv: w_out[co] = w[co] * s, bias_out[co] = beta - mean * s, s = gamma / sqrt(var + eps)
int fold_bn(const float* w, int Cout, long long E, const float* gamma, const float* beta, const float* mean, const float* var,
            float eps, float* w_out, float* bias_out, cudaStream_t st);
// nn.MaxPool2d(3, 2, 1) on an fp32 NHWC map -> fp32 NHWC and / or planes (either output may be nullptr)
int maxpool3x3s2(const float* x, int n_img, int H, int W, int C, float* out_f32, __half* out_hi, long long plane,
                 cudaStream_t st);

// ---- STE elementwise
// x[bt, 0] = cls + pos[0]; x[bt, 1+i] = tok[bt, i] + pos[1+i]; (+ temp[bt % T] when temp != nullptr)
int embed_assemble(const float* tok, const float* cls, const float* pos, const float* temp, int BT, int T, int ntok,
                   int C, float* x, cudaStream_t st);
// LayerNorm over the last dim (eps) of fp32 rows -> planes.  row_stride lets the tail normalise only row 0 of
// each frame (reference vision_transformer.py:406).
int layernorm_planes(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C,
                     float eps, __half* out_hi, long long out_plane, cudaStream_t st);
int layernorm_f32(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C,
                  float eps, float* out, cudaStream_t st);
// mean over the ntok tokens of each frame: in fp32 [BT, ntok, C] -> out[BT, out_ld] at column offset col0
int token_mean(const float* x, int BT, int ntok, int C, float* out, int out_ld, int col0, cudaStream_t st);
// mean over tokens of two fp32 tensors at once -> planes [BT, 2C] = [mean(a) | mean(b)] (deterministic two-stage
// reduction; `scratch` holds BT * kTokenChunks * 2C floats)
static constexpr int kTokenChunks = 8;
int token_mean2_planes(const float* a, const float* b, int BT, int ntok, int C, float* scratch, __half* out_hi,
                       long long out_plane, cudaStream_t st);
// parallel-mode attentive addition (reference vision_transformer.py:152-158):
//   logits [BT, 2C] (channel c uses entries 2c, 2c+1); out = x_t*softmax_1 + x_s*softmax_0 -> planes
int ts_blend(const float* x_s, const float* x_t, const float* logits, int BT, int ntok, int C, __half* out_hi,
             long long out_plane, cudaStream_t st);
// x[bt, i, :] += v[bt, :]  ('temporal' mode: attention output of the token-mean is broadcast over tokens)
int broadcast_add(float* x, const float* v, int BT, int ntok, int C, cudaStream_t st);

// ---- attention (attention.cu)
// qkv planes: [BT*ntok, 3*H*64] (q | k | v, head-major inside each).  Outputs fp32 or planes [BT*ntok, H*64].
int attn_spatial(const __half* qkv_hi, long long qkv_plane, int BT, int ntok, int heads, float scale, int nsplit,
                 float* out_f32, __half* out_hi, long long out_plane, cudaStream_t st, float* lse = nullptr);   // lse: optional [BT*ntok, heads] log2-domain row statistics
// tcgen05 version (attention_temporal_sm100.cu): T in {4, 8, 16, 32}, split precision; attn_temporal dispatches to it
bool attn_temporal_tc_supported(int T, long long qkv_plane);
int attn_temporal_tc(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale, float* out_f32,
                     __half* out_hi, long long out_plane, cudaStream_t st, float* lse = nullptr);
int attn_temporal(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale,
                  float* out_f32, __half* out_hi, long long out_plane, cudaStream_t st, float* lse = nullptr);
// generic CUDA-core fp32 attention over `seq` tokens addressed as row = base(b) + i*row_step (coupling mode,
// and the cross-check of the tcgen05 kernel in tests)
int attn_generic(const __half* qkv_hi, long long qkv_plane, int batch, int seq, int heads, float scale,
                 int tokens_per_frame, int frames_per_batch, float* out_f32, __half* out_hi, long long out_plane,
                 cudaStream_t st);

// ---- tail / decoders (decoder.cu): fp32 CUDA-core helpers (the KTD / pre_logits GEMMs themselves run as split-precision
// tcgen05 GEMMs from engine.cu; linear_f32 serves the iterative regressor and the tests)
// out[R,N] = act(x[R,K] @ W[N,K]^T + bias) (+ residual);  act: 0 none, 3 tanh
int linear_f32(const float* x, int ldx, const float* W, int ldw, const float* bias, int R, int N, int K, int act,
               const float* residual, int ldr, float* out, int ldo, cudaStream_t st);
// KTD kinematic-tree pass (reference ktd.py:81-86): pose6d[r, j] = base[r, j] + W_anc[j] . pose6d[r, ancestors(j)]
// `base` has row stride ld (>= 157): columns [0,144) joint bases, [144,154) shape, [154,157) cam (copied out).
int ktd_tree(const float* base, int ld, const float* w_anc, int R, float* pose6d, float* shape, float* cam,
             cudaStream_t st);
// rot6d -> rotmat -> angle-axis; theta = [cam, aa, shape]; kp_2d from kp_3d and cam
// (reference geometry.py:320-334,58-223; ktd.py:94-124; spin.py:113-157)
int decode_outputs(const float* pose6d, const float* shape, const float* cam, int R, const float* kp3d, int n_joints,
                   float* rotmat, float* theta, float* kp2d, cudaStream_t st);
// training-path geometry tail (decode_bwd.cu): backward of pose6d -> rotmat / angle-axis, and the keypoint projection
int decode_pose_backward(const float* pose6d, int R, const float* d_rotmat, const float* d_aa, int ld_aa, float* d_pose6d,
                         cudaStream_t st);
int project_keypoints_forward(const float* kp3d, const float* cam, int R, int J, float* kp2d, cudaStream_t st);
int project_keypoints_backward(const float* kp3d, const float* cam, int R, int J, const float* d_kp2d, float* d_cam, float* d_kp3d,
                               cudaStream_t st);
int concat_cols(const float* a, int ca, const float* b, int cb, const float* c, int cc, const float* d, int cd, int R,
                float* out, cudaStream_t st);

}  // namespace maed
