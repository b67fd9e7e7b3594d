// extern "C" surface of libmaed_b200.so (declared in include/maed_b200.h).
#include "../../include/maed_b200.h"

#include <atomic>

#include "common.h"
#include "gemm_host.h"
#include "engine.h"
#include "bwd_kernels.h"
#include "kernels.h"
#include "loss.h"
#include "smpl.h"
#include "train.h"

using namespace maed;

extern "C" {

const char* maed_last_error(void) { return last_error(); }
int maed_version(void) { return 1; }
long long maed_launch_count(void) { return launch_count(); }

int maed_op_gemm(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb, int M, int N,
                 int K, int nsplit, const float* bias, const float* residual, int act, int out_mode, void* out,
                 long long out_plane, int ldc, int force_block_n, void* stream) {
  GemmArgs g;
  g.A = (const __half*)A; g.a_plane = a_plane; g.lda = lda;
  g.B = (const __half*)B; g.b_plane = b_plane; g.ldb = ldb;
  g.M = M; g.N = N; g.K = K; g.nsplit = nsplit;
  g.bias = bias; g.residual = residual; g.act = act; g.out_mode = out_mode; g.out = out; g.out_plane = out_plane;
  g.ldc = ldc; g.force_block_n = force_block_n;
  return launch_gemm(g, (cudaStream_t)stream);
}

int maed_op_gemm_bottleneck(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb, int M, int N,
                            int K, int nsplit, const float* bias, const void* res_hi, long long res_plane, int act_post,
                            int out_mode, void* out, long long out_plane, int ldc, void* stream) {
  GemmArgs g;
  g.A = (const __half*)A; g.a_plane = a_plane; g.lda = lda;
  g.B = (const __half*)B; g.b_plane = b_plane; g.ldb = ldb;
  g.M = M; g.N = N; g.K = K; g.nsplit = nsplit;
  g.bias = bias; g.res_hi = (const __half*)res_hi; g.res_plane = res_plane; g.act_post = act_post;
  g.out_mode = out_mode; g.out = out; g.out_plane = out_plane; g.ldc = ldc;
  return launch_gemm(g, (cudaStream_t)stream);
}

int maed_op_conv_gemm(const void* A, long long a_plane, const void* B, long long b_plane, int n_img, int H, int W,
                      int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, int nsplit, int out_mode, void* out,
                      long long out_plane, int force_block_n, void* stream) {
  GemmArgs g;
  g.A = (const __half*)A; g.a_plane = a_plane;
  g.B = (const __half*)B; g.b_plane = b_plane;
  g.M = n_img * H * W; g.N = Cout; g.K = KH * KW * Cin; g.nsplit = nsplit;
  g.out_mode = out_mode; g.out = out; g.out_plane = out_plane; g.ldc = Cout;
  g.conv = 1; g.n_img = n_img; g.H = H; g.W = W; g.Cin = Cin; g.KH = KH; g.KW = KW; g.pad_h = pad_h; g.pad_w = pad_w;
  g.force_block_n = force_block_n;
  return launch_gemm(g, (cudaStream_t)stream);
}

int maed_op_fold_bn(const float* w, int Cout, long long E, const float* gamma, const float* beta, const float* mean,
                    const float* var, float eps, float* w_out, float* bias_out, void* stream) {
  return fold_bn(w, Cout, E, gamma, beta, mean, var, eps, w_out, bias_out, (cudaStream_t)stream);
}
int maed_op_maxpool3x3s2(const float* x, int n_img, int H, int W, int C, float* out_f32, void* out_hi, long long plane,
                         void* stream) {
  return maxpool3x3s2(x, n_img, H, W, C, out_f32, (__half*)out_hi, plane, (cudaStream_t)stream);
}

int maed_op_split_f32(const float* in, void* out_hi, long long plane, long long n, void* stream) {
  return split_f32(in, (__half*)out_hi, plane, n, (cudaStream_t)stream);
}

int maed_op_prep_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int k_pad, int standardize, void* out_hi,
                             long long plane, void* stream) {
  return prep_conv_weight(w, Cout, Cin, KH, KW, k_pad, standardize, (__half*)out_hi, plane, (cudaStream_t)stream);
}
int maed_op_im2col_stem(const float* x, int n_img, int Cin, int H, int W, int KH, int KW, int stride, int pad_t, int pad_l,
                        int OH, int OW, int k_pad, void* out_hi, long long plane, void* stream) {
  return im2col_stem(x, n_img, Cin, H, W, KH, KW, stride, pad_t, pad_l, OH, OW, k_pad, (__half*)out_hi, plane,
                     (cudaStream_t)stream);
}
int maed_op_im2col_nhwc(const void* in_hi, long long in_plane, int n_img, int H, int W, int C, int KH, int KW, int stride,
                        int pad_t, int pad_l, int OH, int OW, void* out_hi, long long out_plane, void* stream) {
  return im2col_nhwc((const __half*)in_hi, in_plane, n_img, H, W, C, KH, KW, stride, pad_t, pad_l, OH, OW, (__half*)out_hi,
                     out_plane, (cudaStream_t)stream);
}
int maed_op_conv_gn(const void* A, long long a_plane, const void* W, long long w_plane, int n_img, int H, int Wd, int Cin, int C,
                    int KH, int KW, int nsplit, const float* gamma, const float* beta, float eps, int relu, const void* res_hi,
                    long long res_plane, void* out_hi, long long out_plane, long long* dbg, void* stream) {
  ConvGnArgs f;
  f.A = (const __half*)A; f.a_plane = a_plane; f.B = (const __half*)W; f.b_plane = w_plane;
  f.n_img = n_img; f.H_out = H; f.W_out = Wd; f.C = C; f.K = KH * KW * Cin; f.nsplit = nsplit;
  f.conv = (KH * KW > 1) ? 1 : 0; f.Cin = Cin; f.KH = KH; f.KW = KW; f.pad_h = (KH - 1) / 2; f.pad_w = (KW - 1) / 2;
  f.gamma = gamma; f.beta = beta; f.eps = eps; f.relu = relu; f.res = (const __half*)res_hi; f.res_plane = res_plane;
  f.out = (__half*)out_hi; f.out_plane = out_plane; f.dbg = dbg;
  return conv_gn_fused(f, (cudaStream_t)stream);
}
int maed_op_stem_conv(const float* x, int n_img, const void* w_hi, long long w_plane, int k_pad, int nsplit, float* out,
                      double* stats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)n_img * 64 * sizeof(double), st));
  return stem_conv(x, n_img, (const __half*)w_hi, w_plane, k_pad, nsplit, out, stats, st);
}
int maed_op_groupnorm(const float* x, int n_img, int HW, int C, const float* gamma, const float* beta, float eps, int relu,
                      const void* res_hi, long long res_plane, void* out_hi, long long out_plane, double* stats_scratch,
                      void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, (size_t)n_img * 64 * sizeof(double), st));
  MAED_PROPAGATE(gn_stats(x, n_img, HW, C, stats_scratch, st));
  return gn_apply(x, stats_scratch, gamma, beta, n_img, HW, C, eps, relu, (const __half*)res_hi, res_plane, (__half*)out_hi,
                  out_plane, st);
}
int maed_op_groupnorm_train(const float* x, int n_img, int HW, int C, const float* gamma, const float* beta, float eps, int relu,
                            const void* res_hi, long long res_plane, void* out_hi, long long out_plane, double* stats,
                            void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
#ifndef MAED_EMU
  const int rc = groupnorm_fwd_cluster(x, gamma, beta, n_img, HW, C, eps, relu, (const __half*)res_hi, res_plane, (__half*)out_hi,
                                       out_plane, stats, st);
  if (rc != MAED_ERR_UNSUPPORTED) return rc;
#endif
  return maed_op_groupnorm(x, n_img, HW, C, gamma, beta, eps, relu, res_hi, res_plane, out_hi, out_plane, stats, stream);
}
int maed_op_groupnorm_maxpool(const float* x, int n_img, int H, int W, int C, const float* gamma, const float* beta, float eps,
                              void* out_hi, long long out_plane, double* stats_scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, (size_t)n_img * 64 * sizeof(double), st));
  MAED_PROPAGATE(gn_stats(x, n_img, H * W, C, stats_scratch, st));
  return gn_apply_maxpool(x, stats_scratch, gamma, beta, n_img, H, W, C, eps, (__half*)out_hi, out_plane, st);
}
int maed_op_layernorm(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C, float eps,
                      void* out_hi, long long out_plane, void* stream) {
  return layernorm_planes(x, row_stride, gamma, beta, rows, C, eps, (__half*)out_hi, out_plane, (cudaStream_t)stream);
}
int maed_op_attention(int kind, const void* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale,
                      int nsplit, float* out_f32, void* out_hi, long long out_plane, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const __half* q = (const __half*)qkv_hi;
  if (kind == 0) return attn_spatial(q, qkv_plane, B * T, ntok, heads, scale, nsplit, out_f32, (__half*)out_hi, out_plane, st);
  if (kind == 1) return attn_temporal(q, nsplit == 3 ? qkv_plane : 0, B, T, ntok, heads, scale, out_f32, (__half*)out_hi, out_plane, st);
  if (kind == 2) return attn_generic(q, nsplit == 3 ? qkv_plane : 0, B, T * ntok, heads, scale, ntok, T, out_f32, (__half*)out_hi, out_plane, st);
  set_error("maed_op_attention: unknown kind %d", kind);
  return MAED_ERR_ARG;
}
int maed_op_linear_f32(const float* x, int ldx, const float* W, int ldw, const float* bias, int R, int N, int K, int act,
                       const float* residual, int ldr, float* out, int ldo, void* stream) {
  return linear_f32(x, ldx, W, ldw, bias, R, N, K, act, residual, ldr, out, ldo, (cudaStream_t)stream);
}
int maed_op_decode_outputs(const float* pose6d, const float* shape, const float* cam, int R, const float* kp3d, int n_joints,
                           float* rotmat, float* theta, float* kp2d, void* stream) {
  return decode_outputs(pose6d, shape, cam, R, kp3d, n_joints, rotmat, theta, kp2d, (cudaStream_t)stream);
}

// ---- engine
static_assert(sizeof(maed_config) == sizeof(EngineConfig), "maed_config must mirror EngineConfig");
static_assert(sizeof(maed_outputs) == sizeof(EngineOutputs), "maed_outputs must mirror EngineOutputs");
static_assert(MAED_TAP_COUNT == TAP_COUNT, "tap count mismatch");
int maed_engine_create(const maed_config* cfg, maed_engine** out) {
  return engine_create(reinterpret_cast<const EngineConfig*>(cfg), reinterpret_cast<Engine**>(out));
}
void maed_engine_destroy(maed_engine* e) { engine_destroy(reinterpret_cast<Engine*>(e)); }
int maed_engine_num_params(const maed_engine* e) { return engine_num_params(reinterpret_cast<const Engine*>(e)); }
const char* maed_engine_param_name(const maed_engine* e, int i) { return engine_param_name(reinterpret_cast<const Engine*>(e), i); }
long long maed_engine_param_numel(const maed_engine* e, int i) { return engine_param_numel(reinterpret_cast<const Engine*>(e), i); }
size_t maed_engine_packed_bytes(const maed_engine* e) { return engine_packed_bytes(reinterpret_cast<const Engine*>(e)); }
size_t maed_engine_workspace_bytes(const maed_engine* e, int n_frames) {
  return engine_workspace_bytes(reinterpret_cast<const Engine*>(e), n_frames);
}
int maed_engine_pack(const maed_engine* e, const void* const* params, void* packed, void* stream) {
  return engine_pack(reinterpret_cast<const Engine*>(e), params, packed, (cudaStream_t)stream);
}
int maed_engine_forward(const maed_engine* e, const void* const* params, const void* packed, const float* x, int N, int T,
                        void* workspace, size_t workspace_bytes, const maed_outputs* outs, float* const* taps, void* stream) {
  return engine_forward(reinterpret_cast<const Engine*>(e), params, packed, x, N, T, workspace, workspace_bytes,
                        reinterpret_cast<const EngineOutputs*>(outs), taps, (cudaStream_t)stream);
}

// ---- training path
static_assert(sizeof(maed_train_outputs) == sizeof(TrainOutputs), "maed_train_outputs must mirror TrainOutputs");
int maed_train_set_exchange(maed_engine* e, maed_exchange_fn fn, void* user, double* buffer, int capacity_doubles) {
  return train_set_exchange(reinterpret_cast<Engine*>(e), fn, user, buffer, capacity_doubles);
}
int maed_train_set_progress(maed_engine* e, maed_progress_fn fn, void* user) {
  return train_set_progress(reinterpret_cast<Engine*>(e), fn, user);
}
size_t maed_train_pack_bytes(const maed_engine* e) { return train_pack_bytes(reinterpret_cast<const Engine*>(e)); }
size_t maed_train_workspace_bytes(const maed_engine* e, int n_frames) {
  return train_workspace_bytes(reinterpret_cast<const Engine*>(e), n_frames);
}
int maed_train_pack(const maed_engine* e, const void* const* params, void* tpack, void* stream) {
  return train_pack(reinterpret_cast<const Engine*>(e), params, tpack, (cudaStream_t)stream);
}
int maed_train_forward(const maed_engine* e, const void* const* params, const void* packed, const float* x, int N, int T,
                       void* workspace, size_t workspace_bytes, float dropout_p, unsigned long long seed,
                       const maed_train_outputs* outs, void* stream) {
  return train_forward(reinterpret_cast<const Engine*>(e), params, packed, x, N, T, workspace, workspace_bytes, dropout_p, seed,
                       reinterpret_cast<const TrainOutputs*>(outs), (cudaStream_t)stream);
}
int maed_train_backward(const maed_engine* e, const void* const* params, const void* packed, const void* tpack,
                        const float* x, int N, int T, void* workspace, size_t workspace_bytes, const float* d_pose6d,
                        const float* d_shape, const float* d_cam, float loss_scale, float dropout_p, float* const* grads,
                        void* stream) {
  return train_backward(reinterpret_cast<const Engine*>(e), params, packed, tpack, x, N, T, workspace, workspace_bytes, d_pose6d,
                        d_shape, d_cam, loss_scale, dropout_p, grads, (cudaStream_t)stream);
}
int maed_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int step, float grad_scale, void* stream) {
  return adam_step(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
}

// ---- per-op backward entry points (unit tests)
int maed_bwd_transpose_planes(const void* in_hi, long long in_plane, int R, int C, int ld_in, void* out_hi, long long out_plane,
                              int ld_out, void* stream) {
  return transpose_planes((const __half*)in_hi, in_plane, R, C, ld_in, (__half*)out_hi, out_plane, ld_out, (cudaStream_t)stream);
}
int maed_bwd_colsum_chunks(void) { return kColsumChunks; }
int maed_bwd_colsum(const float* in, long long ld, int R, int C, float scale, int accumulate, float* scratch, float* out,
                    void* stream) {
  return colsum_f32(in, ld, R, C, scale, accumulate, scratch, out, (cudaStream_t)stream);
}
int maed_bwd_layernorm_partial_rows(void) { return ln_bwd_partial_rows(); }
int maed_bwd_layernorm(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* gamma, int rows,
                       int C, float eps, const float* dx_add, float* dx_out, long long dx_stride, float* partial,
                       float* scratch, float* dgamma, float* dbeta, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(layernorm_bwd(dy, dy_stride, x, x_stride, gamma, rows, C, eps, dx_add, dx_out, dx_stride, partial, st));
  const int pr = ln_bwd_partial_rows();
  MAED_PROPAGATE(colsum_f32(partial, 2 * C, pr, C, 1.f, 0, scratch, dgamma, st));
  return colsum_f32(partial + C, 2 * C, pr, C, 1.f, 0, scratch, dbeta, st);
}
int maed_bwd_groupnorm(const float* dy, const float* x, int n_img, int HW, int C, const float* gamma, float eps,
                       double* stats_scratch, float* red, float* dgb_partial, void* dx_hi, long long dx_plane,
                       const float* relu_beta, int order, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, (size_t)n_img * 64 * sizeof(double), st));
  MAED_PROPAGATE(gn_stats(x, n_img, HW, C, stats_scratch, st, order == 1));
  return groupnorm_bwd(dy, x, stats_scratch, gamma, n_img, HW, C, eps, red, dgb_partial, (__half*)dx_hi, dx_plane, st, relu_beta,
                       order);
}
size_t maed_bwd_batchnorm_scratch_doubles(long long M, int C) { return bn_scratch_doubles(M, C); }
int maed_bwd_batchnorm(const float* x, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       float* running_mean, float* running_var, int relu, const void* res_hi, long long res_plane, void* y_hi,
                       long long y_plane, float* mean, float* rstd, const float* dy, float scale, float* dgamma, float* dbeta,
                       void* dx_hi, long long dx_plane, double* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(bn_train_stats(x, M, C, eps, momentum, scratch, mean, rstd, running_mean, running_var, nullptr, st));
  MAED_PROPAGATE(bn_apply(x, mean, rstd, gamma, beta, M, C, relu, (const __half*)res_hi, res_plane, (__half*)y_hi, y_plane, st));
  if (!dy) return MAED_OK;
  return bn_bwd(dy, x, mean, rstd, gamma, M, C, scale, scratch, dgamma, dbeta, (__half*)dx_hi, dx_plane, nullptr, st);
}
int maed_bwd_maxpool3x3s2(const float* x, int n_img, int H, int W, int C, void* out_hi, long long out_plane, unsigned char* idx,
                          const float* d_out, float* d_x, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(maxpool3x3s2_idx(x, nullptr, nullptr, nullptr, nullptr, n_img, H, W, C, (__half*)out_hi, out_plane, idx, st));
  if (!d_out) return MAED_OK;
  return maxpool3x3s2_bwd(d_out, idx, n_img, H, W, C, d_x, st);
}
int maed_bwd_wstd(const float* g, int k_pad, const float* w, int Cout, int Cin, int KH, int KW, float eps, float scale, float* dw,
                  void* stream) {
  return wstd_bwd(g, k_pad, w, Cout, Cin, KH, KW, eps, scale, dw, (cudaStream_t)stream);
}
int maed_bwd_gelu(const float* d_hid, const float* pre, long long n, void* out_hi, long long out_plane, void* stream) {
  return gelu_bwd(d_hid, pre, n, (__half*)out_hi, out_plane, (cudaStream_t)stream);
}
int maed_bwd_relu_mask(float* d, const void* act_hi, long long n, void* stream) {
  return relu_mask_f32(d, (const __half*)act_hi, n, (cudaStream_t)stream);
}
int maed_bwd_maxpool(const float* x, int n_img, int H, int W, int C, const float* gamma, const float* beta, float eps,
                     double* stats_scratch, void* out_hi, long long out_plane, unsigned char* idx, const float* d_pool,
                     float* d_y, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_CUDA_CHECK(cudaMemsetAsync(stats_scratch, 0, (size_t)n_img * 64 * sizeof(double), st));
  MAED_PROPAGATE(gn_stats(x, n_img, H * W, C, stats_scratch, st));
  MAED_PROPAGATE(gn_apply_maxpool_idx(x, stats_scratch, gamma, beta, n_img, H, W, C, eps, (__half*)out_hi, out_plane, idx, st));
  return maxpool_gn_relu_bwd(d_pool, idx, x, stats_scratch, gamma, beta, n_img, H, W, C, eps, d_y, st);
}
int maed_bwd_dilate2(const void* in_hi, long long in_plane, int n_img, int OH, int OW, int C, int H, int W, void* out_hi,
                     long long out_plane, void* stream) {
  return dilate2_planes((const __half*)in_hi, in_plane, n_img, OH, OW, C, H, W, (__half*)out_hi, out_plane, (cudaStream_t)stream);
}
int maed_bwd_scatter_stride2(const float* src, int n_img, int OH, int OW, int C, int H, int W, const float* add, float* d_in,
                             void* stream) {
  return scatter_stride2_f32(src, n_img, OH, OW, C, H, W, add, d_in, (cudaStream_t)stream);
}
int maed_bwd_blend(const float* d_ao, const float* x_s, const float* x_t, const float* logits, const float* d_pool, int BT,
                   int ntok, int C, float* d_logits, float* d_xs, float* d_xt, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(blend_bwd(d_ao, x_s, x_t, logits, BT, ntok, C, d_logits, d_xs, d_xt, st));
  if (d_pool) return blend_bwd_pool(d_pool, BT, ntok, C, d_xs, d_xt, st);
  return MAED_OK;
}
int maed_bwd_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
                   float beta, float* C, int ldc, void* stream) {
  return sgemm_f32(transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, (cudaStream_t)stream);
}
int maed_bwd_ktd_tree(const float* d_pose6d, const float* d_shape, const float* d_cam, const float* w_anc, const float* pose6d,
                      int R, float scale, float* g_total, float* d_base, int ld, float* d_w_anc, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(ktd_tree_bwd(d_pose6d, d_shape, d_cam, w_anc, R, g_total, d_base, ld, st));
  return ktd_anc_wgrad(g_total, pose6d, R, scale, d_w_anc, st);
}
int maed_bwd_attention(int kind, const void* qkv_hi, long long qkv_plane, const float* d_out, int B, int T, int ntok, int heads,
                       float scale, int accumulate, float* d_qkv, float* scratch, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == 0) return attn_spatial_bwd((const __half*)qkv_hi, qkv_plane, d_out, B * T, ntok, heads, scale, accumulate, d_qkv, st);
  if (kind == 1) return attn_temporal_bwd((const __half*)qkv_hi, qkv_plane, d_out, B, T, ntok, heads, scale, accumulate, d_qkv, st);
  if (kind == 2) {   // 'coupling': joint attention over the T * ntok tokens of a clip; scratch = softmax statistics
    MAED_CHECK_ARG(scratch, "maed_bwd_attention(kind 2): scratch of B * heads * T * ntok * 3 floats required");
    return attn_generic_bwd((const __half*)qkv_hi, qkv_plane, d_out, B, T * ntok, heads, scale, accumulate, d_qkv, scratch, st);
  }
  if (kind == 3) {   // spatial on the tensor cores; scratch = 2 * B*T*ntok * heads*64 halfs (hi / lo planes of d_out)
    MAED_CHECK_ARG(scratch, "maed_bwd_attention(kind 3): scratch of B*T*ntok * heads*64 floats required");
    const long long n = (long long)B * T * ntok * heads * 64;
    MAED_PROPAGATE(split_f32(d_out, (__half*)scratch, n, n, st));
    return attn_spatial_bwd_tc((const __half*)qkv_hi, qkv_plane, (const __half*)scratch, n, B * T, ntok, heads, scale, accumulate,
                               d_qkv, st);
  }
  if (kind == 4) {   // temporal on the tensor cores; scratch as for kind 3
    MAED_CHECK_ARG(scratch, "maed_bwd_attention(kind 4): scratch of B*T*ntok * heads*64 floats required");
    const long long n = (long long)B * T * ntok * heads * 64;
    MAED_PROPAGATE(split_f32(d_out, (__half*)scratch, n, n, st));
    return attn_temporal_bwd_tc((const __half*)qkv_hi, qkv_plane, (const __half*)scratch, n, B, T, ntok, heads, scale, accumulate,
                                d_qkv, st);
  }
  if (kind == 5 || kind == 6) {
    // the train step's arrangement: forward kernel with row statistics, D = rowsum(dO o O), then the one-pass backward.
    // scratch (floats): [n] planes of d_out | [n] forward output | [rows*heads] lse | [rows*heads] D,  n = rows * heads * 64
    MAED_CHECK_ARG(scratch, "maed_bwd_attention(kind 5/6): scratch of 2 * rows*heads*64 + 2 * rows*heads floats required");
    const long long rows = (long long)B * T * ntok, n = rows * heads * 64;
    float* o = scratch + n;
    float* lse = o + n;
    float* Dv = lse + rows * heads;
    if (kind == 5) MAED_PROPAGATE(attn_spatial((const __half*)qkv_hi, qkv_plane, B * T, ntok, heads, scale, 3, o, nullptr, 0, st, lse));
    else MAED_PROPAGATE(attn_temporal_tc((const __half*)qkv_hi, qkv_plane, B, T, ntok, heads, scale, o, nullptr, 0, st, lse));
    MAED_PROPAGATE(attn_rowdot(d_out, o, nullptr, 0, rows, heads, Dv, st));
    MAED_PROPAGATE(split_f32(d_out, (__half*)scratch, n, n, st));
    if (kind == 5)
      return attn_spatial_bwd_tc((const __half*)qkv_hi, qkv_plane, (const __half*)scratch, n, B * T, ntok, heads, scale, accumulate,
                                 d_qkv, st, lse, Dv);
    return attn_temporal_bwd_tc((const __half*)qkv_hi, qkv_plane, (const __half*)scratch, n, B, T, ntok, heads, scale, accumulate,
                                d_qkv, st, lse, Dv);
  }
  set_error("maed_bwd_attention: unknown kind %d", kind);
  return MAED_ERR_ARG;
}
size_t maed_bwd_wgrad_slab_floats(int Mo, int No, int R) { return splitk_slab_floats(Mo, No, R); }
int maed_bwd_wgrad_splitk(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb, int Mo, int No,
                          int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream) {
  return gemm_wgrad_splitk((const __half*)A, a_plane, lda, (const __half*)B, b_plane, ldb, Mo, No, R, nsplit, scale, accumulate,
                           slabs, D, ldd, (cudaStream_t)stream);
}
int maed_bwd_wgrad_rows(const void* dY, long long dy_plane, int ld_dy, const void* X, long long x_plane, int ld_x, int No_x, int Mo,
                        int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream) {
  return gemm_wgrad_rows((const __half*)dY, dy_plane, ld_dy, (const __half*)X, x_plane, ld_x, No_x, Mo, No, R, nsplit, scale,
                         accumulate, slabs, D, ldd, (cudaStream_t)stream);
}
int maed_bwd_wgrad_conv(const void* dY, long long dy_plane, const void* X, long long x_plane, int n_img, int H, int W, int Cin,
                        int Cout, int KH, int KW, int pad, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream) {
  return gemm_wgrad_conv((const __half*)dY, dy_plane, (const __half*)X, x_plane, n_img, H, W, Cin, Cout, KH, KW, pad, 3, scale,
                         accumulate, slabs, D, ldd, (cudaStream_t)stream);
}
int maed_bwd_split_transposed(const float* w, int N, int K, void* out_hi, long long plane, void* stream) {
  return split_f32_transposed(w, N, K, (__half*)out_hi, plane, (cudaStream_t)stream);
}
int maed_bwd_prep_conv_weight_dgrad(const float* w, int Cout, int Cin, int KH, int KW, int standardize, void* out_hi,
                                    long long plane, void* stream) {
  return prep_conv_weight_dgrad(w, Cout, Cin, KH, KW, standardize, (__half*)out_hi, plane, (cudaStream_t)stream);
}
int maed_bwd_dropout(float* x, long long n, float p, unsigned long long seed, unsigned char* mask, float* d, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MAED_PROPAGATE(dropout_fwd(x, n, p, seed, mask, st));
  if (d) return dropout_bwd(d, n, p, mask, st);
  return MAED_OK;
}

// ---- SMPL
static_assert(sizeof(maed_smpl_assets) == sizeof(SmplAssets), "maed_smpl_assets must mirror SmplAssets");
size_t maed_smpl_scratch_bytes(int n_frames) { return smpl_scratch_bytes(n_frames); }
int maed_smpl_forward(const maed_smpl_assets* assets, const float* betas, const float* rotmat, int R, const float* J_regressor,
                      int n_reg, float* verts, float* joints, void* scratch, size_t scratch_bytes, void* stream) {
  return smpl_forward(reinterpret_cast<const SmplAssets*>(assets), betas, rotmat, R, J_regressor, n_reg, verts, joints, scratch,
                      scratch_bytes, (cudaStream_t)stream);
}
size_t maed_smpl_backward_scratch_bytes(int n_frames) { return smpl_backward_scratch_bytes(n_frames); }
int maed_smpl_backward(const maed_smpl_assets* assets, const float* betas, const float* rotmat, int R, const float* J_regressor,
                       int n_reg, const float* d_verts, const float* d_joints, float* d_betas, float* d_rotmat, void* scratch,
                       size_t scratch_bytes, void* stream) {
  return smpl_backward(reinterpret_cast<const SmplAssets*>(assets), betas, rotmat, R, J_regressor, n_reg, d_verts, d_joints, d_betas,
                       d_rotmat, scratch, scratch_bytes, (cudaStream_t)stream);
}

// ---- geometry tail of the training path
int maed_decode_pose_backward(const float* pose6d, int R, const float* d_rotmat, const float* d_aa, int ld_aa, float* d_pose6d,
                              void* stream) {
  return decode_pose_backward(pose6d, R, d_rotmat, d_aa, ld_aa, d_pose6d, (cudaStream_t)stream);
}
int maed_project_keypoints(const float* kp3d, const float* cam, int R, int J, float* kp2d, const float* d_kp2d, float* d_cam,
                           float* d_kp3d, void* stream) {
  if (!d_kp2d) return project_keypoints_forward(kp3d, cam, R, J, kp2d, (cudaStream_t)stream);
  return project_keypoints_backward(kp3d, cam, R, J, d_kp2d, d_cam, d_kp3d, (cudaStream_t)stream);
}

// ---- fused loss
static_assert(sizeof(maed_loss_weights) == sizeof(LossWeights), "maed_loss_weights must mirror LossWeights");
size_t maed_loss_scratch_bytes(int M2, int M3) { return loss_scratch_bytes(M2, M3); }
int maed_loss_forward_backward(const float* pred_kp2d, const float* gt_kp2d, int M2, int J2, const float* pred_kp3d,
                               const float* gt_kp3d, int M3, int J3, const float* pred_theta, const float* gt_theta,
                               const unsigned char* valid, int T, const maed_loss_weights* w, float* losses, float* d_kp2d,
                               float* d_kp3d, float* d_theta, void* scratch, size_t scratch_bytes, void* stream) {
  return loss_forward_backward(pred_kp2d, gt_kp2d, M2, J2, pred_kp3d, gt_kp3d, M3, J3, pred_theta, gt_theta, valid, T,
                               reinterpret_cast<const LossWeights*>(w), losses, d_kp2d, d_kp3d, d_theta, scratch, scratch_bytes,
                               (cudaStream_t)stream);
}

}  // extern "C"
