// Launchers of the backward / optimiser kernels of the MAED training path (bwd_kernels.cu, attention_bwd.cu,
// gemm_splitk_sm100.cu).  Conventions as in kernels.h: token-major / NHWC activations, "planes" = fp16 hi/lo pair
// `plane` elements apart, fp32 for the residual stream and every reduction.  Activation gradients carry the loss
// scale of the caller; parameter gradients are un-scaled (`scale` arguments) when they are written.
//
// Reference for the arithmetic: autograd of the reference's forward ops (lib/models/resnetv2.py:35-93,
// vision_transformer.py:96-261, ktd.py:69-92); the training oracle is oracle/maed_oracle.py::maed_param_grads.
#pragma once
#include "common.h"

namespace maed {

// ---- layout helpers
// out[c, r] = in[r, c] for both planes: [R, C] (row stride ld_in) -> [C, R] (row stride ld_out >= R, multiple of 8)
int transpose_planes(const __half* in_hi, long long in_plane, int R, int C, int ld_in, __half* out_hi, long long out_plane,
                     int ld_out, cudaStream_t st);
// fp32 [R, C] rows (row stride ld_in) -> planes [R, C] dense (split_f32 with a row gather)
int split_rows_f32(const float* in, long long ld_in, int R, int C, __half* out_hi, long long out_plane, cudaStream_t st);
// out[c] (+)= scale * sum_r in[r, c];  scratch: kColsumChunks * C floats.  Deterministic two-stage reduction.
static constexpr int kColsumChunks = 256;
int colsum_f32(const float* in, long long ld, int R, int C, float scale, int accumulate, float* scratch, float* out,
               cudaStream_t st);
int colsum_planes(const __half* in_hi, long long plane, long long ld, int R, int C, float scale, int accumulate,
                  float* scratch, float* out, cudaStream_t st);
// a[i] += b[i]
int add_f32(float* a, const float* b, long long n, cudaStream_t st);
// out[i] = a[i] * s
int scale_f32(const float* a, float s, long long n, float* out, cudaStream_t st);

// ---- elementwise backward
// d[i] = hi[i] > 0 ? d[i] : 0            (ReLU mask taken from the saved post-ReLU activation planes)
int relu_mask_f32(float* d, const __half* act_hi, long long n, cudaStream_t st, int reverse = 0);
// d_pre = d_hid * gelu'(pre) -> planes    (exact erf GELU, reference vision_transformer.py:100)
int gelu_bwd(const float* d_hid, const float* pre, long long n, __half* out_hi, long long out_plane, cudaStream_t st);
// y = gelu(pre) -> planes                 (training forward keeps `pre`)
int gelu_fwd_planes(const float* pre, long long n, __half* out_hi, long long out_plane, cudaStream_t st);
// d_pre = d_y * (1 - y*y)                 (tanh)
int tanh_bwd(const float* d_y, const float* y, long long n, float* d_pre, cudaStream_t st);
// dropout: mask byte = keep; y = x * keep / (1-p).  Counter-based hash RNG (seed, element index).
int dropout_fwd(float* x, long long n, float p, unsigned long long seed, unsigned char* mask, cudaStream_t st);
int dropout_bwd(float* d, long long n, float p, const unsigned char* mask, cudaStream_t st);

// ---- LayerNorm backward (eps inside the sqrt): statistics are recomputed from the saved input x.
//   dx_out[row] = (dx_add ? dx_add[row] : 0) + LN'(dy[row]);   partial: [ln_bwd_partial_rows()] x [2C] (dgamma | dbeta)
int ln_bwd_partial_rows();
int layernorm_bwd(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* gamma, int rows,
                  int C, float eps, const float* dx_add, float* dx_out, long long dx_stride, float* partial,
                  cudaStream_t st);

// ---- GroupNorm(32) backward over NHWC fp32 conv outputs x with the forward's (sum, sumsq) statistics
//   dx = rstd * (dy*gamma - mean_g(dy*gamma) - xhat * mean_g(dy*gamma*xhat)) -> planes
//   dgb_partial: [n_img][2][C] per-image (sum dy*xhat | sum dy); red: [n_img][32][2] floats (scratch)
int groupnorm_bwd(const float* dy, const float* x, const double* stats, const float* gamma, int n_img, int HW, int C,
                  float eps, float* red, float* dgb_partial, __half* dx_hi, long long dx_plane, cudaStream_t st,
                  const float* relu_beta = nullptr, int order = 0);
//   relu_beta: the layer's beta when a ReLU followed the norm and dy is the gradient behind that ReLU (the mask
//   xhat*gamma+beta > 0 is recomputed on the fly); order: image order of the two passes (0 up/up, 1 down/up, 2 up/down)

// gn_cluster.cu: the same two operations with one thread-block cluster per image (single HBM read of x / dy, second pass from
// L2); MAED_ERR_UNSUPPORTED (nothing launched) for shapes they do not cover or with MAED_B200_GN_CLUSTER=0 -> multi-kernel path.
// The forward variant also writes the (sum, sumsq) statistics [n_img][32][2] the backward reads.
int groupnorm_fwd_cluster(const float* x, const float* gamma, const float* beta, int n_img, int HW, int C, float eps, int relu,
                          const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane, double* stats,
                          cudaStream_t st);
int groupnorm_bwd_cluster(const float* dy, const float* x, const double* stats, const float* gamma, const float* relu_beta,
                          int n_img, int HW, int C, float eps, float* dgb_partial, __half* dx_hi, long long dx_plane,
                          cudaStream_t st);

// ---- weight standardisation backward (reference resnetv2.py:86-89): g = dL/dW_hat in the packed layout
// [Cout][kh][kw][Cin] (row stride k_pad) -> dW OIHW = scale * ((g - mean g)/(std+eps) - w_hat * mean(g*w_hat)/std)
int wstd_bwd(const float* g, int k_pad, const float* w, int Cout, int Cin, int KH, int KW, float eps, float scale,
             float* dw, cudaStream_t st);
// plain permute for non-standardised convs / copies: dW[co][ci][kh][kw] = scale * g[co][(kh,kw),ci]
// ---- derived weights for the data-gradient GEMMs
// linear W [N, K] fp32 -> planes of W^T [K, N]
int split_f32_transposed(const float* w, int N, int K, __half* out_hi, long long plane, cudaStream_t st);
// conv weight OIHW (standardised when `standardize`) -> planes [Cin][(KH-1-kh, KW-1-kw), Cout]: the B operand of the
// stride-1 data-gradient convolution (flipped taps, in/out channels swapped)
int prep_conv_weight_dgrad(const float* w, int Cout, int Cin, int KH, int KW, int standardize, __half* out_hi,
                           long long plane, cudaStream_t st);

// ---- stem: GN + ReLU + MaxPool2dSame(3,2) keeping the arg-max tap (0..8) of every output element, and its backward
int gn_apply_maxpool_idx(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int H, int W,
                         int C, float eps, __half* out_hi, long long out_plane, unsigned char* idx, cudaStream_t st);
// d_y[n, h, w, c] (gradient w.r.t. the GN output, ReLU applied) from d_pool [n, OH, OW, c]
int maxpool_gn_relu_bwd(const float* d_pool, const unsigned char* idx, const float* x, const double* stats,
                        const float* gamma, const float* beta, int n_img, int H, int W, int C, float eps, float* d_y,
                        cudaStream_t st);

// ---- stride-2 helpers of the data gradients
// out planes [n, 2*OH(+odd), 2*OW, C] zero everywhere except out[n, 2*oh, 2*ow, :] = in[n, oh, ow, :]
int dilate2_planes(const __half* in_hi, long long in_plane, int n_img, int OH, int OW, int C, int H, int W, __half* out_hi,
                   long long out_plane, cudaStream_t st);
// d_in fp32 [n, H, W, C] = (add ? add : 0) with d_in[n, 2*oh, 2*ow, :] += src[n, oh, ow, :]
int scatter_stride2_f32(const float* src, int n_img, int OH, int OW, int C, int H, int W, const float* add, float* d_in,
                        cudaStream_t st);

// ---- STE pieces
// parallel-mode attentive addition backward (forward: kernels.h ts_blend)
//   d_logits [BT, 2C];  d_xs / d_xt [BT*ntok, C] = d_ao * alpha (the token-mean term is added by blend_bwd_pool)
int blend_bwd(const float* d_ao, const float* x_s, const float* x_t, const float* logits, int BT, int ntok, int C,
              float* d_logits, float* d_xs, float* d_xt, cudaStream_t st);
//   d_xs[bt, tok, :] += d_pool[bt, 0:C] / ntok;  d_xt[bt, tok, :] += d_pool[bt, C:2C] / ntok
int blend_bwd_pool(const float* d_pool, int BT, int ntok, int C, float* d_xs, float* d_xt, cudaStream_t st);
// sum over the tokens of each frame: out[bt, c] = sum_tok x[bt, tok, c]
int token_sum(const float* x, int BT, int ntok, int C, float* out, cudaStream_t st);

// ---- small fp32 GEMM for the tail (CUDA cores):  C[M,N] = alpha * op(A) * op(B) + beta * C
//   transA = 0: A is [M,K] (lda);  1: A is [K,M].   transB = 0: B is [K,N] (ldb);  1: B is [N,K].
int sgemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
              float beta, float* C, int ldc, cudaStream_t st);

// ---- KTD kinematic-tree backward (forward: kernels.h ktd_tree)
//   g_total [R,144] (gradient w.r.t. every joint output incl. descendants' contributions), d_base [R, ld] (cols 0..156)
int ktd_tree_bwd(const float* d_pose6d, const float* d_shape, const float* d_cam, const float* w_anc, int R, float* g_total,
                 float* d_base, int ld, cudaStream_t st);
//   d_w_anc[36*95] = scale * sum_r g_total[r, j] (x) pose6d[r, ancestors(j)]   (same packing as the forward's w_anc)
int ktd_anc_wgrad(const float* g_total, const float* pose6d, int R, float scale, float* d_w_anc, cudaStream_t st);

// ---- Adam (torch.optim.Adam semantics: L2 weight decay added to the gradient, bias-corrected moments)
int adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
              double weight_decay, int step, float grad_scale, cudaStream_t st);

// weight gradient of a stride-1 k x k conv without an im2col matrix (implicit operand: 5-D TMA boxes of the NHWC planes shifted by
// the tap): D[Cout, k*k*Cin] (+)= scale * dY^T im2col(x); dY [n_img*H*W, Cout], x [n_img, H, W, Cin]; Cin, Cout multiples of 64
int gemm_wgrad_conv(const __half* dY, long long dy_plane, const __half* X, long long x_plane, int n_img, int H, int W, int Cin,
                    int Cout, int KH, int KW, int pad, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                    cudaStream_t st);

// ---- attention backward (attention_bwd.cu); qkv planes as in the forward, d_out fp32 [BT*ntok, H*64],
// d_qkv fp32 [BT*ntok, 3*H*64]; `accumulate` adds to d_qkv instead of overwriting it
int attn_spatial_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int BT, int ntok, int heads, float scale,
                     int accumulate, float* d_qkv, cudaStream_t st);
// tcgen05 version (attention_bwd_sm100.cu): d_out given as fp16 hi/lo planes [BT*ntok, heads*64]; ntok <= 208
// lse / Dv (optional, together, both [rows, heads]): the forward's log2-domain row log-sum-exp (attn_spatial / attn_temporal `lse`
// output) and D = rowsum(dO o O) (attn_rowdot) — with them the score-orientation pass needs no reductions
int attn_spatial_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int BT, int ntok,
                        int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse = nullptr,
                        const float* Dv = nullptr);
// D[row, h] = sum_d dO[row, h*64 + d] * O[row, h*64 + d]; O as fp32 [rows, heads*64] or as fp16 hi/lo planes (exactly one non-null)
int attn_rowdot(const float* d_out, const float* o_f32, const __half* o_hi, long long o_plane, long long rows, int heads, float* D,
                cudaStream_t st);
// temporal attention backward on the same tcgen05 kernel (TMA-gathered {64, 128/T, T} tiles); T in {4, 8, 16, 32}
int attn_temporal_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int B, int T,
                         int ntok, int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse = nullptr,
                         const float* Dv = nullptr);
int attn_temporal_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int B, int T, int ntok, int heads,
                      float scale, int accumulate, float* d_qkv, cudaStream_t st);

// ---- split-K weight-gradient GEMM on the tensor cores (gemm_splitk_sm100.cu)
//   D[Mo, No] (+)= scale * sum_r A[Mo, r] * B[No, r]      A, B: planes, r contiguous (row strides lda / ldb)
//   slabs: splitk_slab_floats(Mo, No) floats of scratch
size_t splitk_slab_floats(int Mo, int No, int R);
int gemm_wgrad_splitk(const __half* A, long long a_plane, int lda, const __half* B, long long b_plane, int ldb, int Mo,
                      int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                      cudaStream_t st);
// the same contraction on row-major operands (no transposed copies): dY [R, Mo] (ld_dy), X [R, No_x] (ld_x), planes;
// D [Mo, No], No >= No_x a multiple of 32, columns No_x..No are written as zeros
int gemm_wgrad_rows(const __half* dY, long long dy_plane, int ld_dy, const __half* X, long long x_plane, int ld_x, int No_x,
                    int Mo, int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                    cudaStream_t st);

// generic attention backward over `seq` contiguous rows per batch element ('coupling' mode, vision_transformer.py:180-204);
// stats: batch * heads * seq * 3 floats of scratch
int attn_generic_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int batch, int seq, int heads, float scale,
                     int accumulate, float* d_qkv, float* stats, cudaStream_t st);

// ---- 'cnn' encoder training (cnn_kernels.cu): BatchNorm2d with batch statistics, MaxPool2d(3, 2, 1) with arg-max, pools
// SyncBatchNorm hook: when `fn` is set, the per-channel sums (2C + 1 doubles: sum, sum of products, row count) are placed in
// `buf` (device memory owned by the caller) and fn(user, n) must add the first n doubles up over the data-parallel ranks
// (e.g. torch.distributed.all_reduce on the current stream) before the statistics are used.
struct BnExchange { int (*fn)(void* user, int n_doubles); void* user; double* buf; int capacity; };
size_t bn_partial_doubles(long long M, int C);
size_t bn_scratch_doubles(long long M, int C);            // scratch (`partial`) of bn_train_stats / bn_bwd
// mean / rstd over the M rows of x [M, C] (biased variance, eps inside the sqrt); running buffers (may be nullptr) get the
// momentum update with the unbiased variance, like nn.BatchNorm2d in train()
int bn_train_stats(const float* x, long long M, int C, float eps, float momentum, double* partial, float* mean, float* rstd,
                   float* running_mean, float* running_var, const BnExchange* ex, cudaStream_t st);
int bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, long long M, int C,
             int relu, const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane, cudaStream_t st);
int bn_relu_mask(float* d, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, long long M,
                 int C, cudaStream_t st);
// dgamma / dbeta = scale * this rank's column sums (written), dx -> planes (means over all ranks when `ex` exchanges)
int bn_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, long long M, int C,
           float scale, double* partial, float* dgamma, float* dbeta, __half* dx_hi, long long dx_plane, const BnExchange* ex,
           cudaStream_t st);
// MaxPool2d(3, 2, 1) of relu(BatchNorm(x)) (mean == nullptr: of x itself) -> planes + the arg-max tap (0..8) per element
int maxpool3x3s2_idx(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int n_img, int H,
                     int W, int C, __half* out_hi, long long plane, unsigned char* idx, cudaStream_t st);
int maxpool3x3s2_bwd(const float* d_out, const unsigned char* idx, int n_img, int H, int W, int C, float* d_x, cudaStream_t st);
int avgpool_bwd(const float* d_feat, int BT, int P, int C, float* d_map, cudaStream_t st);
int add_cols_f32(float* dst, int ldd, const float* src, int lds, int R, int n, int accumulate, cudaStream_t st);
int wgrad_permute(const float* g, int k_pad, int Cout, int Cin, int KH, int KW, float scale, float* dw, cudaStream_t st);

}  // namespace maed
