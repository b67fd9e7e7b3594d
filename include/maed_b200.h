/* libmaed_b200.so — C ABI of the B200-native MAED hot path.
 *
 * Boundary being replaced: the reference has NO native code; its hot path is the Python call
 *   preds = model(inp)                         (reference lib/core/trainer.py:253, lib/core/evaluate.py:78)
 * on `lib.models.MAED` (lib/models/maed.py:52-66).  `maed_b200.models.MAED` keeps that Python surface and
 * forwards to the entry points below through ctypes (see INTEGRATION.md for the binding).
 *
 * Conventions: every function returns 0 on success, non-zero on failure (maed_last_error() gives the
 * message).  All pointers are DEVICE pointers owned by the caller (PyTorch); nothing is allocated on the
 * hot path; `stream` is a cudaStream_t passed as void*.  No torch types appear in any signature.
 */
#ifndef MAED_B200_H
#define MAED_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* maed_last_error(void);
int maed_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long maed_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Per-op entry points (unit-test / microbenchmark granularity).
 * "planes": an activation or weight matrix in split precision is two fp16 matrices, hi = rn(x) and
 * lo = rn(x - hi), `plane` elements apart.  nsplit = 3 uses both (3 MMAs per K step), nsplit = 1 only hi.
 * ---------------------------------------------------------------------------------------------- */

/* out = act(A[M,K] * B[N,K]^T + bias) + residual   — tcgen05 GEMM (replaces every nn.Linear / 1x1 conv
 * matmul: reference vision_transformer.py:105-111,139-177; resnetv2.py:91-93).
 * act: 0 none, 1 exact GELU, 2 ReLU.  out_mode: 0 fp32, 1 fp16, 2 fp16 hi+lo planes. */
int maed_op_gemm(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb,
                 int M, int N, int K, int nsplit, const float* bias, const float* residual, int act,
                 int out_mode, void* out, long long out_plane, int ldc, int force_block_n, void* stream);

/* out = act_post(A * B^T + bias + residual), the residual given as fp16 planes [M, ldc] (hi at res_hi, lo res_plane elements
 * further; res_plane = 0: hi only); act_post: 0 none, 2 ReLU — the tail of a ResNet bottleneck, conv3 + folded BatchNorm +
 * identity + ReLU in one launch (torchvision models/resnet.py Bottleneck.forward; reference lib/models/maed.py:36). */
int maed_op_gemm_bottleneck(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb, int M, int N,
                            int K, int nsplit, const float* bias, const void* res_hi, long long res_plane, int act_post,
                            int out_mode, void* out, long long out_plane, int ldc, void* stream);

/* Implicit-GEMM stride-1 KHxKW convolution over an NHWC activation [n_img,H,W,Cin] (x planes) with
 * weights [Cout, KH*KW*Cin] (x planes); zero padding (pad_h, pad_w) on the top/left, SAME-style on the
 * bottom/right (reference resnetv2.py:54-59,91-93).  Output [n_img*H*W, Cout]. */
int maed_op_conv_gemm(const void* A, long long a_plane, const void* B, long long b_plane, int n_img, int H,
                      int W, int Cin, int Cout, int KH, int KW, int pad_h, int pad_w, int nsplit, int out_mode,
                      void* out, long long out_plane, int force_block_n, void* stream);

/* fp32 -> fp16 hi/lo planes (elementwise) */
int maed_op_split_f32(const float* in, void* out_hi, long long plane, long long n, void* stream);

/* Conv weight preparation: per-output-channel standardisation (w-mean)/(std_biased+1e-5) (reference
 * resnetv2.py:86-89) when `standardize`, OIHW -> [Cout][kh][kw][Cin], zero-pad K to k_pad, split. */
int maed_op_prep_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int k_pad, int standardize,
                             void* out_hi, long long plane, void* stream);
/* explicit im2col gathers (stem from fp32 NCHW; strided convs from NHWC planes) */
int maed_op_im2col_stem(const float* x, int n_img, int Cin, int H, int W, int KH, int KW, int stride, int pad_t,
                        int pad_l, int OH, int OW, int k_pad, void* out_hi, long long plane, void* stream);
int maed_op_im2col_nhwc(const void* in_hi, long long in_plane, int n_img, int H, int W, int C, int KH, int KW,
                        int stride, int pad_t, int pad_l, int OH, int OW, void* out_hi, long long out_plane,
                        void* stream);
/* Fused StdConv (1x1 plain GEMM when KH = KW = 1, else stride-1 KHxKW implicit GEMM) -> GroupNorm(32, eps) -> (+ residual
 * planes) -> optional ReLU -> planes [n_img*H*W, C] (reference resnetv2.py:189-204).  Returns 3 (unsupported shape, nothing
 * launched) when the image does not fit the TMEM of one cluster.  dbg: NULL or items*8 int64 clock stamps. */
int maed_op_conv_gn(const void* A, long long a_plane, const void* W, long long w_plane, int n_img, int H, int Wd, int Cin,
                    int C, int KH, int KW, int nsplit, const float* gamma, const float* beta, float eps, int relu,
                    const void* res_hi, long long res_plane, void* out_hi, long long out_plane, long long* dbg,
                    void* stream);
/* stem: StdConv 7x7/2 SAME 3->64 on fp32 NCHW frames [n,3,224,224] -> fp32 NHWC [n*112*112, 64]; tcgen05 implicit
 * GEMM with the im2col tile built in shared memory; accumulates the GroupNorm (sum, sumsq) of the output into
 * stats[n][32][2] (zeroed by this call).  w: prep_conv_weight(k_pad) planes (reference resnetv2.py:245-274). */
int maed_op_stem_conv(const float* x, int n_img, const void* w_hi, long long w_plane, int k_pad, int nsplit,
                      float* out, double* stats, void* stream);
/* GroupNorm(32, eps) over an NHWC fp32 map (+ optional residual planes, ReLU) -> planes
 * (reference resnetv2.py:35-49).  `stats_scratch`: n_img*64 doubles. */
int maed_op_groupnorm(const float* x, int n_img, int HW, int C, const float* gamma, const float* beta, float eps,
                      int relu, const void* res_hi, long long res_plane, void* out_hi, long long out_plane,
                      double* stats_scratch, void* stream);
/* stem: GN + ReLU + MaxPool2dSame(3,2) (reference resnetv2.py:61-72,245-274) */
/* GroupNorm of the training forward: same result as maed_op_groupnorm, one thread-block cluster per image when the shape allows
   (x leaves HBM once); `stats` [n_img][32][2] doubles (sum, sum of squares) is an OUTPUT kept on the tape for maed_bwd_groupnorm */
int maed_op_groupnorm_train(const float* x, int n_img, int HW, int C, const float* gamma, const float* beta, float eps, int relu,
                            const void* res_hi, long long res_plane, void* out_hi, long long out_plane, double* stats,
                            void* stream);
int maed_op_groupnorm_maxpool(const float* x, int n_img, int H, int W, int C, const float* gamma, const float* beta,
                              float eps, void* out_hi, long long out_plane, double* stats_scratch, void* stream);
/* 'cnn' encoder (torchvision ResNet-50, reference lib/models/maed.py:35-37), inference:
 * eval-mode BatchNorm2d folded into the preceding bias-free conv: w_out[co,:] = w[co,:] * s, bias_out[co] = beta - mean * s
 * with s = gamma / sqrt(var + eps); E = Cin*KH*KW weights per output channel. */
int maed_op_fold_bn(const float* w, int Cout, long long E, const float* gamma, const float* beta, const float* mean,
                    const float* var, float eps, float* w_out, float* bias_out, void* stream);
/* nn.MaxPool2d(3, stride 2, padding 1) on an fp32 NHWC map [n,H,W,C] -> fp32 NHWC and / or planes (either may be NULL) */
int maed_op_maxpool3x3s2(const float* x, int n_img, int H, int W, int C, float* out_f32, void* out_hi, long long plane,
                         void* stream);
/* LayerNorm(eps) rows of fp32 -> planes (reference vision_transformer.py:258-261) */
int maed_op_layernorm(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C,
                      float eps, void* out_hi, long long out_plane, void* stream);
/* attention over qkv planes [BT*ntok, 3*heads*64]; outputs fp32 and/or planes [BT*ntok, heads*64]
 * (reference vision_transformer.py:206-228,180-204).  kind: 0 spatial (tcgen05), 1 temporal, 2 generic. */
int maed_op_attention(int kind, const void* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads,
                      float scale, int nsplit, float* out_f32, void* out_hi, long long out_plane, void* stream);
/* small fp32 linear (tail): out = act(x W^T + b) + residual; act 0 none / 3 tanh */
int maed_op_linear_f32(const float* x, int ldx, const float* W, int ldw, const float* bias, int R, int N, int K,
                       int act, const float* residual, int ldr, float* out, int ldo, void* stream);
/* rot6d -> rotmat, angle-axis, theta, kp_2d (reference geometry.py:320-334,58-223; spin.py:113-157) */
int maed_op_decode_outputs(const float* pose6d, const float* shape, const float* cam, int R, const float* kp3d,
                           int n_joints, float* rotmat, float* theta, float* kp2d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Whole-model engine: MAED(encoder='ste' | 'cnn').forward (reference lib/models/maed.py:52-66).
 * ---------------------------------------------------------------------------------------------- */
typedef struct maed_engine maed_engine;
typedef struct maed_config {
  int num_blocks;   /* MODEL.ENCODER.NUM_BLOCKS (6) */
  int num_heads;    /* MODEL.ENCODER.NUM_HEADS (12; head_dim must be 64) */
  int mode;         /* st_mode: 0 vanilla, 1 parallel, 2 series, 3 coupling, 4 temporal */
  int decoder;      /* 0 ktd, 1 iterative */
  int hidden_dim;   /* MODEL.DECODER.HIDDEN_DIM (1024) */
  int nsplit;       /* 3 split-fp16 operands (parity mode), 1 plain fp16 (fast mode) */
  int temp_frames;  /* rows of encoder.temp_embed (16 in the reference; 32 for the T=32 extension) */
  int encoder;      /* MODEL.ENCODER.BACKBONE: 0 'ste' (hybrid ResNetV2 + STE blocks, 768 features), 1 'cnn' (torchvision
                       ResNet-50 conv+BN+ReLU, 2048 features, reference lib/models/maed.py:35-37; inference only,
                       BatchNorm uses its running statistics; num_blocks / num_heads / mode are ignored) */
} maed_config;
typedef struct maed_outputs {
  float* feat;       /* [N*T, 768] ('cnn': [N*T, 2048])  encoder feature (MAED.extract_feature) */
  float* pose6d;     /* [N*T, 144] */
  float* shape;      /* [N*T, 10] */
  float* cam;        /* [N*T, 3] */
  float* rotmat;     /* [N*T, 24, 3, 3] */
  float* theta;      /* [N*T, 85] = cam | angle-axis pose | shape */
  float* kp2d;       /* [N*T, n_joints, 2] */
  const float* kp3d; /* [N*T, n_joints, 3] joints to project, or NULL (zeros) */
  int n_joints;
} maed_outputs;

int maed_engine_create(const maed_config* cfg, maed_engine** out);
void maed_engine_destroy(maed_engine* e);
/* parameter table: the engine wants one fp32 device pointer per reference state_dict key, in this order */
int maed_engine_num_params(const maed_engine* e);
const char* maed_engine_param_name(const maed_engine* e, int i);
long long maed_engine_param_numel(const maed_engine* e, int i);
size_t maed_engine_packed_bytes(const maed_engine* e);
size_t maed_engine_workspace_bytes(const maed_engine* e, int n_frames);
/* derive the packed tensor-core weights (standardised, K-major fp16 planes) from the fp32 parameters */
int maed_engine_pack(const maed_engine* e, const void* const* params, void* packed, void* stream);
/* x: fp32 [N, T, 3, 224, 224].  taps: NULL or MAED_TAP_COUNT pointers (NULL entries skipped). */
#define MAED_TAP_COUNT 13
int maed_engine_forward(const maed_engine* e, const void* const* params, const void* packed, const float* x, int N,
                        int T, void* workspace, size_t workspace_bytes, const maed_outputs* outs,
                        float* const* taps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training path (maed_b200/csrc/train.cu): forward with a saved-activation tape + backward to every parameter.
 * Replaces `loss.backward()` through lib/models/maed.py:52-66 (reference lib/core/trainer.py:238-255); the boundary
 * is the decoder output pose6d / shape / cam (the geometry tail behind it stays under autograd).
 * Supported: encoder 'ste' with all five st_modes, encoder 'cnn' (BatchNorm on batch statistics, optional
 * SyncBatchNorm exchange), both decoders (KTD, iterative regressor), split precision.
 * ---------------------------------------------------------------------------------------------- */
typedef struct maed_train_outputs {
  float* feat;    /* [N*T, 768] or NULL */
  float* pose6d;  /* [N*T, 144] */
  float* shape;   /* [N*T, 10] */
  float* cam;     /* [N*T, 3] */
} maed_train_outputs;
/* SyncBatchNorm hook for encoder = 'cnn' (reference train.py:95 converts every BatchNorm to SyncBatchNorm under DDP): when set,
 * every BatchNorm of maed_train_forward / maed_train_backward places its per-channel sums (2C + 1 doubles: sum, sum of products,
 * row count) in `buffer` (device memory owned by the caller, >= 2 * 2048 + 1 doubles) and calls fn(user, n): the callback must
 * add the first n doubles up over the data-parallel ranks in stream order (torch.distributed.all_reduce on the current stream:
 * NCCL over NVLink on the GPU box, gloo in the CPU tests) and return 0.  fn == NULL: statistics of this rank's batch only. */
typedef int (*maed_exchange_fn)(void* user, int n_doubles);
int maed_train_set_exchange(maed_engine* e, maed_exchange_fn fn, void* user, double* buffer, int capacity_doubles);
/* Backward progress hook (overlap of the data-parallel gradient exchange with the backward; the role of DDP's bucketed
 * all-reduce, reference train.py:113): during maed_train_backward fn(user, first, end) is called on the host, in stream order,
 * when the gradients of the engine parameters with table index in [first, end) are final — encoder = 'ste': after STE block k
 * (k = num_blocks - 1 ... 0) the range [first parameter of block k, first parameter of block k + 1 or table end for the last
 * block: final norm, pre_logits and the decoder are behind the blocks in the table], finally [0, first parameter of block 0)
 * (embeddings, backbone, patch projection).  fn must return 0.  fn == NULL removes the hook. */
typedef int (*maed_progress_fn)(void* user, int first_param, int end_param);
int maed_train_set_progress(maed_engine* e, maed_progress_fn fn, void* user);
size_t maed_train_pack_bytes(const maed_engine* e);
size_t maed_train_workspace_bytes(const maed_engine* e, int n_frames);
/* derived weights of the data-gradient GEMMs; redo after every parameter update (like maed_engine_pack) */
int maed_train_pack(const maed_engine* e, const void* const* params, void* tpack, void* stream);
/* dropout_p = 0: the reference in eval() mode (parity configuration); > 0: KTD dropout (ktd.py:54-56) */
int maed_train_forward(const maed_engine* e, const void* const* params, const void* packed, const float* x, int N, int T,
                       void* workspace, size_t workspace_bytes, float dropout_p, unsigned long long seed,
                       const maed_train_outputs* outs, void* stream);
/* grads[i]: fp32 [maed_engine_param_numel(e, i)], overwritten with dL/dparam_i.  workspace: the one the matching
 * maed_train_forward call filled.  loss_scale multiplies the activation gradients internally (fp16 range). */
int maed_train_backward(const maed_engine* e, const void* const* params, const void* packed, const void* tpack,
                        const float* x, int N, int T, void* workspace, size_t workspace_bytes, const float* d_pose6d,
                        const float* d_shape, const float* d_cam, float loss_scale, float dropout_p, float* const* grads,
                        void* stream);
/* torch.optim.Adam step on flat fp32 buffers (reference lib/utils/utils.py:127-131); g is multiplied by grad_scale.  The
 * hyper-parameters are doubles: like torch, the scalar coefficients (1 - beta, lr / (1 - beta1^step), ...) are formed in double
 * on the host and rounded to fp32 once. */
int maed_adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int step, float grad_scale, void* stream);

/* per-op entry points of the backward kernels (unit tests; semantics in maed_b200/csrc/bwd_kernels.h) */
int maed_bwd_transpose_planes(const void* in_hi, long long in_plane, int R, int C, int ld_in, void* out_hi, long long out_plane,
                              int ld_out, void* stream);
/* column sums (bias / affine gradients): two-stage, fixed order.  `scratch` of maed_bwd_colsum / maed_bwd_layernorm holds
 * maed_bwd_colsum_chunks() * C floats (2 * C for the LayerNorm entry: gamma and beta partials) */
int maed_bwd_colsum_chunks(void);
int maed_bwd_colsum(const float* in, long long ld, int R, int C, float scale, int accumulate, float* scratch, float* out,
                    void* stream);
int maed_bwd_layernorm(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* gamma, int rows,
                       int C, float eps, const float* dx_add, float* dx_out, long long dx_stride, float* partial,
                       float* scratch, float* dgamma, float* dbeta, void* stream);
int maed_bwd_layernorm_partial_rows(void);
/* relu_beta: NULL, or the layer's beta when a ReLU followed the norm and dy is the gradient behind that ReLU (the mask is
   recomputed from x); order: image order of the two passes over dy / x (0 up/up, 1 down/up, 2 up/down — L2 reuse only) */
int maed_bwd_groupnorm(const float* dy, const float* x, int n_img, int HW, int C, const float* gamma, float eps,
                       double* stats_scratch, float* red, float* dgb_partial, void* dx_hi, long long dx_plane,
                       const float* relu_beta, int order, void* stream);
int maed_bwd_wstd(const float* g, int k_pad, const float* w, int Cout, int Cin, int KH, int KW, float eps, float scale, float* dw,
                  void* stream);
int maed_bwd_gelu(const float* d_hid, const float* pre, long long n, void* out_hi, long long out_plane, void* stream);
int maed_bwd_relu_mask(float* d, const void* act_hi, long long n, void* stream);
int maed_bwd_maxpool(const float* x, int n_img, int H, int W, int C, const float* gamma, const float* beta, float eps,
                     double* stats_scratch, void* out_hi, long long out_plane, unsigned char* idx, const float* d_pool,
                     float* d_y, void* stream);
int maed_bwd_dilate2(const void* in_hi, long long in_plane, int n_img, int OH, int OW, int C, int H, int W, void* out_hi,
                     long long out_plane, void* stream);
int maed_bwd_scatter_stride2(const float* src, int n_img, int OH, int OW, int C, int H, int W, const float* add, float* d_in,
                             void* stream);
int maed_bwd_blend(const float* d_ao, const float* x_s, const float* x_t, const float* logits, const float* d_pool, int BT,
                   int ntok, int C, float* d_logits, float* d_xs, float* d_xt, void* stream);
int maed_bwd_sgemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
                   float beta, float* C, int ldc, void* stream);
int maed_bwd_ktd_tree(const float* d_pose6d, const float* d_shape, const float* d_cam, const float* w_anc, const float* pose6d,
                      int R, float scale, float* g_total, float* d_base, int ld, float* d_w_anc, void* stream);
/* kind: 0 spatial (per frame; CUDA-core cross-check), 1 temporal (per token across the T frames), 2 coupling (all T * ntok tokens
 * of a clip; needs scratch of B * heads * T * ntok * 3 floats), 3 spatial on the tensor cores (tcgen05, ntok <= 208: what the train
 * step runs; scratch of B*T*ntok * heads*64 floats holds the fp16 hi/lo planes of d_out), 4 temporal on the tensor cores (same
 * kernel on TMA-gathered {64, 128/T, T} tiles, T in {4, 8, 16, 32}; scratch as for 3), 5 / 6 = 3 / 4 the way the train step runs
 * them: the forward kernel leaves its row log-sum-exp, D = rowsum(dO o O) is precomputed, and the backward needs one element-wise
 * pass per score tile (scratch: 2 * rows*heads*64 + 2 * rows*heads floats); scratch NULL otherwise */
int maed_bwd_attention(int kind, const void* qkv_hi, long long qkv_plane, const float* d_out, int B, int T, int ntok, int heads,
                       float scale, int accumulate, float* d_qkv, float* scratch, void* stream);
size_t maed_bwd_wgrad_slab_floats(int Mo, int No, int R);
int maed_bwd_wgrad_splitk(const void* A, long long a_plane, int lda, const void* B, long long b_plane, int ldb, int Mo, int No,
                          int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream);
/* the same on row-major operands: D[Mo, No] (+)= scale * dY[R, Mo]^T X[R, No_x] (MN-major tcgen05 operands, no transposes) */
int maed_bwd_wgrad_rows(const void* dY, long long dy_plane, int ld_dy, const void* X, long long x_plane, int ld_x, int No_x, int Mo,
                        int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream);
/* weight gradient of a stride-1 KH x KW convolution with an implicit im2col operand (5-D TMA boxes of the NHWC planes shifted by
 * the tap): D[Cout, KH*KW*Cin] (+)= scale * dY[n_img*H*W, Cout]^T im2col(x[n_img, H, W, Cin]); layout [Cout][kh][kw][Cin];
 * slabs: maed_bwd_wgrad_slab_floats(Cout, KH*KW*Cin, n_img*H*W) floats */
int maed_bwd_wgrad_conv(const void* dY, long long dy_plane, const void* X, long long x_plane, int n_img, int H, int W, int Cin,
                        int Cout, int KH, int KW, int pad, float scale, int accumulate, float* slabs, float* D, int ldd, void* stream);
int maed_bwd_split_transposed(const float* w, int N, int K, void* out_hi, long long plane, void* stream);
/* nn.BatchNorm2d in train() mode over x [M, C] (rows = N*H*W): batch statistics (mean / rstd out; running buffers updated with
 * `momentum` and the unbiased variance when given), y = relu?(xhat * gamma + beta (+ residual planes)) -> planes; when dy is
 * given also the backward: dgamma / dbeta (scaled) and dx planes.  scratch: maed_bwd_batchnorm_scratch_doubles(M, C) doubles. */
size_t maed_bwd_batchnorm_scratch_doubles(long long M, int C);
int maed_bwd_batchnorm(const float* x, long long M, int C, const float* gamma, const float* beta, float eps, float momentum,
                       float* running_mean, float* running_var, int relu, const void* res_hi, long long res_plane, void* y_hi,
                       long long y_plane, float* mean, float* rstd, const float* dy, float scale, float* dgamma, float* dbeta,
                       void* dx_hi, long long dx_plane, double* scratch, void* stream);
/* nn.MaxPool2d(3, 2, 1) on fp32 NHWC -> planes + arg-max tap (0..8) per element; when d_out [n,OH,OW,C] is given also the
 * backward d_x [n,H,W,C] (gather over the windows of each pixel). */
int maed_bwd_maxpool3x3s2(const float* x, int n_img, int H, int W, int C, void* out_hi, long long out_plane, unsigned char* idx,
                          const float* d_out, float* d_x, void* stream);
int maed_bwd_prep_conv_weight_dgrad(const float* w, int Cout, int Cin, int KH, int KW, int standardize, void* out_hi,
                                    long long plane, void* stream);
int maed_bwd_dropout(float* x, long long n, float p, unsigned long long seed, unsigned char* mask, float* d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SMPL body model forward (maed_b200/csrc/smpl.cu): replaces the `self.smpl(...)` call of the decoders
 * (reference lib/models/ktd.py:100-114 -> lib/models/smpl.py:84-106 -> smplx.lbs.lbs, smplx not vendored: the published
 * algorithm is restated; parity with smplx itself is unpinned).  fp32 / int32 device pointers.
 * ---------------------------------------------------------------------------------------------- */
typedef struct maed_smpl_assets {
  const float* v_template;        /* [6890, 3] */
  const float* shapedirs;         /* [6890*3, 10] */
  const float* posedirs;          /* [207, 6890*3] */
  const float* J_template;        /* [24, 3]    = J_regressor @ v_template */
  const float* J_shapedirs;       /* [24*3, 10] = J_regressor @ shapedirs */
  const float* lbs_weights;       /* [6890, 24] */
  const float* J_regressor_extra; /* [9, 6890] */
  const int* parents;             /* [24] */
  const int* extra_vertex_ids;    /* [21] */
  const int* joint_map;           /* [49] */
} maed_smpl_assets;
size_t maed_smpl_scratch_bytes(int n_frames);
/* betas [R,10], rotmat [R,24,3,3] -> verts [R,6890,3]; joints [R,49,3], or [R,n_reg,3] = J_regressor @ verts when given */
int maed_smpl_forward(const maed_smpl_assets* assets, const float* betas, const float* rotmat, int R, const float* J_regressor,
                      int n_reg, float* verts, float* joints, void* scratch, size_t scratch_bytes, void* stream);
/* backward of the same: d_verts [R,6890,3] (or NULL) and d_joints [R,49 | n_reg,3] -> d_betas [R,10], d_rotmat [R,24,3,3]
 * (what the reference's keypoint losses back-propagate through the body model, lib/core/loss.py:178-192) */
size_t maed_smpl_backward_scratch_bytes(int n_frames);
int maed_smpl_backward(const maed_smpl_assets* assets, const float* betas, const float* rotmat, int R, const float* J_regressor,
                       int n_reg, const float* d_verts, const float* d_joints, float* d_betas, float* d_rotmat, void* scratch,
                       size_t scratch_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Geometry tail of the training path (maed_b200/csrc/decode_bwd.cu): gradients through rot6d -> rotation matrix -> angle-axis
 * (reference lib/utils/geometry.py:320-334,58-223) and the weak-perspective keypoint projection (lib/models/spin.py:113-157);
 * the forward of the first is maed_op_decode_outputs.
 * ---------------------------------------------------------------------------------------------- */
/* d_pose6d [R,144] = J_rotmat^T d_rotmat [R,24,9] (or NULL) + J_aa^T d_aa (72 values per frame at row stride ld_aa, e.g. theta + 3
 * with ld_aa = 85; or NULL) */
int maed_decode_pose_backward(const float* pose6d, int R, const float* d_rotmat, const float* d_aa, int ld_aa, float* d_pose6d,
                              void* stream);
/* d_kp2d == NULL: kp2d [R,J,2] = project(kp3d [R,J,3] or NULL (zeros), cam [R,3]);  otherwise the backward: d_cam [R,3] and, when
 * d_kp3d != NULL, d_kp3d [R,J,3] */
int maed_project_keypoints(const float* kp3d, const float* cam, int R, int J, float* kp2d, const float* d_kp2d, float* d_cam,
                           float* d_kp3d, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused training loss (maed_b200/csrc/loss.cu): replaces `self.criterion(preds, ...)` (reference lib/core/trainer.py:254 ->
 * lib/core/loss.py:159-210 LossVideo / :214-283 LossImage) — every term of the reference and its gradient with respect to
 * the predictions in three launches, no host synchronisation.
 * ---------------------------------------------------------------------------------------------- */
typedef struct maed_loss_weights {
  float kp2d;   /* e_loss_weight         (LOSS.KP_2D_W) */
  float kp3d;   /* e_3d_loss_weight      (LOSS.KP_3D_W) */
  float pose;   /* e_pose_loss_weight    (LOSS.POSE_W);  pose and shape terms exist only when both weights are > 0 */
  float shape;  /* e_shape_loss_weight   (LOSS.SHAPE_W) */
  float norm;   /* e_smpl_norm_loss */
  float accl;   /* e_smpl_accl_loss      (LOSS.ACCL_W; video only) */
} maed_loss_weights;
size_t maed_loss_scratch_bytes(int M2, int M3);
/* pred_kp2d [M2,J2,2], gt_kp2d [M2,J2,3] = (x, y, conf) or NULL; pred_kp3d [M3,J3,3], gt_kp3d [M3,J3,4] = (x, y, z, conf) or
 * NULL (49-joint layout: pelvis = mean of joints 27, 28); pred_theta / gt_theta [M3,85]; valid [M3] bytes (w_smpl) or NULL =
 * all frames; T = frames per clip (M3 % T == 0).  losses[8] = kp2d, kp3d, shape, pose, norm, accl (weighted, the order of the
 * reference's loss_dict), their total, n_valid.  d_kp2d / d_kp3d / d_theta = d total / d prediction (fully written). */
int maed_loss_forward_backward(const float* pred_kp2d, const float* gt_kp2d, int M2, int J2, const float* pred_kp3d,
                               const float* gt_kp3d, int M3, int J3, const float* pred_theta, const float* gt_theta,
                               const unsigned char* valid, int T, const maed_loss_weights* w, float* losses, float* d_kp2d,
                               float* d_kp3d, float* d_theta, void* scratch, size_t scratch_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAED_B200_H */
