// Memory-bound kernels of the 'cnn' encoder (torchvision ResNet-50, reference lib/models/maed.py:35-37): BatchNorm folding
// at pack time and the 3x3/2 max-pool behind the stem.  HBM-bound: 16-byte accesses along the channel dimension, grids capped
// at a few waves of the SMs.
#include "bwd_kernels.h"
#include "device_utils.cuh"

namespace maed {

using bw::grid_for;
using bw::store_split4;

// ------------------------------------------------------------------------------------ BatchNorm folding (eval mode)
// y = (conv(x, w) - mean) / sqrt(var + eps) * gamma + beta  ==  conv(x, w * s) + (beta - mean * s),  s = gamma / sqrt(var + eps)
// One block per output channel; E = Cin*KH*KW weights per channel.
__global__ void fold_bn_kernel(const float* __restrict__ w, long long E, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ w_out, float* __restrict__ bias_out) {
  const int co = blockIdx.x;
  const float s = gamma[co] / sqrtf(var[co] + eps);
  const float* wc = w + (long long)co * E;
  float* oc = w_out + (long long)co * E;
  for (long long i = threadIdx.x; i < E; i += blockDim.x) oc[i] = wc[i] * s;
  if (threadIdx.x == 0) bias_out[co] = beta[co] - mean[co] * s;
}
int fold_bn(const float* w, int Cout, long long E, const float* gamma, const float* beta, const float* mean, const float* var,
            float eps, float* w_out, float* bias_out, cudaStream_t st) {
  MAED_CHECK_ARG(w && gamma && beta && mean && var && w_out && bias_out, "fold_bn: null argument");
  MAED_CHECK_ARG(Cout >= 1 && E >= 1, "fold_bn: bad shape Cout=%d E=%lld", Cout, E);
  fold_bn_kernel<<<Cout, 256, 0, st>>>(w, E, gamma, beta, mean, var, eps, w_out, bias_out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ----------------------------------------------------------------------------- nn.MaxPool2d(3, stride 2, padding 1)
// fp32 NHWC in -> fp32 NHWC out (the identity of the first bottleneck) and planes (its tensor-core operand).  Padding
// positions never win (PyTorch pads with -inf): taps outside the image are skipped.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, int H, int W, int C, int OH, int OW, long long total4,
                                    float* __restrict__ out_f32, __half* __restrict__ out_hi, long long plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long m = i / c4n;
    const int ow = (int)(m % OW);
    const int oh = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - 1;
        if (iw < 0 || iw >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + ((n * H + ih) * W + iw) * C + c);
        best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y); best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
      }
    }
    const long long o = m * C + c;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = best;
    if (out_hi) store_split4(out_hi + o, plane, best);
  }
}
int maxpool3x3s2(const float* x, int n_img, int H, int W, int C, float* out_f32, __half* out_hi, long long plane,
                 cudaStream_t st) {
  MAED_CHECK_ARG(x && (out_f32 || out_hi), "maxpool3x3s2: null argument");
  MAED_CHECK_ARG(C % 4 == 0 && H >= 1 && W >= 1 && n_img >= 1, "maxpool3x3s2: bad shape n=%d H=%d W=%d C=%d", n_img, H, W, C);
  MAED_CHECK_ARG(!out_hi || plane % 4 == 0, "maxpool3x3s2: plane stride must be a multiple of 4");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total4 = (long long)n_img * OH * OW * (C / 4);
  maxpool3x3s2_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x, H, W, C, OH, OW, total4, out_f32, out_hi, plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// =====================================================================================================================
// Training path of the 'cnn' encoder: BatchNorm with batch statistics (forward + backward), the max-pool with its arg-max,
// average-pool backward, weight-gradient permute.  Deterministic (fixed-order two-stage reductions, no atomics).
// =====================================================================================================================
static inline int bn_chunks(long long M) {
  long long c = (M + 255) / 256;
  return (int)(c < 1 ? 1 : (c > 1024 ? 1024 : c));
}
size_t bn_partial_doubles(long long M, int C) { return (size_t)bn_chunks(M) * 2 * C; }
size_t bn_scratch_doubles(long long M, int C) { return bn_partial_doubles(M, C) + 2 * (size_t)C + 1; }

// partial[chunk][0][c] = sum_r a, partial[chunk][1][c] = sum_r a * b over the rows of the chunk;
// mode 0 (forward statistics): a = b = x;  mode 1 (backward): a = dy, b = xhat = (x - mean) * rstd.  Thread = channel.
__global__ void bn_partial_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                  const float* __restrict__ rstd, long long M, int C, int mode, double* __restrict__ partial) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const long long per = (M + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * per, r1 = (r0 + per < M) ? r0 + per : M;
  double s0 = 0.0, s1 = 0.0;
  if (mode == 0) {
    for (long long r = r0; r < r1; ++r) {
      const float v = x[r * C + c];
      s0 += (double)v; s1 += (double)v * (double)v;
    }
  } else {
    const float mu = mean[c], rs = rstd[c];
    for (long long r = r0; r < r1; ++r) {
      const float g = dy[r * C + c];
      const float xh = (x[r * C + c] - mu) * rs;
      s0 += (double)g; s1 += (double)g * (double)xh;
    }
  }
  partial[((long long)blockIdx.x * 2 + 0) * C + c] = s0;
  partial[((long long)blockIdx.x * 2 + 1) * C + c] = s1;
}
static int bn_partial(const float* dy, const float* x, const float* mean, const float* rstd, long long M, int C, int mode,
                      double* partial, cudaStream_t st) {
  bn_partial_kernel<<<dim3(bn_chunks(M), (C + 127) / 128), 128, 0, st>>>(dy, x, mean, rstd, M, C, mode, partial);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// sums[0..C) = sum_r a, sums[C..2C) = sum_r a*b over all chunks (fixed order), sums[2C] = M: the unit a SyncBatchNorm
// exchange adds up over the ranks
__global__ void bn_reduce_chunks_kernel(const double* __restrict__ partial, int chunks, long long M, int C,
                                        double* __restrict__ sums) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) sums[2 * C] = (double)M;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < chunks; ++k) { s += partial[((long long)k * 2) * C + c]; q += partial[((long long)k * 2 + 1) * C + c]; }
  sums[c] = s;
  sums[C + c] = q;
}
__global__ void bn_finalize_fwd_kernel(const double* __restrict__ sums, int C, float eps, float momentum, float* __restrict__ mean,
                                       float* __restrict__ rstd, float* __restrict__ running_mean,
                                       float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double M = sums[2 * C];
  const double mu = sums[c] / M;
  double var = sums[C + c] / M - mu * mu;                  // biased: what the normalisation uses
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) {                                      // nn.BatchNorm2d.train(): momentum update, UNBIASED variance
    const double unb = M > 1.0 ? var * M / (M - 1.0) : var;
    running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mu);
    running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
  }
}
// local column sums -> (optional) sum over the data-parallel ranks -> `sums` (device, 2C + 1 doubles)
static int bn_sums(const float* dy, const float* x, const float* mean, const float* rstd, long long M, int C, int mode,
                   double* partial, const BnExchange* ex, double** sums_out, cudaStream_t st) {
  MAED_PROPAGATE(bn_partial(dy, x, mean, rstd, M, C, mode, partial, st));
  double* sums = partial + bn_partial_doubles(M, C);       // tail of the scratch buffer (see bn_partial_doubles' caller contract)
  if (ex && ex->fn) {
    MAED_CHECK_ARG(ex->buf && ex->capacity >= 2 * C + 1, "BatchNorm exchange buffer too small (%d < %d doubles)", ex->capacity,
                   2 * C + 1);
    sums = ex->buf;
  }
  bn_reduce_chunks_kernel<<<(C + 127) / 128, 128, 0, st>>>(partial, bn_chunks(M), M, C, sums);
  MAED_BW_LAUNCH_CHECK();
  if (ex && ex->fn) {
    const int rc = ex->fn(ex->user, 2 * C + 1);            // e.g. torch.distributed.all_reduce on the caller's stream
    MAED_CHECK_ARG(rc == 0, "BatchNorm statistics exchange failed (callback returned %d)", rc);
  }
  *sums_out = sums;
  return MAED_OK;
}
int bn_train_stats(const float* x, long long M, int C, float eps, float momentum, double* partial, float* mean, float* rstd,
                   float* running_mean, float* running_var, const BnExchange* ex, cudaStream_t st) {
  MAED_CHECK_ARG(x && partial && mean && rstd && M >= 1 && C >= 1, "bn_train_stats: bad argument");
  MAED_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "bn_train_stats: running buffers come in pairs");
  double* sums;
  MAED_PROPAGATE(bn_sums(nullptr, x, nullptr, nullptr, M, C, 0, partial, ex, &sums, st));
  bn_finalize_fwd_kernel<<<(C + 127) / 128, 128, 0, st>>>(sums, C, eps, momentum, mean, rstd, running_mean, running_var);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// y = relu?((x - mean) * rstd * gamma + beta (+ residual planes)) -> planes
__global__ void bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, long long total4, int C, int relu,
                                const __half* __restrict__ res_hi, long long res_plane, __half* __restrict__ out_hi,
                                long long out_plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c), b = *reinterpret_cast<const float4*>(beta + c);
    float4 y = make_float4((v.x - mu.x) * rs.x * g.x + b.x, (v.y - mu.y) * rs.y * g.y + b.y, (v.z - mu.z) * rs.z * g.z + b.z,
                           (v.w - mu.w) * rs.w * g.w + b.w);
    if (res_hi) {
      const float4 r = bw::load_planes4(res_hi + 4 * i, res_plane);
      y.x += r.x; y.y += r.y; y.z += r.z; y.w += r.w;
    }
    if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
    store_split4(out_hi + 4 * i, out_plane, y);
  }
}
int bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, long long M, int C,
             int relu, const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(x && mean && rstd && gamma && beta && out_hi, "bn_apply: null argument");
  MAED_CHECK_ARG(C % 4 == 0 && out_plane % 4 == 0 && (!res_hi || res_plane % 4 == 0), "bn_apply: C and plane strides must be multiples of 4");
  const long long total4 = M * (C / 4);
  bn_apply_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x, mean, rstd, gamma, beta, total4, C, relu, res_hi, res_plane, out_hi,
                                                        out_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// d *= (bn(x) > 0): the ReLU behind a BatchNorm whose output was not kept (the stem, whose output only survives max-pooled)
__global__ void bn_relu_mask_kernel(float* __restrict__ d, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    long long total, int C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    if ((x[i] - mean[c]) * rstd[c] * gamma[c] + beta[c] <= 0.f) d[i] = 0.f;
  }
}
int bn_relu_mask(float* d, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, long long M,
                 int C, cudaStream_t st) {
  bn_relu_mask_kernel<<<grid_for(M * C, 256), 256, 0, st>>>(d, x, mean, rstd, gamma, beta, M * C, C);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// dgamma / dbeta from THIS rank's sums (the data-parallel gradient reduction adds the ranks up, as with SyncBatchNorm)
__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float scale, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = scale * (float)sums[c];
  dgamma[c] = scale * (float)sums[C + c];
}
// dx = gamma * rstd * (dy - mean(dy) - xhat * mean(dy * xhat)) -> planes; the means are over all rows of all ranks
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean,
                                    const float* __restrict__ rstd, const float* __restrict__ gamma, const double* __restrict__ sums,
                                    long long total4, int C, __half* __restrict__ dx_hi, long long dx_plane) {
  const int c4n = C >> 2;
  const float inv_m = (float)(1.0 / sums[2 * C]);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const float4 g = reinterpret_cast<const float4*>(dy)[i], v = reinterpret_cast<const float4*>(x)[i];
    const float gv[4] = {g.x, g.y, g.z, g.w}, xv[4] = {v.x, v.y, v.z, v.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float rs = rstd[c + k];
      const float xh = (xv[k] - mean[c + k]) * rs;
      o[k] = gamma[c + k] * rs * (gv[k] - (float)sums[c + k] * inv_m - xh * (float)sums[C + c + k] * inv_m);
    }
    store_split4(dx_hi + 4 * i, dx_plane, make_float4(o[0], o[1], o[2], o[3]));
  }
}
int bn_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, long long M, int C,
           float scale, double* partial, float* dgamma, float* dbeta, __half* dx_hi, long long dx_plane, const BnExchange* ex,
           cudaStream_t st) {
  MAED_CHECK_ARG(dy && x && mean && rstd && gamma && partial && dgamma && dbeta && dx_hi, "bn_bwd: null argument");
  MAED_CHECK_ARG(C % 4 == 0 && dx_plane % 4 == 0, "bn_bwd: C and the plane stride must be multiples of 4");
  // parameter gradients need the LOCAL sums: reduce without the exchange first, then exchange for dx
  double* local;
  MAED_PROPAGATE(bn_sums(dy, x, mean, rstd, M, C, 1, partial, nullptr, &local, st));
  bn_param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(local, C, scale, dgamma, dbeta);
  MAED_BW_LAUNCH_CHECK();
  const double* sums = local;
  if (ex && ex->fn) {
    MAED_CHECK_ARG(ex->buf && ex->capacity >= 2 * C + 1, "BatchNorm exchange buffer too small");
    MAED_CUDA_CHECK(cudaMemcpyAsync(ex->buf, local, (size_t)(2 * C + 1) * 8, cudaMemcpyDeviceToDevice, st));
    const int rc = ex->fn(ex->user, 2 * C + 1);
    MAED_CHECK_ARG(rc == 0, "BatchNorm gradient-statistics exchange failed (callback returned %d)", rc);
    sums = ex->buf;
  }
  const long long total4 = M * (C / 4);
  bn_bwd_apply_kernel<<<grid_for(total4, 256), 256, 0, st>>>(dy, x, mean, rstd, gamma, sums, total4, C, dx_hi, dx_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// nn.MaxPool2d(3, 2, 1) keeping the winning tap (0..8, first maximum in scan order like PyTorch) of every output element
__global__ void maxpool3x3s2_idx_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W, int C,
                                        int OH, int OW, long long total4, __half* __restrict__ out_hi, long long plane,
                                        unsigned char* __restrict__ idx) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long m = i / c4n;
    const int ow = (int)(m % OW);
    const int oh = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned char bi[4] = {0, 0, 0, 0};
    float mu[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f}, sh[4] = {0.f, 0.f, 0.f, 0.f};
    if (mean) {                                            // x -> relu(BatchNorm(x)) on the fly (the stem)
#pragma unroll
      for (int k = 0; k < 4; ++k) { mu[k] = mean[c + k]; sc[k] = rstd[c + k]; sh[k] = beta[c + k]; }
    }
    const float4 gm = mean ? *reinterpret_cast<const float4*>(gamma + c) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float gk[4] = {gm.x, gm.y, gm.z, gm.w};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - 1;
        if (iw < 0 || iw >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + ((n * H + ih) * W + iw) * C + c);
        float vv[4] = {v.x, v.y, v.z, v.w};
        if (mean) {
#pragma unroll
          for (int k = 0; k < 4; ++k) vv[k] = fmaxf((vv[k] - mu[k]) * sc[k] * gk[k] + sh[k], 0.f);   // same order as bn_apply
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (vv[k] > best[k]) { best[k] = vv[k]; bi[k] = (unsigned char)(r * 3 + s); }
      }
    }
    const long long o = m * C + c;
    store_split4(out_hi + o, plane, make_float4(best[0], best[1], best[2], best[3]));
    *reinterpret_cast<uchar4*>(idx + o) = make_uchar4(bi[0], bi[1], bi[2], bi[3]);
  }
}
int maxpool3x3s2_idx(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, int n_img, int H,
                     int W, int C, __half* out_hi, long long plane, unsigned char* idx, cudaStream_t st) {
  MAED_CHECK_ARG(x && out_hi && idx && C % 4 == 0 && plane % 4 == 0, "maxpool3x3s2_idx: bad argument");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total4 = (long long)n_img * OH * OW * (C / 4);
  maxpool3x3s2_idx_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x, mean, rstd, gamma, beta, H, W, C, OH, OW, total4, out_hi, plane,
                                                                idx);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
// d_x[n, ih, iw, c] = sum of d_out over the (at most 2 x 2) windows whose arg-max is this pixel (gather: deterministic)
__global__ void maxpool3x3s2_bwd_kernel(const float* __restrict__ d_out, const unsigned char* __restrict__ idx, int H, int W, int C,
                                        int OH, int OW, long long total, float* __restrict__ d_x) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long p = i / C;
    const int iw = (int)(p % W);
    const int ih = (int)((p / W) % H);
    const long long n = p / ((long long)W * H);
    float acc = 0.f;
    for (int r = 0; r < 3; ++r) {
      const int t = ih + 1 - r;                            // ih = 2 * oh + r - 1
      if (t < 0 || (t & 1)) continue;
      const int oh = t >> 1;
      if (oh >= OH) continue;
      for (int s = 0; s < 3; ++s) {
        const int u = iw + 1 - s;
        if (u < 0 || (u & 1)) continue;
        const int ow = u >> 1;
        if (ow >= OW) continue;
        const long long o = ((n * OH + oh) * OW + ow) * C + c;
        if (idx[o] == r * 3 + s) acc += d_out[o];
      }
    }
    d_x[i] = acc;
  }
}
int maxpool3x3s2_bwd(const float* d_out, const unsigned char* idx, int n_img, int H, int W, int C, float* d_x, cudaStream_t st) {
  MAED_CHECK_ARG(d_out && idx && d_x, "maxpool3x3s2_bwd: null argument");
  const int OH = (H - 1) / 2 + 1, OW = (W - 1) / 2 + 1;
  const long long total = (long long)n_img * H * W * C;
  maxpool3x3s2_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(d_out, idx, H, W, C, OH, OW, total, d_x);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// AdaptiveAvgPool2d(1) backward: d_map[bt, p, c] = d_feat[bt, c] / P
__global__ void avgpool_bwd_kernel(const float* __restrict__ d_feat, int P, int C, long long total, float inv_p,
                                   float* __restrict__ d_map) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / ((long long)P * C);
    d_map[i] = d_feat[bt * C + c] * inv_p;
  }
}
int avgpool_bwd(const float* d_feat, int BT, int P, int C, float* d_map, cudaStream_t st) {
  const long long total = (long long)BT * P * C;
  avgpool_bwd_kernel<<<grid_for(total, 256), 256, 0, st>>>(d_feat, P, C, total, 1.0f / (float)P, d_map);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// dW OIHW = scale * g[co][(kh, kw), ci]  (split-K output rows of stride k_pad) — convs without weight standardisation
__global__ void wgrad_permute_kernel(const float* __restrict__ g, int k_pad, int Cin, int KH, int KW, long long total, float scale,
                                     float* __restrict__ dw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int kw = (int)(i % KW);
    const int kh = (int)((i / KW) % KH);
    const int ci = (int)((i / ((long long)KW * KH)) % Cin);
    const long long co = i / ((long long)KW * KH * Cin);
    dw[i] = scale * g[co * k_pad + (long long)(kh * KW + kw) * Cin + ci];
  }
}
int wgrad_permute(const float* g, int k_pad, int Cout, int Cin, int KH, int KW, float scale, float* dw, cudaStream_t st) {
  const long long total = (long long)Cout * Cin * KH * KW;
  wgrad_permute_kernel<<<grid_for(total, 256), 256, 0, st>>>(g, k_pad, Cin, KH, KW, total, scale, dw);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// dst[r, 0..n) (+)= src[r, 0..n) for R rows with independent row strides (slices of the iterative regressor's input gradient)
__global__ void add_cols_kernel(float* __restrict__ dst, int ldd, const float* __restrict__ src, int lds, long long total, int n,
                                int accumulate) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / n;
    const int c = (int)(i % n);
    const float v = src[r * lds + c];
    dst[r * ldd + c] = accumulate ? dst[r * ldd + c] + v : v;
  }
}
int add_cols_f32(float* dst, int ldd, const float* src, int lds, int R, int n, int accumulate, cudaStream_t st) {
  const long long total = (long long)R * n;
  add_cols_kernel<<<grid_for(total, 256), 256, 0, st>>>(dst, ldd, src, lds, total, n, accumulate);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
