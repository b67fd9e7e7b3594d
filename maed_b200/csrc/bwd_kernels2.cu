// Backward kernels of the MAED training path, part 2: attentive-addition backward, token sums, the small fp32 GEMM
// of the tail, KTD kinematic-tree backward, Adam.
#include "bwd_kernels.h"

#include "device_utils.cuh"

namespace maed {
using namespace bw;

// ------------------------------------------------------------- parallel-mode attentive addition backward
// forward (kernels.cu ts_blend): (a_s, a_t) = softmax(logits[bt, 2c], logits[bt, 2c+1]);  ao = x_t*a_t + x_s*a_s.
// One thread per (frame, channel): pass 1 reduces over the tokens, pass 2 writes d_xs / d_xt.
__global__ void blend_bwd_kernel(const float* __restrict__ d_ao, const float* __restrict__ xs, const float* __restrict__ xt,
                                 const float* __restrict__ logits, int ntok, int C, float* __restrict__ d_logits,
                                 float* __restrict__ d_xs, float* __restrict__ d_xt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long bt = blockIdx.y;
  if (c >= C) return;
  const float ls = logits[bt * 2 * C + 2 * c], lt = logits[bt * 2 * C + 2 * c + 1];
  const float m = fmaxf(ls, lt);
  const float es = expf(ls - m), et = expf(lt - m);
  const float as = es / (es + et), at = et / (es + et);
  float gs = 0.f, gt = 0.f;                                  // dL/da_s, dL/da_t
  const long long base = bt * ntok * C + c;
  for (int t = 0; t < ntok; ++t) {
    const float d = d_ao[base + (long long)t * C];
    gs += d * xs[base + (long long)t * C];
    gt += d * xt[base + (long long)t * C];
  }
  const float dot = as * gs + at * gt;
  d_logits[bt * 2 * C + 2 * c] = as * (gs - dot);
  d_logits[bt * 2 * C + 2 * c + 1] = at * (gt - dot);
  for (int t = 0; t < ntok; ++t) {
    const float d = d_ao[base + (long long)t * C];
    d_xs[base + (long long)t * C] = d * as;
    d_xt[base + (long long)t * C] = d * at;
  }
}
int blend_bwd(const float* d_ao, const float* x_s, const float* x_t, const float* logits, int BT, int ntok, int C,
              float* d_logits, float* d_xs, float* d_xt, cudaStream_t st) {
  blend_bwd_kernel<<<dim3(cdiv(C, 128), BT), 128, 0, st>>>(d_ao, x_s, x_t, logits, ntok, C, d_logits, d_xs, d_xt);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
__global__ void blend_bwd_pool_kernel(const float* __restrict__ d_pool, int ntok, int C, float inv_ntok, long long total4,
                                      float* __restrict__ d_xs, float* __restrict__ d_xt) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long bt = i / ((long long)c4n * ntok);
    const float4 ps = *reinterpret_cast<const float4*>(d_pool + bt * 2 * C + c);
    const float4 pt = *reinterpret_cast<const float4*>(d_pool + bt * 2 * C + C + c);
    float4 a = reinterpret_cast<float4*>(d_xs)[i];
    float4 b = reinterpret_cast<float4*>(d_xt)[i];
    a.x += ps.x * inv_ntok; a.y += ps.y * inv_ntok; a.z += ps.z * inv_ntok; a.w += ps.w * inv_ntok;
    b.x += pt.x * inv_ntok; b.y += pt.y * inv_ntok; b.z += pt.z * inv_ntok; b.w += pt.w * inv_ntok;
    reinterpret_cast<float4*>(d_xs)[i] = a;
    reinterpret_cast<float4*>(d_xt)[i] = b;
  }
}
int blend_bwd_pool(const float* d_pool, int BT, int ntok, int C, float* d_xs, float* d_xt, cudaStream_t st) {
  const long long total4 = (long long)BT * ntok * C / 4;
  blend_bwd_pool_kernel<<<grid_for(total4, 256), 256, 0, st>>>(d_pool, ntok, C, 1.0f / (float)ntok, total4, d_xs, d_xt);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void token_sum_kernel(const float* __restrict__ x, int ntok, int C, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const long long bt = blockIdx.y;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < ntok; ++t) s += x[(bt * ntok + t) * C + c];
  out[bt * C + c] = s;
}
int token_sum(const float* x, int BT, int ntok, int C, float* out, cudaStream_t st) {
  token_sum_kernel<<<dim3(cdiv(C, 128), BT), 128, 0, st>>>(x, ntok, C, out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------------------- small fp32 GEMM
// 64x64 output tile, 16-deep K slices, 256 threads each owning a 4x4 micro-tile.  Only used where one dimension is
// the frame count (tail / ts_attn linears): ~1 GFLOP per step in total.
template <int TA, int TB>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float beta, float* __restrict__ C, int ldc) {
  __shared__ float sA[16][65];                               // [k][m]
  __shared__ float sB[16][65];                               // [k][n]
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int kk, mm;
      if (TA) { mm = e & 63; kk = e >> 6; } else { kk = e & 15; mm = e >> 4; }   // contiguous dimension fastest
      const int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < M && k < K) v = TA ? A[(long long)k * lda + m] : A[(long long)m * lda + k];
      sA[kk][mm] = v;
    }
    for (int e = threadIdx.x; e < 16 * 64; e += 256) {
      int kk, nn;
      if (TB) { kk = e & 15; nn = e >> 4; } else { nn = e & 63; kk = e >> 6; }
      const int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < N && k < K) v = TB ? B[(long long)n * ldb + k] : B[(long long)k * ldb + n];
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* c = C + (long long)m * ldc + n;
      *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}
int sgemm_f32(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda, const float* B, int ldb,
              float beta, float* C, int ldc, cudaStream_t st) {
  MAED_CHECK_ARG(M > 0 && N > 0 && K > 0 && A && B && C, "sgemm_f32: bad arguments M=%d N=%d K=%d", M, N, K);
  const dim3 grid(cdiv(N, 64), cdiv(M, 64));
  if (!transA && !transB) sgemm_kernel<0, 0><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (!transA && transB) sgemm_kernel<0, 1><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else if (transA && !transB) sgemm_kernel<1, 0><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  else sgemm_kernel<1, 1><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, B, ldb, beta, C, ldc);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------- KTD kinematic tree backward
// reference lib/models/ktd.py:10-35 (ANCESTOR_INDEX) — same tables as decoder.cu
__constant__ int c_banc_cnt[24] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8};
__constant__ int c_banc[24][8] = {
    {0}, {0}, {0}, {0}, {0, 1}, {0, 2}, {0, 3}, {0, 1, 4}, {0, 2, 5}, {0, 3, 6}, {0, 1, 4, 7}, {0, 2, 5, 8},
    {0, 3, 6, 9}, {0, 3, 6, 9}, {0, 3, 6, 9}, {0, 3, 6, 9, 12}, {0, 3, 6, 9, 13}, {0, 3, 6, 9, 14},
    {0, 3, 6, 9, 13, 16}, {0, 3, 6, 9, 14, 17}, {0, 3, 6, 9, 13, 16, 18}, {0, 3, 6, 9, 14, 17, 19},
    {0, 3, 6, 9, 13, 16, 18, 20}, {0, 3, 6, 9, 14, 17, 19, 21}};
__constant__ int c_banc_woff[24] = {0, 0, 36, 72, 108, 180, 252, 324, 432, 540, 648, 792, 936, 1080, 1224, 1368, 1548,
                                    1728, 1908, 2124, 2340, 2592, 2844, 3132};

// forward: pose_j = base_j + sum_a W_j[:, 6a:6a+6] . pose_{anc(j,a)}.  Joints are visited in reverse order, so every
// descendant has already pushed its contribution into g[anc] when joint `anc` is finalised.  One thread per frame.
__global__ void ktd_tree_bwd_kernel(const float* __restrict__ d_pose, const float* __restrict__ d_shape,
                                    const float* __restrict__ d_cam, const float* __restrict__ w_anc, int R,
                                    float* __restrict__ g_total, float* __restrict__ d_base, int ld) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float g[144];
  for (int i = 0; i < 144; ++i) g[i] = d_pose[(long long)r * 144 + i];
  for (int j = 23; j >= 0; --j) {
    const int cnt = c_banc_cnt[j];
    const int woff = c_banc_woff[j];
    for (int a = 0; a < cnt; ++a) {
      const int aj = c_banc[j][a];
      for (int i = 0; i < 6; ++i) {
        const float gi = g[j * 6 + i];
        const float* wr = w_anc + woff + i * 6 * cnt + a * 6;
#pragma unroll
        for (int e = 0; e < 6; ++e) g[aj * 6 + e] += __ldg(wr + e) * gi;
      }
    }
  }
  float* db = d_base + (long long)r * ld;
  for (int i = 0; i < 144; ++i) { g_total[(long long)r * 144 + i] = g[i]; db[i] = g[i]; }
  for (int i = 0; i < 10; ++i) db[144 + i] = d_shape[(long long)r * 10 + i];
  for (int i = 0; i < 3; ++i) db[154 + i] = d_cam[(long long)r * 3 + i];
  for (int i = 157; i < ld; ++i) db[i] = 0.f;
}
int ktd_tree_bwd(const float* d_pose6d, const float* d_shape, const float* d_cam, const float* w_anc, int R, float* g_total,
                 float* d_base, int ld, cudaStream_t st) {
  MAED_CHECK_ARG(ld >= 157, "ktd_tree_bwd: ld=%d < 157", ld);
  ktd_tree_bwd_kernel<<<cdiv(R, 32), 32, 0, st>>>(d_pose6d, d_shape, d_cam, w_anc, R, g_total, d_base, ld);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
// d_w_anc block of joint j: [6][6*cnt] with entry (i, 6a+e) = sum_r g_total[r, 6j+i] * pose[r, 6*anc(j,a)+e].
// One warp per weight entry (3420 entries), lanes stride over the frames.
__global__ void ktd_anc_wgrad_kernel(const float* __restrict__ g_total, const float* __restrict__ pose, int R, float scale,
                                     float* __restrict__ d_w) {
  const int entry = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (entry >= 36 * 95) return;
  int j = 23;
  while (c_banc_woff[j] > entry) --j;                        // joints with cnt = 0 have an empty block
  while (c_banc_cnt[j] == 0) --j;                            // (entry 0 belongs to joint 1)
  const int cnt = c_banc_cnt[j];
  const int loc = entry - c_banc_woff[j];
  const int i = loc / (6 * cnt), col = loc % (6 * cnt);
  const int a = col / 6, e = col % 6;
  const int aj = c_banc[j][a];
  float s = 0.f;
  for (int r = lane; r < R; r += 32) s += g_total[(long long)r * 144 + j * 6 + i] * pose[(long long)r * 144 + aj * 6 + e];
  s = warp_sum(s);
  if (lane == 0) d_w[entry] = scale * s;
}
int ktd_anc_wgrad(const float* g_total, const float* pose6d, int R, float scale, float* d_w_anc, cudaStream_t st) {
  ktd_anc_wgrad_kernel<<<cdiv(36 * 95, 8), 256, 0, st>>>(g_total, pose6d, R, scale, d_w_anc);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------------------ Adam
// torch.optim.Adam (reference lib/utils/utils.py:127-131): g += wd * p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float omb1, float b2, float omb2, float eps, float wd, float step_size, float bc2_sqrt,
                            float grad_scale) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = g[i] * grad_scale + wd * pi;
    const float m0 = m[i];
    const float mi = m0 + omb1 * (gi - m0);               // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = b2 * v[i] + omb2 * gi * gi;          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}
int adam_step(float* p, const float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
              double weight_decay, int step, float grad_scale, cudaStream_t st) {
  MAED_CHECK_ARG(step >= 1, "adam_step: step counts from 1");
  // scalar coefficients in double, rounded to fp32 once (torch/optim/adam.py::_single_tensor_adam)
  const double bc1 = 1.0 - pow(beta1, (double)step);
  const double bc2_sqrt = sqrt(1.0 - pow(beta2, (double)step));
  adam_kernel<<<grid_for(n, 256), 256, 0, st>>>(p, g, m, v, n, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                               (float)weight_decay, (float)(lr / bc1), (float)bc2_sqrt, grad_scale);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ----------------------------------------------------------------- D = rowsum(dO o O) per (row, head) for the attention backward
// One warp per row; a lane owns every 32nd float4 of the row, 16 lanes share a head (64 columns = 16 float4): heads 2k and 2k + 1
// are reduced in the two half-warps.
__global__ void __launch_bounds__(256)
attn_rowdot_kernel(const float* __restrict__ d_out, const float* __restrict__ o_f32, const __half* __restrict__ o_hi, long long o_plane,
                   long long rows, int heads, float* __restrict__ D) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int C = heads * 64;
  const float* dr = d_out + row * C;
  for (int k = 0; 2 * k < heads; ++k) {
    const int col = (k * 32 + lane) * 4;                   // head = 2k + lane / 16
    float s = 0.f;
    if (col < C) {
      const float4 d = *reinterpret_cast<const float4*>(dr + col);
      float4 o;
      if (o_f32) o = *reinterpret_cast<const float4*>(o_f32 + row * C + col);
      else o = load_planes4(o_hi + row * C + col, o_plane);
      s = d.x * o.x + d.y * o.y + d.z * o.z + d.w * o.w;
    }
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const int h = 2 * k + (lane >> 4);
    if ((lane & 15) == 0 && h < heads) D[row * heads + h] = s;
  }
}
int attn_rowdot(const float* d_out, const float* o_f32, const __half* o_hi, long long o_plane, long long rows, int heads, float* D,
                cudaStream_t st) {
  MAED_CHECK_ARG(d_out && D && ((o_f32 != nullptr) != (o_hi != nullptr)), "attn_rowdot: d_out, D and exactly one form of O required");
  MAED_CHECK_ARG(rows >= 1 && heads >= 1, "attn_rowdot: bad shape");
  attn_rowdot_kernel<<<(unsigned)cdiv(rows, 8), 256, 0, st>>>(d_out, o_f32, o_hi, o_plane, rows, heads, D);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
