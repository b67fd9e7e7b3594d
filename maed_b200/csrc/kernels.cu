// Memory-bound kernels of the MAED hot path: precision split, weight standardisation, im2col gathers,
// GroupNorm, LayerNorm, max-pool, embedding assembly, attentive addition.  All HBM-bound: coalesced
// 16-byte accesses along the channel dimension, grids sized to a few waves of the 148 SMs.
#include "kernels.h"

#include <atomic>

namespace maed {

static std::atomic<long long> g_launches{0};
long long launch_count() { return g_launches.load(); }
void count_launch(int n) { g_launches.fetch_add(n); }

#define LAUNCH_CHECK()                      \
  do {                                      \
    count_launch();                         \
    MAED_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
  hi = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(hi);
  lo = __floats2half2_rn(a - hf.x, b - hf.y);
}
__device__ __forceinline__ void store_split4(__half* hi_ptr, long long plane, float4 v) {
  __half2 h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  uint2 H, L;
  H.x = *reinterpret_cast<uint32_t*>(&h0); H.y = *reinterpret_cast<uint32_t*>(&h1);
  L.x = *reinterpret_cast<uint32_t*>(&l0); L.y = *reinterpret_cast<uint32_t*>(&l1);
  *reinterpret_cast<uint2*>(hi_ptr) = H;
  *reinterpret_cast<uint2*>(hi_ptr + plane) = L;
}
__device__ __forceinline__ float4 load_planes4(const __half* hi_ptr, long long plane) {
  const uint2 H = *reinterpret_cast<const uint2*>(hi_ptr);
  const uint2 L = *reinterpret_cast<const uint2*>(hi_ptr + plane);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&H.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&H.y));
  const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&L.x));
  const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&L.y));
  return make_float4(a.x + c.x, a.y + c.y, b.x + d.x, b.y + d.y);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ------------------------------------------------------------------------------------------ split
__global__ void split_f32_kernel(const float* __restrict__ in, __half* __restrict__ hi, long long plane, long long n) {
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    store_split4(hi + 4 * i, plane, reinterpret_cast<const float4*>(in)[i]);
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    const __half h = __float2half_rn(in[i]);
    hi[i] = h;
    hi[i + plane] = __float2half_rn(in[i] - __half2float(h));
  }
}
int split_f32(const float* in, __half* out_hi, long long plane, long long n, cudaStream_t st) {
  MAED_CHECK_ARG(plane % 4 == 0 || n < 4, "split_f32: plane stride must be a multiple of 4");
  split_f32_kernel<<<grid_for(n / 4 + 1, 256), 256, 0, st>>>(in, out_hi, plane, n);
  LAUNCH_CHECK();
  return MAED_OK;
}

// -------------------------------------------------------------------------- conv weight preparation
// One block per output channel.  Two-pass mean / biased variance in fp32 (E <= 9216 elements).
__global__ void prep_conv_weight_kernel(const float* __restrict__ w, int Cin, int KH, int KW, int k_pad,
                                        int standardize, __half* __restrict__ hi, long long plane) {
  __shared__ float red[32];
  __shared__ float s_mean, s_inv;
  const int co = blockIdx.x;
  const int E = Cin * KH * KW;
  const float* wc = w + (long long)co * E;
  float mean = 0.f, inv = 1.f;
  if (standardize) {
    float s = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) s += wc[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) s_mean = t / (float)E;
    }
    __syncthreads();
    mean = s_mean;
    float q = 0.f;
    for (int i = threadIdx.x; i < E; i += blockDim.x) { const float d = wc[i] - mean; q += d * d; }
    q = warp_sum(q);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
    __syncthreads();
    if (threadIdx.x < 32) {
      float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
      t = warp_sum(t);
      if (threadIdx.x == 0) s_inv = 1.0f / (sqrtf(t / (float)E) + 1e-5f);   // eps added to std (resnetv2.py:88)
    }
    __syncthreads();
    inv = s_inv;
  }
  __half* oh = hi + (long long)co * k_pad;
  for (int k = threadIdx.x; k < k_pad; k += blockDim.x) {
    float v = 0.f;
    if (k < E) {
      const int c = k % Cin, tap = k / Cin;
      const int r = tap / KW, s = tap % KW;
      v = (wc[(c * KH + r) * KW + s] - mean) * inv;
    }
    const __half h = __float2half_rn(v);
    oh[k] = h;
    oh[k + plane] = __float2half_rn(v - __half2float(h));
  }
}
int prep_conv_weight(const float* w, int Cout, int Cin, int KH, int KW, int k_pad, int standardize, __half* out_hi,
                     long long plane, cudaStream_t st) {
  prep_conv_weight_kernel<<<Cout, 256, 0, st>>>(w, Cin, KH, KW, k_pad, standardize, out_hi, plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------------- im2col
// Stem gather: a block builds the A rows of 32 consecutive output pixels of one output row in shared
// memory (reads coalesced along the input row), then writes whole rows (16-byte stores).
// CIN_ / KW_ > 0: compile-time shape (the 7x7 RGB stem: divisions by constants), 0: the runtime arguments
template <int CIN_, int KW_>
__global__ void im2col_stem_kernel(const float* __restrict__ x, int Cin_rt, int H, int W, int KH, int KW_rt, int stride,
                                   int pad_t, int pad_l, int OH, int OW, int k_pad, __half* __restrict__ hi,
                                   long long plane) {
  extern __shared__ float tile[];                       // [32][k_pad + 4]
  const int Cin = CIN_ > 0 ? CIN_ : Cin_rt, KW = KW_ > 0 ? KW_ : KW_rt;
  const int ldt = k_pad + 4;                            // row stride = 28 (mod 32) for k_pad = 152: see the gather below
  const int tiles_w = (OW + 31) / 32;
  const int tw = blockIdx.x % tiles_w;
  const int oh = (blockIdx.x / tiles_w) % OH;
  const int n = blockIdx.x / (tiles_w * OH);
  const int ow0 = tw * 32;
  const int Kreal = KH * KW * Cin;
  // gather: a warp fills 8 pixels x 4 columns per step (lane = 8 * column + pixel): with a row stride of 4 * odd floats the
  // 32 shared-memory writes fall into 32 different banks, and every group of 8 lanes reads one contiguous run of the input row
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int units = 4 * (k_pad / 4);
  for (int u = warp; u < units; u += nwarps) {
    const int px = (u & 3) * 8 + (lane & 7), k = (u >> 2) * 4 + (lane >> 3);
    float v = 0.f;
    const int ow = ow0 + px;
    if (k < Kreal && ow < OW) {
      const int c = k % Cin, tap = k / Cin;
      const int r = tap / KW, s2 = tap % KW;
      const int ih = oh * stride + r - pad_t, iw = ow * stride + s2 - pad_l;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((long long)n * Cin + c) * H + ih) * W + iw];
    }
    tile[px * ldt + k] = v;
  }
  __syncthreads();
  // rows leave as 16-byte stores: 8 columns -> 8 hi halfs + 8 lo halfs (k_pad is a multiple of 8)
  const int chunks = k_pad / 8;
  for (int idx = threadIdx.x; idx < 32 * chunks; idx += blockDim.x) {
    const int px = idx / chunks, ch = idx % chunks;
    const int ow = ow0 + px;
    if (ow >= OW) continue;
    const long long m = ((long long)n * OH + oh) * OW + ow;
    const float4 v0 = *reinterpret_cast<const float4*>(&tile[px * ldt + ch * 8]);
    const float4 v1 = *reinterpret_cast<const float4*>(&tile[px * ldt + ch * 8 + 4]);
    __half2 h0, l0, h1, l1, h2, l2, h3, l3;
    split2(v0.x, v0.y, h0, l0); split2(v0.z, v0.w, h1, l1); split2(v1.x, v1.y, h2, l2); split2(v1.z, v1.w, h3, l3);
    __half* o = hi + m * k_pad + ch * 8;
    *reinterpret_cast<uint4*>(o) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                              *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
    if (plane)
      *reinterpret_cast<uint4*>(o + plane) = make_uint4(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1),
                                                        *reinterpret_cast<uint32_t*>(&l2), *reinterpret_cast<uint32_t*>(&l3));
  }
}
int im2col_stem(const float* x, int n_img, int Cin, int H, int W, int KH, int KW, int stride, int pad_t, int pad_l,
                int OH, int OW, int k_pad, __half* out_hi, long long plane, cudaStream_t st) {
  MAED_CHECK_ARG(k_pad % 8 == 0, "im2col_stem: k_pad must be a multiple of 8");
  const int tiles_w = (OW + 31) / 32;
  const size_t smem = (size_t)32 * (k_pad + 4) * sizeof(float);
  if (Cin == 3 && KW == 7)
    im2col_stem_kernel<3, 7><<<n_img * OH * tiles_w, 256, smem, st>>>(x, Cin, H, W, KH, KW, stride, pad_t, pad_l, OH, OW, k_pad,
                                                                     out_hi, plane);
  else
    im2col_stem_kernel<0, 0><<<n_img * OH * tiles_w, 256, smem, st>>>(x, Cin, H, W, KH, KW, stride, pad_t, pad_l, OH, OW, k_pad,
                                                                     out_hi, plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void im2col_nhwc_kernel(const __half* __restrict__ in, long long in_plane, int H, int W, int C, int KH, int KW,
                                   int stride, int pad_t, int pad_l, int OH, int OW, long long total_chunks,
                                   __half* __restrict__ out, long long out_plane) {
  const int c8n = C >> 3;
  const int taps = KH * KW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_chunks;
       i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const int tap = (int)((i / c8n) % taps);
    const long long m = i / ((long long)c8n * taps);
    const int ow = (int)(m % OW);
    const int oh = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    const int r = tap / KW, s = tap % KW;
    const int ih = oh * stride + r - pad_t, iw = ow * stride + s - pad_l;
    uint4 vh = make_uint4(0, 0, 0, 0), vl = make_uint4(0, 0, 0, 0);
    if (ih >= 0 && ih < H && iw >= 0 && iw < W) {
      const long long src = ((n * H + ih) * W + iw) * C + c8 * 8;
      vh = *reinterpret_cast<const uint4*>(in + src);
      vl = *reinterpret_cast<const uint4*>(in + src + in_plane);
    }
    const long long dst = (m * taps + tap) * C + c8 * 8;
    *reinterpret_cast<uint4*>(out + dst) = vh;
    *reinterpret_cast<uint4*>(out + dst + out_plane) = vl;
  }
}
int im2col_nhwc(const __half* in_hi, long long in_plane, int n_img, int H, int W, int C, int KH, int KW, int stride,
                int pad_t, int pad_l, int OH, int OW, __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(C % 8 == 0, "im2col_nhwc: C must be a multiple of 8");
  const long long total = (long long)n_img * OH * OW * KH * KW * (C / 8);
  im2col_nhwc_kernel<<<grid_for(total, 256), 256, 0, st>>>(in_hi, in_plane, H, W, C, KH, KW, stride, pad_t, pad_l, OH, OW,
                                                          total, out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------------------------- GroupNorm
// grid (chunks, n_img).  A thread owns 4 consecutive channels and strides over the pixel rows of its chunk.
__global__ void gn_stats_kernel(const float* __restrict__ x, int HW, int C, int rows_per_chunk, double* __restrict__ stats,
                                int reverse) {
  __shared__ double gs[32], gq[32];
  if (threadIdx.x < 32) { gs[threadIdx.x] = 0.0; gq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int n = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int c4n = C >> 2;
  const int rows_per_iter = blockDim.x / c4n;            // blockDim is a multiple of c4n (or c4n >= blockDim)
  const int gsz = C / 32;
  const int row_lo = blockIdx.x * rows_per_chunk;
  const int row_hi = min(HW, row_lo + rows_per_chunk);
  const float* xb = x + (long long)n * HW * C;
  if (rows_per_iter >= 1) {
    const int c4 = threadIdx.x % c4n, rsub = threadIdx.x / c4n;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (rsub < rows_per_iter) {
      for (int r = row_lo + rsub; r < row_hi; r += rows_per_iter) {
        const float4 v = *reinterpret_cast<const float4*>(xb + (long long)r * C + c4 * 4);
        s[0] += v.x; q[0] += v.x * v.x; s[1] += v.y; q[1] += v.y * v.y;
        s[2] += v.z; q[2] += v.z * v.z; s[3] += v.w; q[3] += v.w * v.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (c4 * 4 + j) / gsz;
        atomicAdd(&gs[g], (double)s[j]);
        atomicAdd(&gq[g], (double)q[j]);
      }
    }
  } else {                                               // C/4 > blockDim: several channel quads per thread
    for (int c4 = threadIdx.x; c4 < c4n; c4 += blockDim.x) {
      float s = 0.f, q = 0.f;
      for (int r = row_lo; r < row_hi; ++r) {
        const float4 v = *reinterpret_cast<const float4*>(xb + (long long)r * C + c4 * 4);
        s += v.x + v.y + v.z + v.w;
        q += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      const int g = (c4 * 4) / gsz;                      // gsz >= 4 here
      atomicAdd(&gs[g], (double)s);
      atomicAdd(&gq[g], (double)q);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    atomicAdd(&stats[((long long)n * 32 + threadIdx.x) * 2 + 0], gs[threadIdx.x]);
    atomicAdd(&stats[((long long)n * 32 + threadIdx.x) * 2 + 1], gq[threadIdx.x]);
  }
}
int gn_stats(const float* x, int n_img, int HW, int C, double* stats, cudaStream_t st, int reverse) {
  MAED_CHECK_ARG(C % 32 == 0 && C >= 32, "gn_stats: C=%d must be a multiple of 32", C);
  const int threads = 256;
  const int c4n = C / 4;
  const int rows_per_iter = threads / c4n > 0 ? threads / c4n : 1;
  int chunks = cdiv((long long)sm_count() * 8, n_img);
  const int max_chunks = cdiv(HW, rows_per_iter * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  const int rows_per_chunk = cdiv(HW, chunks);
  chunks = cdiv(HW, rows_per_chunk);
  gn_stats_kernel<<<dim3(chunks, n_img), threads, 0, st>>>(x, HW, C, rows_per_chunk, stats, reverse);
  LAUNCH_CHECK();
  return MAED_OK;
}

__device__ __forceinline__ void gn_load_stats(const double* stats, int n, int HW, int C, float eps, float* s_mean,
                                              float* s_rstd) {
  if (threadIdx.x < 32) {
    const double cnt = (double)HW * (C / 32);
    const double m = stats[((long long)n * 32 + threadIdx.x) * 2] / cnt;
    double var = stats[((long long)n * 32 + threadIdx.x) * 2 + 1] / cnt - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)m;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
}

// grid (blocks_per_img, n_img)
__global__ void gn_apply_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int HW, int C, float eps,
                                int relu, const __half* __restrict__ res, long long res_plane, __half* __restrict__ out,
                                long long out_plane, int reverse) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  gn_load_stats(stats, n, HW, C, eps, s_mean, s_rstd);
  const int c4n = C >> 2, gsz = C / 32;
  const long long total = (long long)HW * c4n;
  const long long base = (long long)n * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long off = base + (i / c4n) * C + c;
    float4 v = *reinterpret_cast<const float4*>(x + off);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    const int g0 = c / gsz, g1 = (c + 1) / gsz, g2 = (c + 2) / gsz, g3 = (c + 3) / gsz;
    v.x = (v.x - s_mean[g0]) * s_rstd[g0] * g.x + b.x;
    v.y = (v.y - s_mean[g1]) * s_rstd[g1] * g.y + b.y;
    v.z = (v.z - s_mean[g2]) * s_rstd[g2] * g.z + b.z;
    v.w = (v.w - s_mean[g3]) * s_rstd[g3] * g.w + b.w;
    if (res) {
      const float4 r = load_planes4(res + off, res_plane);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    store_split4(out + off, out_plane, v);
  }
}
int gn_apply(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int HW, int C,
             float eps, int relu, const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane,
             cudaStream_t st, int reverse) {
  const long long per_img = (long long)HW * C / 4;
  int bpi = cdiv((long long)sm_count() * 8, n_img);
  const int maxb = cdiv(per_img, 256);
  if (bpi > maxb) bpi = maxb;
  if (bpi < 1) bpi = 1;
  gn_apply_kernel<<<dim3(bpi, n_img), 256, 0, st>>>(x, stats, gamma, beta, HW, C, eps, relu, res_hi, res_plane, out_hi,
                                                    out_plane, reverse);
  LAUNCH_CHECK();
  return MAED_OK;
}

// stem: GN + ReLU + 3x3/2 max-pool with TF-SAME padding (-inf; extra pad on bottom/right)
__global__ void gn_apply_maxpool_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W, int C,
                                        float eps, int OH, int OW, int pad_t, int pad_l, __half* __restrict__ out,
                                        long long out_plane) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.y;
  gn_load_stats(stats, n, H * W, C, eps, s_mean, s_rstd);
  const int c4n = C >> 2, gsz = C / 32;
  const long long total = (long long)OH * OW * c4n;
  const float* xb = x + (long long)n * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int ow = (int)((i / c4n) % OW);
    const int oh = (int)(i / ((long long)c4n * OW));
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    const int g0 = c / gsz, g1 = (c + 1) / gsz, g2 = (c + 2) / gsz, g3 = (c + 3) / gsz;
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - pad_t;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - pad_l;
        if (iw < 0 || iw >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(xb + ((long long)ih * W + iw) * C + c);
        m.x = fmaxf(m.x, (v.x - s_mean[g0]) * s_rstd[g0] * g.x + b.x);
        m.y = fmaxf(m.y, (v.y - s_mean[g1]) * s_rstd[g1] * g.y + b.y);
        m.z = fmaxf(m.z, (v.z - s_mean[g2]) * s_rstd[g2] * g.z + b.z);
        m.w = fmaxf(m.w, (v.w - s_mean[g3]) * s_rstd[g3] * g.w + b.w);
      }
    }
    m.x = fmaxf(m.x, 0.f); m.y = fmaxf(m.y, 0.f); m.z = fmaxf(m.z, 0.f); m.w = fmaxf(m.w, 0.f);   // ReLU commutes with max
    const long long off = (((long long)n * OH + oh) * OW + ow) * C + c;
    store_split4(out + off, out_plane, m);
  }
}
int gn_apply_maxpool(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int H, int W,
                     int C, float eps, __half* out_hi, long long out_plane, cudaStream_t st) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const int pad_h = max((OH - 1) * 2 + 3 - H, 0), pad_w = max((OW - 1) * 2 + 3 - W, 0);
  const long long per_img = (long long)OH * OW * C / 4;
  int bpi = cdiv((long long)sm_count() * 8, n_img);
  const int maxb = cdiv(per_img, 256);
  if (bpi > maxb) bpi = maxb;
  if (bpi < 1) bpi = 1;
  gn_apply_maxpool_kernel<<<dim3(bpi, n_img), 256, 0, st>>>(x, stats, gamma, beta, H, W, C, eps, OH, OW, pad_h / 2,
                                                            pad_w / 2, out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------------------------------- STE
__global__ void embed_assemble_kernel(const float* __restrict__ tok, const float* __restrict__ cls,
                                      const float* __restrict__ pos, const float* __restrict__ temp, int T, int ntok, int C,
                                      long long total4, float* __restrict__ x) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int t = (int)((i / c4n) % ntok);
    const long long bt = i / ((long long)c4n * ntok);
    float4 v = (t == 0) ? *reinterpret_cast<const float4*>(cls + c)
                        : *reinterpret_cast<const float4*>(tok + (bt * (ntok - 1) + (t - 1)) * C + c);
    const float4 p = *reinterpret_cast<const float4*>(pos + (long long)t * C + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    if (temp) {
      const float4 e = *reinterpret_cast<const float4*>(temp + (bt % T) * C + c);
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    }
    *reinterpret_cast<float4*>(x + (bt * ntok + t) * C + c) = v;
  }
}
int embed_assemble(const float* tok, const float* cls, const float* pos, const float* temp, int BT, int T, int ntok, int C,
                   float* x, cudaStream_t st) {
  const long long total4 = (long long)BT * ntok * C / 4;
  embed_assemble_kernel<<<grid_for(total4, 256), 256, 0, st>>>(tok, cls, pos, temp, T, ntok, C, total4, x);
  LAUNCH_CHECK();
  return MAED_OK;
}

// LayerNorm: one warp per row, two-pass statistics in registers (C <= 1024, C % 128 == 0).
template <bool kPlanes>
__global__ void layernorm_kernel(const float* __restrict__ x, long long row_stride, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int rows, int C, float eps, float* __restrict__ out_f,
                                 __half* __restrict__ out_h, long long out_plane) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + (long long)row * row_stride;
  float4 v[8];
  const int nv = C >> 7;                                   // float4 per lane
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      v[j] = *reinterpret_cast<const float4*>(xr + (j * 32 + lane) * 4);
      s += v[j].x + v[j].y + v[j].z + v[j].w;
    }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const int c = (j * 32 + lane) * 4;
      const float4 g = *reinterpret_cast<const float4*>(gamma + c);
      const float4 b = *reinterpret_cast<const float4*>(beta + c);
      float4 o;
      o.x = (v[j].x - mean) * rstd * g.x + b.x;
      o.y = (v[j].y - mean) * rstd * g.y + b.y;
      o.z = (v[j].z - mean) * rstd * g.z + b.z;
      o.w = (v[j].w - mean) * rstd * g.w + b.w;
      if (kPlanes) store_split4(out_h + (long long)row * C + c, out_plane, o);
      else *reinterpret_cast<float4*>(out_f + (long long)row * C + c) = o;
    }
}
int layernorm_planes(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C, float eps,
                     __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(C % 128 == 0 && C <= 1024, "layernorm: C=%d unsupported", C);
  layernorm_kernel<true><<<cdiv(rows, 8), 256, 0, st>>>(x, row_stride, gamma, beta, rows, C, eps, nullptr, out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}
int layernorm_f32(const float* x, long long row_stride, const float* gamma, const float* beta, int rows, int C, float eps,
                  float* out, cudaStream_t st) {
  MAED_CHECK_ARG(C % 128 == 0 && C <= 1024, "layernorm: C=%d unsupported", C);
  layernorm_kernel<false><<<cdiv(rows, 8), 256, 0, st>>>(x, row_stride, gamma, beta, rows, C, eps, out, nullptr, 0);
  LAUNCH_CHECK();
  return MAED_OK;
}

// grid (C/128, BT): thread = one channel of a 128-channel slab, loops over tokens (coalesced rows)
__global__ void token_mean_kernel(const float* __restrict__ x, int ntok, int C, float* __restrict__ out, int out_ld, int col0) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const long long bt = blockIdx.y;
  const float* xb = x + bt * ntok * C + c;
  float s0 = 0.f, s1 = 0.f;
  int t = 0;
  for (; t + 1 < ntok; t += 2) { s0 += xb[(long long)t * C]; s1 += xb[(long long)(t + 1) * C]; }
  if (t < ntok) s0 += xb[(long long)t * C];
  out[bt * out_ld + col0 + c] = (s0 + s1) / (float)ntok;
}
int token_mean(const float* x, int BT, int ntok, int C, float* out, int out_ld, int col0, cudaStream_t st) {
  MAED_CHECK_ARG(C % 128 == 0, "token_mean: C must be a multiple of 128");
  token_mean_kernel<<<dim3(C / 128, BT), 128, 0, st>>>(x, ntok, C, out, out_ld, col0);
  LAUNCH_CHECK();
  return MAED_OK;
}

// stage 1: grid (kTokenChunks, BT, 2): partial column sums of a token chunk; thread = float4 of channels
__global__ void token_sum_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int ntok, int C,
                                         float* __restrict__ scratch) {
  const float* x = blockIdx.z == 0 ? a : b;
  const long long bt = blockIdx.y;
  const int per = (ntok + kTokenChunks - 1) / kTokenChunks;
  const int t0 = blockIdx.x * per, t1 = min(ntok, t0 + per);
  const int c = threadIdx.x * 4;
  if (c >= C) return;
  float4 s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
  const float* xb = x + bt * ntok * C + c;
  int t = t0;
  for (; t + 1 < t1; t += 2) {
    const float4 u = *reinterpret_cast<const float4*>(xb + (long long)t * C);
    const float4 v = *reinterpret_cast<const float4*>(xb + (long long)(t + 1) * C);
    s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
    s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
  }
  if (t < t1) {
    const float4 u = *reinterpret_cast<const float4*>(xb + (long long)t * C);
    s0.x += u.x; s0.y += u.y; s0.z += u.z; s0.w += u.w;
  }
  s0.x += s1.x; s0.y += s1.y; s0.z += s1.z; s0.w += s1.w;
  *reinterpret_cast<float4*>(scratch + ((bt * kTokenChunks + blockIdx.x) * 2 + blockIdx.z) * C + c) = s0;
}
// stage 2: sum the chunk partials in a fixed order, scale by 1/ntok, write planes [BT, 2C]
__global__ void token_mean_finalize_kernel(const float* __restrict__ scratch, int ntok, int C, long long total4,
                                           __half* __restrict__ out, long long out_plane) {
  const int c4n = (2 * C) >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % c4n) * 4;
    const long long bt = i / c4n;
    const int which = col / C, c = col % C;
    float4 s = make_float4(0, 0, 0, 0);
#pragma unroll
    for (int k = 0; k < kTokenChunks; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(scratch + ((bt * kTokenChunks + k) * 2 + which) * C + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const float inv = 1.0f / (float)ntok;
    s.x *= inv; s.y *= inv; s.z *= inv; s.w *= inv;
    store_split4(out + bt * 2 * C + col, out_plane, s);
  }
}
int token_mean2_planes(const float* a, const float* b, int BT, int ntok, int C, float* scratch, __half* out_hi,
                       long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(C % 4 == 0 && C <= 4096, "token_mean2: C=%d unsupported", C);
  token_sum_partial_kernel<<<dim3(kTokenChunks, BT, 2), (C / 4 + 31) / 32 * 32, 0, st>>>(a, b, ntok, C, scratch);
  LAUNCH_CHECK();
  const long long total4 = (long long)BT * 2 * C / 4;
  token_mean_finalize_kernel<<<grid_for(total4, 256), 256, 0, st>>>(scratch, ntok, C, total4, out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void ts_blend_kernel(const float* __restrict__ xs, const float* __restrict__ xt, const float* __restrict__ logits,
                                int ntok, int C, long long total4, __half* __restrict__ out, long long out_plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long row = i / c4n;
    const long long bt = row / ntok;
    const float4 l0 = *reinterpret_cast<const float4*>(logits + bt * 2 * C + 2 * c);       // (s,t) for c, c+1
    const float4 l1 = *reinterpret_cast<const float4*>(logits + bt * 2 * C + 2 * c + 4);   // (s,t) for c+2, c+3
    const float4 s = *reinterpret_cast<const float4*>(xs + row * C + c);
    const float4 t = *reinterpret_cast<const float4*>(xt + row * C + c);
    float4 o;
    {
      // softmax over the pair: a_s = 1/(1+exp(lt-ls)), a_t = 1 - a_s (computed like torch: exp(x-max)/sum)
      float m, es, et;
      m = fmaxf(l0.x, l0.y); es = expf(l0.x - m); et = expf(l0.y - m); o.x = (t.x * et + s.x * es) / (es + et);
      m = fmaxf(l0.z, l0.w); es = expf(l0.z - m); et = expf(l0.w - m); o.y = (t.y * et + s.y * es) / (es + et);
      m = fmaxf(l1.x, l1.y); es = expf(l1.x - m); et = expf(l1.y - m); o.z = (t.z * et + s.z * es) / (es + et);
      m = fmaxf(l1.z, l1.w); es = expf(l1.z - m); et = expf(l1.w - m); o.w = (t.w * et + s.w * es) / (es + et);
    }
    store_split4(out + row * C + c, out_plane, o);
  }
}
int ts_blend(const float* x_s, const float* x_t, const float* logits, int BT, int ntok, int C, __half* out_hi,
             long long out_plane, cudaStream_t st) {
  const long long total4 = (long long)BT * ntok * C / 4;
  ts_blend_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x_s, x_t, logits, ntok, C, total4, out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void broadcast_add_kernel(float* __restrict__ x, const float* __restrict__ v, int ntok, int C, long long total4) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long bt = i / ((long long)c4n * ntok);
    float4 a = *reinterpret_cast<float4*>(x + i * 4);
    const float4 b = *reinterpret_cast<const float4*>(v + bt * C + c);
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    *reinterpret_cast<float4*>(x + i * 4) = a;
  }
}
int broadcast_add(float* x, const float* v, int BT, int ntok, int C, cudaStream_t st) {
  const long long total4 = (long long)BT * ntok * C / 4;
  broadcast_add_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x, v, ntok, C, total4);
  LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
