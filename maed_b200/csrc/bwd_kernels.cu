// Backward kernels of the MAED training path, part 1: layout helpers, reductions, elementwise backward,
// LayerNorm / GroupNorm / weight-standardisation backward, stem max-pool, stride-2 helpers.
// All HBM-bound: coalesced accesses along the channel dimension; every reduction is a fixed-order two-stage
// sum (no floating-point atomics), so gradients are bit-reproducible run to run.
#include "bwd_kernels.h"

#include "device_utils.cuh"

namespace maed {
using namespace bw;

// ------------------------------------------------------------------------------------------ transpose
// 64x64 tiles, both planes through shared memory; out[c, r] = in[r, c].
__global__ void transpose_planes_kernel(const __half* __restrict__ in, long long in_plane, int R, int C, int ld_in,
                                        __half* __restrict__ out, long long out_plane, int ld_out) {
  __shared__ __half th[64][66];
  __shared__ __half tl[64][66];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
  const int c0 = blockIdx.x * 64;
  for (int r0 = blockIdx.y * 64; r0 < R; r0 += gridDim.y * 64) {
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int r = r0 + ty + 8 * k;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = c0 + 2 * tx + e;
      __half h = __float2half(0.f), l = __float2half(0.f);
      if (r < R && c < C) {
        h = in[(long long)r * ld_in + c];
        l = in[(long long)r * ld_in + c + in_plane];
      }
      th[ty + 8 * k][2 * tx + e] = h;
      tl[ty + 8 * k][2 * tx + e] = l;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = c0 + ty + 8 * k;                            // output row
    if (c >= C) continue;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = r0 + 2 * tx + e;                          // output column
      if (r < R) {
        out[(long long)c * ld_out + r] = th[2 * tx + e][ty + 8 * k];
        out[(long long)c * ld_out + r + out_plane] = tl[2 * tx + e][ty + 8 * k];
      }
    }
  }
  }
}
int transpose_planes(const __half* in_hi, long long in_plane, int R, int C, int ld_in, __half* out_hi, long long out_plane,
                     int ld_out, cudaStream_t st) {
  MAED_CHECK_ARG(R > 0 && C > 0 && ld_in >= C && ld_out >= R, "transpose_planes: bad shape R=%d C=%d ld_in=%d ld_out=%d", R, C,
                 ld_in, ld_out);
  const int row_blocks = cdiv(R, 64) < 32768 ? cdiv(R, 64) : 32768;      // the kernel strides over further row tiles
  transpose_planes_kernel<<<dim3(cdiv(C, 64), row_blocks), 256, 0, st>>>(in_hi, in_plane, R, C, ld_in, out_hi, out_plane, ld_out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void split_rows_kernel(const float* __restrict__ in, long long ld_in, int C, long long total4,
                                  __half* __restrict__ out, long long out_plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / c4n;
    const int c = (int)(i % c4n) * 4;
    store_split4(out + r * C + c, out_plane, *reinterpret_cast<const float4*>(in + r * ld_in + c));
  }
}
int split_rows_f32(const float* in, long long ld_in, int R, int C, __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(C % 4 == 0 && ld_in % 4 == 0, "split_rows_f32: C and ld must be multiples of 4");
  const long long total4 = (long long)R * C / 4;
  split_rows_kernel<<<grid_for(total4, 256), 256, 0, st>>>(in, ld_in, C, total4, out_hi, out_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------ column sums
template <bool kPlanes>
__global__ void colsum_stage1_kernel(const float* __restrict__ in_f, const __half* __restrict__ in_h, long long plane,
                                     long long ld, int R, int C, float* __restrict__ partial) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y;
  const int per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = chunk * per, r1 = min(R, r0 + per);
  if (c >= C) return;
  float s = 0.f;
  for (int r = r0; r < r1; ++r) {
    if (kPlanes) s += __half2float(in_h[(long long)r * ld + c]) + __half2float(in_h[(long long)r * ld + c + plane]);
    else s += in_f[(long long)r * ld + c];
  }
  partial[(long long)chunk * C + c] = s;
}
// Vectorised stage 1 (round 2): a thread owns 4 adjacent columns (one 16-byte load per row; planes: 8 bytes hi + 8 bytes lo)
// and walks its row chunk four rows at a time with independent accumulators, so every thread keeps 4 loads in flight (the
// scalar version ran the 310 MB fc1 bias gradient at 1.6 TB/s).  Fixed summation order: deterministic like the scalar kernel.
template <bool kPlanes>
__global__ void colsum_stage1_vec4_kernel(const float* __restrict__ in_f, const __half* __restrict__ in_h, long long plane,
                                          long long ld, int R, int C, float* __restrict__ partial) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int chunk = blockIdx.y;
  const int per = (R + gridDim.y - 1) / gridDim.y;
  const int r0 = chunk * per, r1 = min(R, r0 + per);
  if (c >= C) return;
  float4 acc[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto load = [&](int r) -> float4 {
    if (kPlanes) {
      const uint2 h = *reinterpret_cast<const uint2*>(in_h + (long long)r * ld + c);
      const uint2 l = *reinterpret_cast<const uint2*>(in_h + (long long)r * ld + c + plane);
      const float2 h01 = __half22float2(*reinterpret_cast<const __half2*>(&h.x)), h23 = __half22float2(*reinterpret_cast<const __half2*>(&h.y));
      const float2 l01 = __half22float2(*reinterpret_cast<const __half2*>(&l.x)), l23 = __half22float2(*reinterpret_cast<const __half2*>(&l.y));
      return make_float4(h01.x + l01.x, h01.y + l01.y, h23.x + l23.x, h23.y + l23.y);
    }
    return *reinterpret_cast<const float4*>(in_f + (long long)r * ld + c);
  };
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = load(r + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
  }
  for (; r < r1; ++r) { const float4 v = load(r); acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w; }
  float4 s;
  s.x = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
  s.y = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
  s.z = (acc[0].z + acc[1].z) + (acc[2].z + acc[3].z);
  s.w = (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w);
  *reinterpret_cast<float4*>(partial + (long long)chunk * C + c) = s;
}
// stage 2: 32 columns per block, 8 threads per column walk the chunk partials k = g, g + 8, ... (independent loads in flight),
// then a fixed-order sum of the 8 group sums: deterministic, and ~4 us instead of 14 us for 256 chunks (192 launches per step).
__global__ void __launch_bounds__(256)
colsum_stage2_kernel(const float* __restrict__ partial, int chunks, int C, float scale, int accumulate,
                     float* __restrict__ out) {
  __shared__ float s_g[8][33];
  const int cl = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < C)
    for (int k = g; k < chunks; k += 8) s += partial[(long long)k * C + c];
  s_g[g][cl] = s;
  __syncthreads();
  if (g == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += s_g[k][cl];
    out[c] = (accumulate ? out[c] : 0.f) + scale * t;
  }
}
static int colsum_impl(const float* in_f, const __half* in_h, long long plane, long long ld, int R, int C, float scale,
                       int accumulate, float* scratch, float* out, cudaStream_t st) {
  MAED_CHECK_ARG(R >= 1 && C >= 1 && scratch && out, "colsum: bad arguments R=%d C=%d", R, C);
  int chunks = R < kColsumChunks ? R : kColsumChunks;
  const bool vec4 = (C % 4 == 0) && (ld % 4 == 0) && (reinterpret_cast<uintptr_t>(scratch) % 16 == 0) &&
                    (in_h ? (reinterpret_cast<uintptr_t>(in_h) % 8 == 0 && plane % 4 == 0) : (reinterpret_cast<uintptr_t>(in_f) % 16 == 0));
  if (vec4) {
    const int threads = C >= 256 ? 64 : 32;                // 4 columns per thread; narrow blocks keep >= 148 CTAs for C = 768
    const dim3 vgrid(cdiv(C, 4 * threads), chunks);
    if (in_h) colsum_stage1_vec4_kernel<true><<<vgrid, threads, 0, st>>>(nullptr, in_h, plane, ld, R, C, scratch);
    else colsum_stage1_vec4_kernel<false><<<vgrid, threads, 0, st>>>(in_f, nullptr, 0, ld, R, C, scratch);
    MAED_BW_LAUNCH_CHECK();
    colsum_stage2_kernel<<<cdiv(C, 32), 256, 0, st>>>(scratch, chunks, C, scale, accumulate, out);
    MAED_BW_LAUNCH_CHECK();
    return MAED_OK;
  }
  const dim3 grid(cdiv(C, 128), chunks);
  if (in_h) colsum_stage1_kernel<true><<<grid, 128, 0, st>>>(nullptr, in_h, plane, ld, R, C, scratch);
  else colsum_stage1_kernel<false><<<grid, 128, 0, st>>>(in_f, nullptr, 0, ld, R, C, scratch);
  MAED_BW_LAUNCH_CHECK();
  colsum_stage2_kernel<<<cdiv(C, 32), 256, 0, st>>>(scratch, chunks, C, scale, accumulate, out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
int colsum_f32(const float* in, long long ld, int R, int C, float scale, int accumulate, float* scratch, float* out,
               cudaStream_t st) {
  return colsum_impl(in, nullptr, 0, ld, R, C, scale, accumulate, scratch, out, st);
}
int colsum_planes(const __half* in_hi, long long plane, long long ld, int R, int C, float scale, int accumulate,
                  float* scratch, float* out, cudaStream_t st) {
  return colsum_impl(nullptr, in_hi, plane, ld, R, C, scale, accumulate, scratch, out, st);
}

__global__ void add_f32_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] += b[i];
}
int add_f32(float* a, const float* b, long long n, cudaStream_t st) {
  add_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(a, b, n);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
__global__ void scale_f32_kernel(const float* __restrict__ a, float s, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = a[i] * s;
}
int scale_f32(const float* a, float s, long long n, float* out, cudaStream_t st) {
  scale_f32_kernel<<<grid_for(n, 256), 256, 0, st>>>(a, s, n, out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------------------- elementwise backward
// reverse: walk from the end of the tensor (the part its producer wrote last is still in L2)
__global__ void relu_mask_kernel(float* __restrict__ d, const __half* __restrict__ act, long long n4, int reverse) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n4; j += (long long)gridDim.x * blockDim.x) {
    const long long i = reverse ? n4 - 1 - j : j;
    float4 v = reinterpret_cast<float4*>(d)[i];
    const uint2 a = reinterpret_cast<const uint2*>(act)[i];
    const float2 a01 = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
    const float2 a23 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    if (!(a01.x > 0.f)) v.x = 0.f;
    if (!(a01.y > 0.f)) v.y = 0.f;
    if (!(a23.x > 0.f)) v.z = 0.f;
    if (!(a23.y > 0.f)) v.w = 0.f;
    reinterpret_cast<float4*>(d)[i] = v;
  }
}
int relu_mask_f32(float* d, const __half* act_hi, long long n, cudaStream_t st, int reverse) {
  MAED_CHECK_ARG(n % 4 == 0, "relu_mask_f32: n must be a multiple of 4");
  relu_mask_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(d, act_hi, n / 4, reverse);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
  return cdf + x * pdf;
}
__global__ void gelu_bwd_kernel(const float* __restrict__ d_hid, const float* __restrict__ pre, long long n4,
                                __half* __restrict__ out, long long out_plane) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 d = reinterpret_cast<const float4*>(d_hid)[i];
    const float4 x = reinterpret_cast<const float4*>(pre)[i];
    store_split4(out + 4 * i, out_plane,
                 make_float4(d.x * gelu_grad_f(x.x), d.y * gelu_grad_f(x.y), d.z * gelu_grad_f(x.z), d.w * gelu_grad_f(x.w)));
  }
}
int gelu_bwd(const float* d_hid, const float* pre, long long n, __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(n % 4 == 0, "gelu_bwd: n must be a multiple of 4");
  gelu_bwd_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(d_hid, pre, n / 4, out_hi, out_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
__global__ void gelu_fwd_kernel(const float* __restrict__ pre, long long n4, __half* __restrict__ out, long long out_plane) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(pre)[i];
    store_split4(out + 4 * i, out_plane, make_float4(gelu_f(x.x), gelu_f(x.y), gelu_f(x.z), gelu_f(x.w)));
  }
}
int gelu_fwd_planes(const float* pre, long long n, __half* out_hi, long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(n % 4 == 0, "gelu_fwd_planes: n must be a multiple of 4");
  gelu_fwd_kernel<<<grid_for(n / 4, 256), 256, 0, st>>>(pre, n / 4, out_hi, out_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dy[i] * (1.0f - y[i] * y[i]);
}
int tanh_bwd(const float* d_y, const float* y, long long n, float* d_pre, cudaStream_t st) {
  tanh_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(d_y, y, n, d_pre);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// counter-based RNG: splitmix64 of (seed, index) -> uniform in [0,1)
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}
__global__ void dropout_fwd_kernel(float* __restrict__ x, long long n, float p, float inv_keep, unsigned long long seed,
                                   unsigned char* __restrict__ mask) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned char keep = hash_uniform(seed, (unsigned long long)i) >= p ? 1 : 0;
    mask[i] = keep;
    x[i] = keep ? x[i] * inv_keep : 0.f;
  }
}
int dropout_fwd(float* x, long long n, float p, unsigned long long seed, unsigned char* mask, cudaStream_t st) {
  MAED_CHECK_ARG(p >= 0.f && p < 1.f, "dropout_fwd: p=%f", (double)p);
  dropout_fwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(x, n, p, 1.0f / (1.0f - p), seed, mask);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
__global__ void dropout_bwd_kernel(float* __restrict__ d, long long n, float inv_keep, const unsigned char* __restrict__ mask) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    d[i] = mask[i] ? d[i] * inv_keep : 0.f;
}
int dropout_bwd(float* d, long long n, float p, const unsigned char* mask, cudaStream_t st) {
  dropout_bwd_kernel<<<grid_for(n, 256), 256, 0, st>>>(d, n, 1.0f / (1.0f - p), mask);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------ LayerNorm backward
// One warp walks rows wg, wg + W, ...; statistics recomputed from x (two-pass, like the forward); dgamma / dbeta
// accumulate in registers and are written once per warp to partial[wg][2C].
static constexpr int kLnBwdBlocks = 148 * 2;
static constexpr int kLnBwdWarps = 8;
int ln_bwd_partial_rows() { return kLnBwdBlocks * kLnBwdWarps; }

// One warp per row, NV float4 per lane (C = 128 NV).  Round 2: templated on NV (C = 768 used arrays sized for 1024), x and dy of
// a row are both requested before the first reduction (12 independent 16-byte loads per lane in flight), gamma is re-read
// through L1 instead of living in registers, and two 256-thread blocks fit an SM (the first version ran at 20 % of HBM peak).
template <int NV>
__global__ void __launch_bounds__(kLnBwdWarps * 32, 2)
layernorm_bwd_kernel(const float* __restrict__ dy, long long dy_stride, const float* __restrict__ x, long long x_stride,
                     const float* __restrict__ gamma, int rows, float eps, const float* __restrict__ dx_add,
                     float* __restrict__ dx_out, long long dx_stride, float* __restrict__ partial) {
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * kLnBwdWarps + (threadIdx.x >> 5);
  const int W = gridDim.x * kLnBwdWarps;
  float4 dgam[NV], dbet[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    dgam[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    dbet[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = wg; row < rows; row += W) {
    const float* xr = x + (long long)row * x_stride;
    const float* dr = dy + (long long)row * dy_stride;
    float4 v[NV], g[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = *reinterpret_cast<const float4*>(xr + (j * 32 + lane) * 4);
#pragma unroll
    for (int j = 0; j < NV; ++j) g[j] = *reinterpret_cast<const float4*>(dr + (j * 32 + lane) * 4);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) s += v[j].x + v[j].y + v[j].z + v[j].w;
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
      q += v[j].x * v[j].x + v[j].y * v[j].y + v[j].z * v[j].z + v[j].w * v[j].w;
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
    float s1 = 0.f, s2 = 0.f;                              // sum(dy*gamma), sum(dy*gamma*xhat)
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 d = g[j];
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + (j * 32 + lane) * 4));
      v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;          // xhat
      dgam[j].x += d.x * v[j].x; dgam[j].y += d.y * v[j].y; dgam[j].z += d.z * v[j].z; dgam[j].w += d.w * v[j].w;
      dbet[j].x += d.x; dbet[j].y += d.y; dbet[j].z += d.z; dbet[j].w += d.w;
      g[j] = make_float4(d.x * gm.x, d.y * gm.y, d.z * gm.z, d.w * gm.w);
      s1 += g[j].x + g[j].y + g[j].z + g[j].w;
      s2 += g[j].x * v[j].x + g[j].y * v[j].y + g[j].z * v[j].z + g[j].w * v[j].w;
    }
    const float m1 = warp_sum(s1) / (float)C, m2 = warp_sum(s2) / (float)C;
    float* o = dx_out + (long long)row * dx_stride;
    const float* a = dx_add ? dx_add + (long long)row * dx_stride : nullptr;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 r;
      r.x = rstd * (g[j].x - m1 - v[j].x * m2);
      r.y = rstd * (g[j].y - m1 - v[j].y * m2);
      r.z = rstd * (g[j].z - m1 - v[j].z * m2);
      r.w = rstd * (g[j].w - m1 - v[j].w * m2);
      if (a) {
        const float4 b = *reinterpret_cast<const float4*>(a + (j * 32 + lane) * 4);
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      *reinterpret_cast<float4*>(o + (j * 32 + lane) * 4) = r;
    }
  }
  float* pr = partial + (long long)wg * 2 * C;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    *reinterpret_cast<float4*>(pr + (j * 32 + lane) * 4) = dgam[j];
    *reinterpret_cast<float4*>(pr + C + (j * 32 + lane) * 4) = dbet[j];
  }
}
int layernorm_bwd(const float* dy, long long dy_stride, const float* x, long long x_stride, const float* gamma, int rows,
                  int C, float eps, const float* dx_add, float* dx_out, long long dx_stride, float* partial,
                  cudaStream_t st) {
  MAED_CHECK_ARG(C % 128 == 0 && C <= 1024, "layernorm_bwd: C=%d unsupported (multiple of 128, <= 1024)", C);
#define MAED_LN_BWD(NV)                                                                                                       \
  case NV:                                                                                                                    \
    layernorm_bwd_kernel<NV><<<kLnBwdBlocks, kLnBwdWarps * 32, 0, st>>>(dy, dy_stride, x, x_stride, gamma, rows, eps, dx_add, \
                                                                         dx_out, dx_stride, partial);                         \
    break;
  switch (C / 128) {
    MAED_LN_BWD(1) MAED_LN_BWD(2) MAED_LN_BWD(3) MAED_LN_BWD(4) MAED_LN_BWD(5) MAED_LN_BWD(6) MAED_LN_BWD(7) MAED_LN_BWD(8)
  }
#undef MAED_LN_BWD
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------ GroupNorm backward
__device__ __forceinline__ void gn_group_stats(const double* stats, int n, int HW, int C, float eps, float* s_mean,
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
// stage 1: per (image, hw-chunk, channel): A = sum dy*xhat, B = sum dy.   grid (chunks, n_img), 256 threads.
// relu_gb (gamma, beta of the layer) != nullptr: the GroupNorm was followed by a ReLU and dy is the gradient w.r.t. the ReLU's
// OUTPUT; the mask (xhat * gamma + beta > 0, the forward's own expression) is applied on the fly instead of by a separate pass
// over dy.  reverse: images in descending order (see groupnorm_bwd).
__global__ void gn_bwd_stage1_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const double* __restrict__ stats, int HW, int C, float eps, float* __restrict__ part,
                                     const float* __restrict__ relu_g, const float* __restrict__ relu_b, int reverse) {
  __shared__ float s_mean[32], s_rstd[32];
  __shared__ float4 s_a[256], s_b[256];
  const int n = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
  gn_group_stats(stats, n, HW, C, eps, s_mean, s_rstd);
  const int gsz = C / 32, c4n = C >> 2;
  const int per = (HW + chunks - 1) / chunks;
  const int hw0 = chunk * per, hw1 = min(HW, hw0 + per);
  const long long base = (long long)n * HW * C;
  float* po = part + ((long long)n * chunks + chunk) * 2 * C;
  // a thread owns 4 neighbouring channels (16-byte loads of dy and x, both issued before either is used); with C / 4 <= 256
  // threads per row the block walks 256 / (C / 4) rows per step, wider maps loop over their channel quads
  const bool tiled = c4n <= 256 && 256 % c4n == 0;
  const int rstep = tiled ? 256 / c4n : 1;
  for (int c4 = tiled ? threadIdx.x % c4n : threadIdx.x; c4 < c4n; c4 += 256) {
    const int c = c4 * 4, rr = tiled ? threadIdx.x / c4n : 0;
    float m[4], r[4], mg[4], mb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      m[e] = s_mean[(c + e) / gsz]; r[e] = s_rstd[(c + e) / gsz];
      mg[e] = relu_g ? relu_g[c + e] : 0.f; mb[e] = relu_g ? relu_b[c + e] : 1.f;     // no ReLU: xhat * 0 + 1 > 0 always
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int hw = hw0 + rr; hw < hw1; hw += rstep) {
      const long long off = base + (long long)hw * C + c;
      const float4 xv = *reinterpret_cast<const float4*>(x + off);
      const float4 dv = *reinterpret_cast<const float4*>(dy + off);
      const float xx[4] = {xv.x, xv.y, xv.z, xv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xhat = (xx[e] - m[e]) * r[e];
        const float d = (xhat * mg[e] + mb[e] > 0.f) ? dd[e] : 0.f;
        a[e] += d * xhat;
        b[e] += d;
      }
    }
    float4 av = make_float4(a[0], a[1], a[2], a[3]), bv = make_float4(b[0], b[1], b[2], b[3]);
    if (tiled && rstep > 1) {
      s_a[threadIdx.x] = av;
      s_b[threadIdx.x] = bv;
      __syncthreads();                                     // uniform: tiled blocks run the c4 loop exactly once
      if (threadIdx.x >= c4n) break;
      for (int k = 1; k < rstep; ++k) {
        const float4 pa = s_a[threadIdx.x + k * c4n], pb = s_b[threadIdx.x + k * c4n];
        av.x += pa.x; av.y += pa.y; av.z += pa.z; av.w += pa.w;
        bv.x += pb.x; bv.y += pb.y; bv.z += pb.z; bv.w += pb.w;
      }
    }
    *reinterpret_cast<float4*>(po + c) = av;
    *reinterpret_cast<float4*>(po + C + c) = bv;
  }
}
// stage 2: sum the chunks -> dgb_partial[n][2][C]; group sums red[n][g] = (sum_c gamma*B, sum_c gamma*A).  grid n_img.
__global__ void gn_bwd_stage2_kernel(const float* __restrict__ part, int chunks, const float* __restrict__ gamma, int C,
                                     float* __restrict__ dgb_partial, float* __restrict__ red) {
  extern __shared__ float sm[];                            // [2][C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < chunks; ++k) s += part[((long long)n * chunks + k) * 2 * C + c];
    dgb_partial[(long long)n * 2 * C + c] = s;
    sm[c] = s * gamma[c % C];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int gsz = C / 32, g = threadIdx.x;
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < gsz; ++k) { s2 += sm[g * gsz + k]; s1 += sm[C + g * gsz + k]; }
    red[((long long)n * 32 + g) * 2] = s1;                 // sum dy*gamma
    red[((long long)n * 32 + g) * 2 + 1] = s2;             // sum dy*gamma*xhat
  }
}
// stage 3: dx -> planes.  grid (blocks_per_img, n_img)
__global__ void gn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                    const double* __restrict__ stats, const float* __restrict__ gamma,
                                    const float* __restrict__ red, int HW, int C, float eps, __half* __restrict__ out,
                                    long long out_plane, const float* __restrict__ relu_b, int reverse) {
  __shared__ float s_mean[32], s_rstd[32], s_m1[32], s_m2[32];
  const int n = reverse ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  gn_group_stats(stats, n, HW, C, eps, s_mean, s_rstd);
  if (threadIdx.x < 32) {
    const float cnt = (float)HW * (float)(C / 32);
    s_m1[threadIdx.x] = red[((long long)n * 32 + threadIdx.x) * 2] / cnt;
    s_m2[threadIdx.x] = red[((long long)n * 32 + threadIdx.x) * 2 + 1] / cnt;
  }
  __syncthreads();
  const int c4n = C >> 2, gsz = C / 32;
  const long long total = (long long)HW * c4n;
  const long long base = (long long)n * HW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long off = base + (i / c4n) * C + c;
    const float4 d = *reinterpret_cast<const float4*>(dy + off);
    const float4 v = *reinterpret_cast<const float4*>(x + off);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b4 = relu_b ? *reinterpret_cast<const float4*>(relu_b + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float dd[4] = {d.x, d.y, d.z, d.w}, vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w};
    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float r[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int grp = (c + e) / gsz;
      const float xhat = (vv[e] - s_mean[grp]) * s_rstd[grp];
      const float de = (!relu_b || xhat * gg[e] + bb[e] > 0.f) ? dd[e] : 0.f;
      r[e] = s_rstd[grp] * (de * gg[e] - s_m1[grp] - xhat * s_m2[grp]);
    }
    store_split4(out + off, out_plane, make_float4(r[0], r[1], r[2], r[3]));
  }
}
// relu_beta != nullptr: dy is the gradient behind the ReLU that followed this GroupNorm (mask recomputed, see stage 1).
// order 0: both passes walk the images upwards; 1: the producer of dy wrote it upwards, so the statistics pass walks DOWN (the
// last images are still in L2) and the apply pass UP again (the first images were read last); 2: the mirror image of 1.
int groupnorm_bwd(const float* dy, const float* x, const double* stats, const float* gamma, int n_img, int HW, int C,
                  float eps, float* red, float* dgb_partial, __half* dx_hi, long long dx_plane, cudaStream_t st,
                  const float* relu_beta, int order) {
  MAED_CHECK_ARG(C % 32 == 0 && C <= 4096, "groupnorm_bwd: C=%d unsupported", C);
#ifndef MAED_EMU
  {                                                        // one cluster per image (gn_cluster.cu): dy and x leave HBM once
    const int rc = groupnorm_bwd_cluster(dy, x, stats, gamma, relu_beta, n_img, HW, C, eps, dgb_partial, dx_hi, dx_plane, st);
    if (rc != MAED_ERR_UNSUPPORTED) return rc;
  }
#endif
  // the stage-1 partials live at the tail of `red`'s scratch: caller provides red with n_img*(64 + 16*2*C) floats
  int chunks = cdiv(8LL * sm_count(), n_img);             // several waves of blocks: no tail imbalance at 2-3 blocks per SM
  if (chunks > 16) chunks = 16;
  if (chunks < 1) chunks = 1;
  if (chunks > HW) chunks = HW;
  float* part = red + (long long)n_img * 64;
  gn_bwd_stage1_kernel<<<dim3(chunks, n_img), 256, 0, st>>>(dy, x, stats, HW, C, eps, part, relu_beta ? gamma : nullptr, relu_beta,
                                                            order == 1);
  MAED_BW_LAUNCH_CHECK();
  gn_bwd_stage2_kernel<<<n_img, 256, 2 * C * sizeof(float), st>>>(part, chunks, gamma, C, dgb_partial, red);
  MAED_BW_LAUNCH_CHECK();
  const long long per_img = (long long)HW * C / 4;
  int bpi = cdiv((long long)sm_count() * 8, n_img);
  const int maxb = cdiv(per_img, 256);
  if (bpi > maxb) bpi = maxb;
  if (bpi < 1) bpi = 1;
  gn_bwd_apply_kernel<<<dim3(bpi, n_img), 256, 0, st>>>(dy, x, stats, gamma, red, HW, C, eps, dx_hi, dx_plane, relu_beta,
                                                        order == 2);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------- weight standardisation backward
// One block per output channel; F = Cin*KH*KW <= 9216.
__global__ void wstd_bwd_kernel(const float* __restrict__ g, int k_pad, const float* __restrict__ w, int Cin, int KH, int KW,
                                float eps, float scale, float* __restrict__ dw) {
  __shared__ float buf[32];
  const int co = blockIdx.x;
  const int F = Cin * KH * KW, taps = KH * KW;
  const float* wr = w + (long long)co * F;
  const float* gr = g + (long long)co * k_pad;
  float s = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) s += wr[i];
  const float mean = block_sum(s, buf) / (float)F;
  float q = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) { const float d = wr[i] - mean; q += d * d; }
  const float stdv = sqrtf(block_sum(q, buf) / (float)F);
  const float inv = 1.0f / (stdv + eps);
  float sg = 0.f, sgw = 0.f;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {      // i = OIHW index (ci, kh, kw)
    const int ci = i / taps, tap = i % taps;
    const float gv = gr[tap * Cin + ci];
    sg += gv;
    sgw += gv * (wr[i] - mean) * inv;
  }
  const float mg = block_sum(sg, buf) / (float)F;
  const float mgw = block_sum(sgw, buf) / (float)F;
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const int ci = i / taps, tap = i % taps;
    const float gv = gr[tap * Cin + ci];
    const float what = (wr[i] - mean) * inv;
    dw[(long long)co * F + i] = scale * ((gv - mg) * inv - what * mgw / stdv);
  }
}
int wstd_bwd(const float* g, int k_pad, const float* w, int Cout, int Cin, int KH, int KW, float eps, float scale,
             float* dw, cudaStream_t st) {
  wstd_bwd_kernel<<<Cout, 256, 0, st>>>(g, k_pad, w, Cin, KH, KW, eps, scale, dw);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------- derived weights for dgrad
__global__ void split_transposed_kernel(const float* __restrict__ w, int N, int K, __half* __restrict__ out, long long plane) {
  __shared__ float t[32][33];
  const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty + 8 * j, k = k0 + tx;
    t[ty + 8 * j][tx] = (n < N && k < K) ? w[(long long)n * K + k] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = k0 + ty + 8 * j, n = n0 + tx;
    if (k < K && n < N) {
      const float v = t[tx][ty + 8 * j];
      const __half h = __float2half_rn(v);
      out[(long long)k * N + n] = h;
      out[(long long)k * N + n + plane] = __float2half_rn(v - __half2float(h));
    }
  }
}
int split_f32_transposed(const float* w, int N, int K, __half* out_hi, long long plane, cudaStream_t st) {
  split_transposed_kernel<<<dim3(cdiv(K, 32), cdiv(N, 32)), 256, 0, st>>>(w, N, K, out_hi, plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void prep_conv_weight_dgrad_kernel(const float* __restrict__ w, int Cout, int Cin, int KH, int KW, int standardize,
                                              __half* __restrict__ out, long long plane) {
  __shared__ float buf[32];
  const int co = blockIdx.x;
  const int F = Cin * KH * KW, taps = KH * KW;
  const float* wr = w + (long long)co * F;
  float mean = 0.f, inv = 1.f;
  if (standardize) {
    float s = 0.f;
    for (int i = threadIdx.x; i < F; i += blockDim.x) s += wr[i];
    mean = block_sum(s, buf) / (float)F;
    float q = 0.f;
    for (int i = threadIdx.x; i < F; i += blockDim.x) { const float d = wr[i] - mean; q += d * d; }
    inv = 1.0f / (sqrtf(block_sum(q, buf) / (float)F) + 1e-5f);
  }
  for (int i = threadIdx.x; i < F; i += blockDim.x) {
    const int ci = i / taps, tap = i % taps;
    const int kh = tap / KW, kw = tap % KW;
    const int ftap = (KH - 1 - kh) * KW + (KW - 1 - kw);
    const float v = (wr[i] - mean) * inv;
    const long long o = ((long long)ci * taps + ftap) * Cout + co;
    const __half h = __float2half_rn(v);
    out[o] = h;
    out[o + plane] = __float2half_rn(v - __half2float(h));
  }
}
int prep_conv_weight_dgrad(const float* w, int Cout, int Cin, int KH, int KW, int standardize, __half* out_hi,
                           long long plane, cudaStream_t st) {
  prep_conv_weight_dgrad_kernel<<<Cout, 256, 0, st>>>(w, Cout, Cin, KH, KW, standardize, out_hi, plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ---------------------------------------------------------------------------- stem max-pool (training)
__global__ void gn_apply_maxpool_idx_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                            const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W,
                                            int C, float eps, int OH, int OW, int pad_t, int pad_l, __half* __restrict__ out,
                                            long long out_plane, unsigned char* __restrict__ idx) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.y;
  gn_group_stats(stats, n, H * W, C, eps, s_mean, s_rstd);
  const int c4n = C >> 2, gsz = C / 32;
  const long long total = (long long)OH * OW * c4n;
  const float* xb = x + (long long)n * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int ow = (int)((i / c4n) % OW);
    const int oh = (int)(i / ((long long)c4n * OW));
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned char am[4] = {0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - pad_t;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - pad_l;
        if (iw < 0 || iw >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(xb + ((long long)ih * W + iw) * C + c);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int grp = (c + e) / gsz;
          const float y = (vv[e] - s_mean[grp]) * s_rstd[grp] * gg[e] + bb[e];
          if (y > m[e]) { m[e] = y; am[e] = (unsigned char)(r * 3 + s); }      // first maximum in scan order
        }
      }
    }
    const long long off = (((long long)n * OH + oh) * OW + ow) * C + c;
    store_split4(out + off, out_plane, make_float4(fmaxf(m[0], 0.f), fmaxf(m[1], 0.f), fmaxf(m[2], 0.f), fmaxf(m[3], 0.f)));
    *reinterpret_cast<uchar4*>(idx + off) = make_uchar4(am[0], am[1], am[2], am[3]);
  }
}
int gn_apply_maxpool_idx(const float* x, const double* stats, const float* gamma, const float* beta, int n_img, int H, int W,
                         int C, float eps, __half* out_hi, long long out_plane, unsigned char* idx, cudaStream_t st) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const int pad_h = max((OH - 1) * 2 + 3 - H, 0), pad_w = max((OW - 1) * 2 + 3 - W, 0);
  const long long per_img = (long long)OH * OW * C / 4;
  int bpi = cdiv((long long)sm_count() * 8, n_img);
  const int maxb = cdiv(per_img, 256);
  if (bpi > maxb) bpi = maxb;
  if (bpi < 1) bpi = 1;
  gn_apply_maxpool_idx_kernel<<<dim3(bpi, n_img), 256, 0, st>>>(x, stats, gamma, beta, H, W, C, eps, OH, OW, pad_h / 2,
                                                                pad_w / 2, out_hi, out_plane, idx);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// gather form: every input position sums the pooled gradients of the (at most 2x2) windows whose arg-max it is
__global__ void maxpool_gn_relu_bwd_kernel(const float* __restrict__ d_pool, const unsigned char* __restrict__ idx,
                                           const float* __restrict__ x, const double* __restrict__ stats,
                                           const float* __restrict__ gamma, const float* __restrict__ beta, int H, int W,
                                           int C, float eps, int OH, int OW, int pad_t, int pad_l, float* __restrict__ d_y) {
  __shared__ float s_mean[32], s_rstd[32];
  const int n = blockIdx.y;
  gn_group_stats(stats, n, H * W, C, eps, s_mean, s_rstd);
  const int c4n = C >> 2, gsz = C / 32;
  const long long total = (long long)H * W * c4n;
  const long long xb = (long long)n * H * W * C, pb = (long long)n * OH * OW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int w = (int)((i / c4n) % W);
    const int h = (int)(i / ((long long)c4n * W));
    const long long off = xb + ((long long)h * W + w) * C + c;
    const float4 v = *reinterpret_cast<const float4*>(x + off);
    const float4 g = *reinterpret_cast<const float4*>(gamma + c);
    const float4 b = *reinterpret_cast<const float4*>(beta + c);
    const float vv[4] = {v.x, v.y, v.z, v.w}, gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {b.x, b.y, b.z, b.w};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < 3; ++r) {
      const int t = h + pad_t - r;
      if (t < 0 || (t & 1)) continue;
      const int oh = t >> 1;
      if (oh >= OH) continue;
      for (int s = 0; s < 3; ++s) {
        const int u = w + pad_l - s;
        if (u < 0 || (u & 1)) continue;
        const int ow = u >> 1;
        if (ow >= OW) continue;
        const long long po = pb + ((long long)oh * OW + ow) * C + c;
        const uchar4 am = *reinterpret_cast<const uchar4*>(idx + po);
        const float4 dp = *reinterpret_cast<const float4*>(d_pool + po);
        const unsigned char tap = (unsigned char)(r * 3 + s);
        if (am.x == tap) acc[0] += dp.x;
        if (am.y == tap) acc[1] += dp.y;
        if (am.z == tap) acc[2] += dp.z;
        if (am.w == tap) acc[3] += dp.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int grp = (c + e) / gsz;
      const float y = (vv[e] - s_mean[grp]) * s_rstd[grp] * gg[e] + bb[e];
      if (!(y > 0.f)) acc[e] = 0.f;                                             // ReLU
    }
    *reinterpret_cast<float4*>(d_y + off) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
}
int maxpool_gn_relu_bwd(const float* d_pool, const unsigned char* idx, const float* x, const double* stats,
                        const float* gamma, const float* beta, int n_img, int H, int W, int C, float eps, float* d_y,
                        cudaStream_t st) {
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const int pad_h = max((OH - 1) * 2 + 3 - H, 0), pad_w = max((OW - 1) * 2 + 3 - W, 0);
  const long long per_img = (long long)H * W * C / 4;
  int bpi = cdiv((long long)sm_count() * 8, n_img);
  const int maxb = cdiv(per_img, 256);
  if (bpi > maxb) bpi = maxb;
  if (bpi < 1) bpi = 1;
  maxpool_gn_relu_bwd_kernel<<<dim3(bpi, n_img), 256, 0, st>>>(d_pool, idx, x, stats, gamma, beta, H, W, C, eps, OH, OW,
                                                               pad_h / 2, pad_w / 2, d_y);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ----------------------------------------------------------------------------------- stride-2 helpers
__global__ void dilate2_kernel(const __half* __restrict__ in, long long in_plane, int OH, int OW, int C, int H, int W,
                               long long total4, __half* __restrict__ out, long long out_plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int w = (int)((i / c4n) % W);
    const int h = (int)((i / ((long long)c4n * W)) % H);
    const long long n = i / ((long long)c4n * W * H);
    uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
    if (!(h & 1) && !(w & 1) && (h >> 1) < OH && (w >> 1) < OW) {
      const long long src = ((n * OH + (h >> 1)) * OW + (w >> 1)) * C + c;
      hi = *reinterpret_cast<const uint2*>(in + src);
      lo = *reinterpret_cast<const uint2*>(in + src + in_plane);
    }
    const long long dst = ((n * H + h) * W + w) * C + c;
    *reinterpret_cast<uint2*>(out + dst) = hi;
    *reinterpret_cast<uint2*>(out + dst + out_plane) = lo;
  }
}
int dilate2_planes(const __half* in_hi, long long in_plane, int n_img, int OH, int OW, int C, int H, int W, __half* out_hi,
                   long long out_plane, cudaStream_t st) {
  MAED_CHECK_ARG(C % 4 == 0, "dilate2_planes: C must be a multiple of 4");
  const long long total4 = (long long)n_img * H * W * C / 4;
  dilate2_kernel<<<grid_for(total4, 256), 256, 0, st>>>(in_hi, in_plane, OH, OW, C, H, W, total4, out_hi, out_plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
__global__ void scatter_stride2_kernel(const float* __restrict__ src, int OH, int OW, int C, int H, int W, long long total4,
                                       const float* __restrict__ add, float* __restrict__ d_in) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const int w = (int)((i / c4n) % W);
    const int h = (int)((i / ((long long)c4n * W)) % H);
    const long long n = i / ((long long)c4n * W * H);
    float4 v = add ? reinterpret_cast<const float4*>(add)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    if (!(h & 1) && !(w & 1) && (h >> 1) < OH && (w >> 1) < OW) {
      const float4 s = *reinterpret_cast<const float4*>(src + ((n * OH + (h >> 1)) * OW + (w >> 1)) * C + c);
      v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
    }
    reinterpret_cast<float4*>(d_in)[i] = v;
  }
}
int scatter_stride2_f32(const float* src, int n_img, int OH, int OW, int C, int H, int W, const float* add, float* d_in,
                        cudaStream_t st) {
  MAED_CHECK_ARG(C % 4 == 0, "scatter_stride2_f32: C must be a multiple of 4");
  const long long total4 = (long long)n_img * H * W * C / 4;
  scatter_stride2_kernel<<<grid_for(total4, 256), 256, 0, st>>>(src, OH, OW, C, H, W, total4, add, d_in);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
