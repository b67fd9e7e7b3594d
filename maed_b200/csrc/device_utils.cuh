// Small device / launch helpers shared by the training-path translation units (bwd_kernels.cu, attention_bwd.cu,
// train.cu).  The forward-path files keep their own file-local copies.
#pragma once
#include "common.h"
#include "kernels.h"

namespace maed {
namespace bw {

#define MAED_BW_LAUNCH_CHECK()              \
  do {                                      \
    ::maed::count_launch();                 \
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
__device__ __forceinline__ float load_plane1(const __half* hi_ptr, long long plane) {
  return __half2float(hi_ptr[0]) + (plane ? __half2float(hi_ptr[plane]) : 0.f);
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over a block of up to 1024 threads; every thread gets the result.  `buf` holds 32 floats.
__device__ __forceinline__ float block_sum(float v, float* buf) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) buf[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? buf[threadIdx.x] : 0.f;
  if (warp == 0) {
    t = warp_sum(t);
    if (lane == 0) buf[0] = t;
  }
  __syncthreads();
  return buf[0];
}

static inline int grid_for(long long work_items, int threads, int max_waves = 8) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * max_waves;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace bw
}  // namespace maed
