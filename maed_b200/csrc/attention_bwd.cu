// Backward of the STE attention (reference vision_transformer.py:206-228 under autograd).  First version: CUDA-core fp32
// kernels that recompute the probabilities from the saved q/k/v planes (flash-attention style, nothing but qkv is
// kept from the forward):
//     S = scale Q K^T,  P = softmax(S),  dV = P^T dO,  dP = dO V^T,  D_i = sum_j P_ij dP_ij,
//     dS = scale P o (dP - D),  dQ = dS K,  dK = dS^T Q.
//  * attn_spatial_bwd_kernel  — one CTA per (frame, head): K and V of the frame live in shared memory, queries are walked
//    in blocks of 32 rows, each thread owns one key row of the dK / dV accumulators in registers.
//  * attn_temporal_bwd_kernel — one warp per (clip, head, token): T x T problem entirely in shared memory.
#include "bwd_kernels.h"

#include "device_utils.cuh"

namespace maed {
using namespace bw;

static constexpr int kHd = 64;           // head dim
static constexpr int kLd = 68;           // padded fp32 row (16-byte aligned rows, conflict-free LDS.128 across rows)
static constexpr int kBq = 32;           // query rows per block
static constexpr int kMaxTok = 256;      // one key row per thread

__global__ void __launch_bounds__(256, 1)
attn_spatial_bwd_kernel(const __half* __restrict__ qkv, long long plane, const float* __restrict__ d_out, int ntok, int heads,
                        float scale, int accumulate, float* __restrict__ d_qkv, int pld) {
  extern __shared__ __align__(16) float smem_f[];
  float* sK = smem_f;                               // [ntok][kLd]
  float* sV = sK + (size_t)ntok * kLd;              // [ntok][kLd]
  float* sQ = sV + (size_t)ntok * kLd;              // [kBq][kLd]
  float* sdO = sQ + kBq * kLd;                      // [kBq][kLd]
  float* sP = sdO + kBq * kLd;                      // [kBq][pld]
  float* sdS = sP + (size_t)kBq * pld;              // [kBq][pld]
  const int bt = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * kHd;                   // qkv row
  const int ldo = heads * kHd;
  const long long row0 = (long long)bt * ntok;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // K, V of this (frame, head): planes -> fp32 shared memory
  for (int e = tid; e < ntok * (kHd / 4); e += 256) {
    const int j = e / (kHd / 4), d4 = (e % (kHd / 4)) * 4;
    const __half* kp = qkv + (row0 + j) * ld + heads * kHd + h * kHd + d4;
    const __half* vp = qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd + d4;
    *reinterpret_cast<float4*>(sK + j * kLd + d4) = plane ? load_planes4(kp, plane)
        : make_float4(__half2float(kp[0]), __half2float(kp[1]), __half2float(kp[2]), __half2float(kp[3]));
    *reinterpret_cast<float4*>(sV + j * kLd + d4) = plane ? load_planes4(vp, plane)
        : make_float4(__half2float(vp[0]), __half2float(vp[1]), __half2float(vp[2]), __half2float(vp[3]));
  }
  float dK[kHd], dV[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) { dK[d] = 0.f; dV[d] = 0.f; }
  const int j = tid;                                 // key row owned by this thread
  const bool has_key = j < ntok;

  for (int i0 = 0; i0 < ntok; i0 += kBq) {
    const int nq = min(kBq, ntok - i0);
    __syncthreads();                                 // previous block done with sQ / sdO / sP / sdS (and K/V loaded)
    for (int e = tid; e < kBq * (kHd / 4); e += 256) {
      const int i = e / (kHd / 4), d4 = (e % (kHd / 4)) * 4;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f), g = q;
      if (i < nq) {
        const __half* qp = qkv + (row0 + i0 + i) * ld + h * kHd + d4;
        q = plane ? load_planes4(qp, plane)
                  : make_float4(__half2float(qp[0]), __half2float(qp[1]), __half2float(qp[2]), __half2float(qp[3]));
        g = *reinterpret_cast<const float4*>(d_out + (row0 + i0 + i) * ldo + h * kHd + d4);
      }
      *reinterpret_cast<float4*>(sQ + i * kLd + d4) = q;
      *reinterpret_cast<float4*>(sdO + i * kLd + d4) = g;
    }
    __syncthreads();
    // S = scale Q K^T and dP = dO V^T for key j, 8 query rows at a time
    if (has_key) {
      for (int ib = 0; ib < kBq; ib += 8) {
        float s[8], p[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) { s[r] = 0.f; p[r] = 0.f; }
#pragma unroll 4
        for (int d4 = 0; d4 < kHd; d4 += 4) {
          const float4 k4 = *reinterpret_cast<const float4*>(sK + j * kLd + d4);
          const float4 v4 = *reinterpret_cast<const float4*>(sV + j * kLd + d4);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const float4 q4 = *reinterpret_cast<const float4*>(sQ + (ib + r) * kLd + d4);
            const float4 g4 = *reinterpret_cast<const float4*>(sdO + (ib + r) * kLd + d4);
            s[r] += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
            p[r] += g4.x * v4.x + g4.y * v4.y + g4.z * v4.z + g4.w * v4.w;
          }
        }
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          sP[(ib + r) * pld + j] = s[r] * scale;
          sdS[(ib + r) * pld + j] = p[r];
        }
      }
    }
    __syncthreads();
    // row softmax, D_i and dS: warp w handles rows w, w+8, w+16, w+24
    for (int i = warp; i < kBq; i += 8) {
      float* pr = sP + i * pld;
      float* dr = sdS + i * pld;
      float mx = -INFINITY;
      for (int c = lane; c < ntok; c += 32) mx = fmaxf(mx, pr[c]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int c = lane; c < ntok; c += 32) { const float e = expf(pr[c] - mx); pr[c] = e; sum += e; }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
      float dd = 0.f;
      for (int c = lane; c < ntok; c += 32) { const float pv = pr[c] * inv; pr[c] = pv; dd += pv * dr[c]; }
      dd = warp_sum(dd);
      for (int c = lane; c < ntok; c += 32) dr[c] = scale * pr[c] * (dr[c] - dd);
    }
    __syncthreads();
    // dQ[i, d] = sum_j dS[i, j] K[j, d]: thread -> (row i = tid/8, 8 columns)
    {
      const int i = tid >> 3, d0 = (tid & 7) * 8;
      float a[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] = 0.f;
      const float* dr = sdS + i * pld;
      for (int c = 0; c < ntok; ++c) {
        const float w = dr[c];
        const float4 k0 = *reinterpret_cast<const float4*>(sK + c * kLd + d0);
        const float4 k1 = *reinterpret_cast<const float4*>(sK + c * kLd + d0 + 4);
        a[0] += w * k0.x; a[1] += w * k0.y; a[2] += w * k0.z; a[3] += w * k0.w;
        a[4] += w * k1.x; a[5] += w * k1.y; a[6] += w * k1.z; a[7] += w * k1.w;
      }
      if (i < nq) {
        float* o = d_qkv + (row0 + i0 + i) * ld + h * kHd + d0;
        float4 o0 = make_float4(a[0], a[1], a[2], a[3]), o1 = make_float4(a[4], a[5], a[6], a[7]);
        if (accumulate) {
          const float4 p0 = *reinterpret_cast<const float4*>(o), p1 = *reinterpret_cast<const float4*>(o + 4);
          o0.x += p0.x; o0.y += p0.y; o0.z += p0.z; o0.w += p0.w;
          o1.x += p1.x; o1.y += p1.y; o1.z += p1.z; o1.w += p1.w;
        }
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + 4) = o1;
      }
    }
    // dK[j] += sum_i dS[i, j] Q[i];  dV[j] += sum_i P[i, j] dO[i]   (rows >= nq hold zeros in sQ / sdO)
    if (has_key) {
      for (int i = 0; i < nq; ++i) {
        const float ds = sdS[i * pld + j], pv = sP[i * pld + j];
#pragma unroll
        for (int d4 = 0; d4 < kHd; d4 += 4) {
          const float4 q4 = *reinterpret_cast<const float4*>(sQ + i * kLd + d4);
          const float4 g4 = *reinterpret_cast<const float4*>(sdO + i * kLd + d4);
          dK[d4] += ds * q4.x; dK[d4 + 1] += ds * q4.y; dK[d4 + 2] += ds * q4.z; dK[d4 + 3] += ds * q4.w;
          dV[d4] += pv * g4.x; dV[d4 + 1] += pv * g4.y; dV[d4 + 2] += pv * g4.z; dV[d4 + 3] += pv * g4.w;
        }
      }
    }
  }
  if (has_key) {
    float* ok = d_qkv + (row0 + j) * ld + heads * kHd + h * kHd;
    float* ov = d_qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd;
#pragma unroll
    for (int d4 = 0; d4 < kHd; d4 += 4) {
      float4 a = make_float4(dK[d4], dK[d4 + 1], dK[d4 + 2], dK[d4 + 3]);
      float4 b = make_float4(dV[d4], dV[d4 + 1], dV[d4 + 2], dV[d4 + 3]);
      if (accumulate) {
        const float4 pa = *reinterpret_cast<const float4*>(ok + d4), pb = *reinterpret_cast<const float4*>(ov + d4);
        a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
        b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
      }
      *reinterpret_cast<float4*>(ok + d4) = a;
      *reinterpret_cast<float4*>(ov + d4) = b;
    }
  }
}

int attn_spatial_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int BT, int ntok, int heads, float scale,
                     int accumulate, float* d_qkv, cudaStream_t st) {
  MAED_CHECK_ARG(ntok >= 1 && ntok <= kMaxTok, "attn_spatial_bwd: ntok=%d unsupported (1..%d)", ntok, kMaxTok);
  const int pld = ((ntok + 3) & ~3) + 4;
  const size_t smem = ((size_t)2 * ntok * kLd + 2 * kBq * kLd + (size_t)2 * kBq * pld) * sizeof(float);
  MAED_CHECK_ARG(smem <= 232448, "attn_spatial_bwd: %zu bytes of shared memory needed", smem);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_spatial_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  attn_spatial_bwd_kernel<<<BT * heads, 256, smem, st>>>(qkv_hi, qkv_plane, d_out, ntok, heads, scale, accumulate, d_qkv, pld);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ======================================================================================== temporal
// One warp per (clip, head, token).  Rows of frame t: (b*T + t)*ntok + n.  Shared memory per warp:
// q, k, v, dO [T][kLd] + P, dS [T][T+1].
static constexpr int kTmpWarps = 4;
__global__ void __launch_bounds__(kTmpWarps * 32)
attn_temporal_bwd_kernel(const __half* __restrict__ qkv, long long plane, const float* __restrict__ d_out, int B, int T,
                         int ntok, int heads, float scale, int accumulate, float* __restrict__ d_qkv) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long prob = (long long)blockIdx.x * kTmpWarps + warp;
  const long long nprob = (long long)B * heads * ntok;
  if (prob >= nprob) return;                         // warps are independent: no block-level barrier below
  const int per_warp = 4 * T * kLd + 2 * T * (T + 1);
  float* sq = smem_f + (size_t)warp * per_warp;
  float* sk = sq + T * kLd;
  float* sv = sk + T * kLd;
  float* sg = sv + T * kLd;
  float* sP = sg + T * kLd;
  float* sD = sP + T * (T + 1);
  const int h = (int)(prob % heads);                  // heads fastest: adjacent warps touch adjacent 128-byte head slices
  const int n = (int)((prob / heads) % ntok);
  const int b = (int)(prob / ((long long)ntok * heads));
  const int ld = 3 * heads * kHd, ldo = heads * kHd;
  // gather: lane covers 2 columns (d = 2*lane, 2*lane+1) of every frame row
  for (int t = 0; t < T; ++t) {
    const long long row = ((long long)b * T + t) * ntok + n;
    const __half* qp = qkv + row * ld + h * kHd + 2 * lane;
    const __half* kp = qp + heads * kHd;
    const __half* vp = kp + heads * kHd;
    float2 q = __half22float2(*reinterpret_cast<const __half2*>(qp));
    float2 k = __half22float2(*reinterpret_cast<const __half2*>(kp));
    float2 v = __half22float2(*reinterpret_cast<const __half2*>(vp));
    if (plane) {
      const float2 ql = __half22float2(*reinterpret_cast<const __half2*>(qp + plane));
      const float2 kl = __half22float2(*reinterpret_cast<const __half2*>(kp + plane));
      const float2 vl = __half22float2(*reinterpret_cast<const __half2*>(vp + plane));
      q.x += ql.x; q.y += ql.y; k.x += kl.x; k.y += kl.y; v.x += vl.x; v.y += vl.y;
    }
    const float2 g = *reinterpret_cast<const float2*>(d_out + row * ldo + h * kHd + 2 * lane);
    *reinterpret_cast<float2*>(sq + t * kLd + 2 * lane) = q;
    *reinterpret_cast<float2*>(sk + t * kLd + 2 * lane) = k;
    *reinterpret_cast<float2*>(sv + t * kLd + 2 * lane) = v;
    *reinterpret_cast<float2*>(sg + t * kLd + 2 * lane) = g;
  }
  __syncwarp();
  // S and dP for every (i, j) pair
  for (int e = lane; e < T * T; e += 32) {
    const int i = e / T, j = e % T;
    float s = 0.f, p = 0.f;
#pragma unroll 4
    for (int d4 = 0; d4 < kHd; d4 += 4) {
      const float4 q4 = *reinterpret_cast<const float4*>(sq + i * kLd + d4);
      const float4 k4 = *reinterpret_cast<const float4*>(sk + j * kLd + d4);
      const float4 g4 = *reinterpret_cast<const float4*>(sg + i * kLd + d4);
      const float4 v4 = *reinterpret_cast<const float4*>(sv + j * kLd + d4);
      s += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
      p += g4.x * v4.x + g4.y * v4.y + g4.z * v4.z + g4.w * v4.w;
    }
    sP[i * (T + 1) + j] = s * scale;
    sD[i * (T + 1) + j] = p;
  }
  __syncwarp();
  // softmax rows (lane = row; T <= 32), D_i, dS
  if (lane < T) {
    float* pr = sP + lane * (T + 1);
    float* dr = sD + lane * (T + 1);
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) mx = fmaxf(mx, pr[j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) { const float e = expf(pr[j] - mx); pr[j] = e; sum += e; }
    const float inv = 1.0f / sum;
    float dd = 0.f;
    for (int j = 0; j < T; ++j) { pr[j] *= inv; dd += pr[j] * dr[j]; }
    for (int j = 0; j < T; ++j) dr[j] = scale * pr[j] * (dr[j] - dd);
  }
  __syncwarp();
  // dQ[i], dK[j], dV[j]: lane covers columns d = 2*lane, 2*lane+1
  for (int t = 0; t < T; ++t) {
    float2 aq = make_float2(0.f, 0.f), ak = aq, av = aq;
    for (int u = 0; u < T; ++u) {
      const float ds_tu = sD[t * (T + 1) + u];       // dS[t, u]
      const float ds_ut = sD[u * (T + 1) + t];       // dS[u, t]
      const float p_ut = sP[u * (T + 1) + t];        // P[u, t]
      const float2 ku = *reinterpret_cast<const float2*>(sk + u * kLd + 2 * lane);
      const float2 qu = *reinterpret_cast<const float2*>(sq + u * kLd + 2 * lane);
      const float2 gu = *reinterpret_cast<const float2*>(sg + u * kLd + 2 * lane);
      aq.x += ds_tu * ku.x; aq.y += ds_tu * ku.y;
      ak.x += ds_ut * qu.x; ak.y += ds_ut * qu.y;
      av.x += p_ut * gu.x;  av.y += p_ut * gu.y;
    }
    const long long row = ((long long)b * T + t) * ntok + n;
    float* oq = d_qkv + row * ld + h * kHd + 2 * lane;
    float* ok = oq + heads * kHd;
    float* ov = ok + heads * kHd;
    if (accumulate) {
      const float2 pq = *reinterpret_cast<const float2*>(oq), pk = *reinterpret_cast<const float2*>(ok),
                   pv = *reinterpret_cast<const float2*>(ov);
      aq.x += pq.x; aq.y += pq.y; ak.x += pk.x; ak.y += pk.y; av.x += pv.x; av.y += pv.y;
    }
    *reinterpret_cast<float2*>(oq) = aq;
    *reinterpret_cast<float2*>(ok) = ak;
    *reinterpret_cast<float2*>(ov) = av;
  }
}

int attn_temporal_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int B, int T, int ntok, int heads,
                      float scale, int accumulate, float* d_qkv, cudaStream_t st) {
  MAED_CHECK_ARG(T >= 1 && T <= 32, "attn_temporal_bwd: T=%d unsupported (1..32)", T);
  const size_t smem = (size_t)kTmpWarps * (4 * T * kLd + 2 * T * (T + 1)) * sizeof(float);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_temporal_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem = smem;
  }
  const long long nprob = (long long)B * heads * ntok;
  attn_temporal_bwd_kernel<<<cdiv(nprob, kTmpWarps), kTmpWarps * 32, smem, st>>>(qkv_hi, qkv_plane, d_out, B, T, ntok, heads,
                                                                                scale, accumulate, d_qkv);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ======================================================================================== generic (coupling mode)
// Attention over `seq` contiguous rows per batch element (reference vision_transformer.py:180-204: the T * 197 tokens of a
// clip attend to each other).  CUDA-core fp32, flash-style recomputation, deterministic, slow (an ablation mode of the
// reference: 16x the attention work of the other modes): two kernels, one warp per query row / per key row.
//   kernel 1 (query i): m_i, l_i (softmax statistics), D_i = dO_i . O_i, dQ_i = sum_j dS_ij K_j; keeps (m, l, D) for kernel 2
//   kernel 2 (key j):   dK_j = sum_i dS_ij Q_i,  dV_j = sum_i P_ij dO_i          with dS = scale * P o (dO V^T - D)
static constexpr int kGenWarps = 4;

__device__ __forceinline__ void load_row64(const __half* p, long long plane, float* dst) {
#pragma unroll
  for (int d4 = 0; d4 < kHd; d4 += 4) {
    const float4 v = plane ? load_planes4(p + d4, plane)
                           : make_float4(__half2float(p[d4]), __half2float(p[d4 + 1]), __half2float(p[d4 + 2]), __half2float(p[d4 + 3]));
    dst[d4] = v.x; dst[d4 + 1] = v.y; dst[d4 + 2] = v.z; dst[d4 + 3] = v.w;
  }
}
__device__ __forceinline__ float dot_row64(const __half* p, long long plane, const float* s) {
  float a = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < kHd; d4 += 4) {
    const float4 v = plane ? load_planes4(p + d4, plane)
                           : make_float4(__half2float(p[d4]), __half2float(p[d4 + 1]), __half2float(p[d4 + 2]), __half2float(p[d4 + 3]));
    a += v.x * s[d4] + v.y * s[d4 + 1] + v.z * s[d4 + 2] + v.w * s[d4 + 3];
  }
  return a;
}

__global__ void __launch_bounds__(kGenWarps * 32)
attn_generic_bwd_q_kernel(const __half* __restrict__ qkv, long long plane, const float* __restrict__ d_out, int seq, int heads,
                          float scale, int accumulate, float* __restrict__ d_qkv, float* __restrict__ stats) {
  __shared__ float sq[kGenWarps][kHd], sg[kGenWarps][kHd], so[kGenWarps][kHd];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * kGenWarps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (i >= seq) return;                                    // whole warp leaves together (no block barriers below)
  const int ld = 3 * heads * kHd, ldo = heads * kHd;
  const long long row0 = (long long)b * seq;
  if (lane == 0) {
    load_row64(qkv + (row0 + i) * ld + h * kHd, plane, sq[warp]);
    for (int d = 0; d < kHd; ++d) sg[warp][d] = d_out[(row0 + i) * ldo + h * kHd + d];
  }
  __syncwarp();
  const float* q = sq[warp];
  const float* g = sg[warp];
  // pass A: row maximum
  float mx = -INFINITY;
  for (int j = lane; j < seq; j += 32) mx = fmaxf(mx, scale * dot_row64(qkv + (row0 + j) * ld + heads * kHd + h * kHd, plane, q));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  // pass B: l = sum exp, o = sum exp * V   (per-lane partials, then a fixed-order warp reduction)
  float l = 0.f, acc[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) acc[d] = 0.f;
  for (int j = lane; j < seq; j += 32) {
    const float e = expf(scale * dot_row64(qkv + (row0 + j) * ld + heads * kHd + h * kHd, plane, q) - mx);
    l += e;
    float v[kHd];
    load_row64(qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd, plane, v);
#pragma unroll
    for (int d = 0; d < kHd; ++d) acc[d] += e * v[d];
  }
  l = warp_sum(l);
  float D = 0.f;
#pragma unroll
  for (int d = 0; d < kHd; ++d) D += warp_sum(acc[d]) * g[d];
  D /= l;
  if (lane == 0) {
    float* st = stats + ((long long)(b * heads + h) * seq + i) * 3;
    st[0] = mx; st[1] = l; st[2] = D;
  }
  // pass C: dQ_i = sum_j scale * P_ij (dO_i . V_j - D) K_j
#pragma unroll
  for (int d = 0; d < kHd; ++d) acc[d] = 0.f;
  const float inv_l = 1.0f / l;
  for (int j = lane; j < seq; j += 32) {
    float k[kHd];
    load_row64(qkv + (row0 + j) * ld + heads * kHd + h * kHd, plane, k);
    float sdot = 0.f;
#pragma unroll
    for (int d = 0; d < kHd; ++d) sdot += q[d] * k[d];
    const float pij = expf(scale * sdot - mx) * inv_l;
    const float dp = dot_row64(qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd, plane, g);
    const float ds = scale * pij * (dp - D);
#pragma unroll
    for (int d = 0; d < kHd; ++d) acc[d] += ds * k[d];
  }
#pragma unroll
  for (int d = 0; d < kHd; ++d) {
    const float t = warp_sum(acc[d]);
    if (lane == (d & 31)) so[warp][d] = t;
  }
  __syncwarp();
  for (int d = lane; d < kHd; d += 32) {
    float* o = d_qkv + (row0 + i) * ld + h * kHd + d;
    *o = (accumulate ? *o : 0.f) + so[warp][d];
  }
}

__global__ void __launch_bounds__(kGenWarps * 32)
attn_generic_bwd_kv_kernel(const __half* __restrict__ qkv, long long plane, const float* __restrict__ d_out, int seq, int heads,
                           float scale, int accumulate, float* __restrict__ d_qkv, const float* __restrict__ stats) {
  __shared__ float sk[kGenWarps][kHd], sv[kGenWarps][kHd], so[kGenWarps][2 * kHd];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * kGenWarps + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (j >= seq) return;
  const int ld = 3 * heads * kHd, ldo = heads * kHd;
  const long long row0 = (long long)b * seq;
  if (lane == 0) {
    load_row64(qkv + (row0 + j) * ld + heads * kHd + h * kHd, plane, sk[warp]);
    load_row64(qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd, plane, sv[warp]);
  }
  __syncwarp();
  const float* k = sk[warp];
  const float* v = sv[warp];
  float dk[kHd], dv[kHd];
#pragma unroll
  for (int d = 0; d < kHd; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
  for (int i = lane; i < seq; i += 32) {
    const float* st = stats + ((long long)(b * heads + h) * seq + i) * 3;
    float q[kHd];
    load_row64(qkv + (row0 + i) * ld + h * kHd, plane, q);
    const float* g = d_out + (row0 + i) * ldo + h * kHd;
    float sdot = 0.f, dp = 0.f;
#pragma unroll
    for (int d = 0; d < kHd; ++d) { sdot += q[d] * k[d]; dp += g[d] * v[d]; }
    const float pij = expf(scale * sdot - st[0]) / st[1];
    const float ds = scale * pij * (dp - st[2]);
#pragma unroll
    for (int d = 0; d < kHd; ++d) { dk[d] += ds * q[d]; dv[d] += pij * g[d]; }
  }
#pragma unroll
  for (int d = 0; d < kHd; ++d) {
    const float a = warp_sum(dk[d]), c = warp_sum(dv[d]);
    if (lane == (d & 31)) { so[warp][d] = a; so[warp][kHd + d] = c; }
  }
  __syncwarp();
  for (int d = lane; d < kHd; d += 32) {
    float* ok = d_qkv + (row0 + j) * ld + heads * kHd + h * kHd + d;
    float* ov = d_qkv + (row0 + j) * ld + 2 * heads * kHd + h * kHd + d;
    *ok = (accumulate ? *ok : 0.f) + so[warp][d];
    *ov = (accumulate ? *ov : 0.f) + so[warp][kHd + d];
  }
}

// stats: batch * heads * seq * 3 floats of scratch
int attn_generic_bwd(const __half* qkv_hi, long long qkv_plane, const float* d_out, int batch, int seq, int heads, float scale,
                     int accumulate, float* d_qkv, float* stats, cudaStream_t st) {
  MAED_CHECK_ARG(qkv_hi && d_out && d_qkv && stats && batch >= 1 && seq >= 1 && heads >= 1, "attn_generic_bwd: bad argument");
  const dim3 grid(cdiv(seq, kGenWarps), heads, batch);
  attn_generic_bwd_q_kernel<<<grid, kGenWarps * 32, 0, st>>>(qkv_hi, qkv_plane, d_out, seq, heads, scale, accumulate, d_qkv, stats);
  MAED_BW_LAUNCH_CHECK();
  attn_generic_bwd_kv_kernel<<<grid, kGenWarps * 32, 0, st>>>(qkv_hi, qkv_plane, d_out, seq, heads, scale, accumulate, d_qkv, stats);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
