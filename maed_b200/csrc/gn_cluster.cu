// GroupNorm(32) forward and backward of the TRAINING path as one kernel per pass pair: a thread-block cluster owns an image.
//
// The multi-kernel versions (kernels.cu gn_stats + gn_apply, bwd_kernels.cu stage 1 / 2 / apply) read the fp32 conv output x
// (and dy) from HBM twice: once for the per-image reduction, once to produce the result.  Here the CS CTAs of a cluster split
// the rows of ONE image, reduce over distributed shared memory, and walk their rows a second time straight away — the
// second read comes out of L2 (an image is 0.2 .. 3.2 MB per tensor; the host picks the residency so that the images in
// flight fit).  HBM traffic per element: forward 12 -> 8 bytes, backward 20 -> 12 bytes.
//
//   phase 1   per-channel partial sums over the CTA's rows (a thread owns 4 neighbouring channels, 16-byte loads)
//   exchange  cluster.sync; CTA r sums the channels of ITS 32 / CS groups over all peers (DSMEM loads), forms the group
//             statistics and stores them into every peer (DSMEM stores); cluster.sync
//   phase 2   normalise (forward) / dx (backward) -> fp16 hi / lo planes
//
// Reference semantics: resnetv2.py:35-49 (GroupNormAct), F.group_norm backward; same formulas as the multi-kernel path, which
// stays as the fallback for shapes this file does not cover and is what the CPU emulator build runs.
#include <cooperative_groups.h>

#include <cstdlib>

#include "bwd_kernels.h"
#include "device_utils.cuh"

namespace cg = cooperative_groups;

namespace maed {
using namespace bw;

namespace {


struct GnClusterParams {
  int HW, C, CS, relu;
  float eps;
  const float* x; const float* dy;
  const float* gamma; const float* beta;        // backward: beta != nullptr <=> ReLU mask recomputed
  const __half* res; long long res_plane;       // forward: shortcut planes added before the ReLU (or nullptr)
  __half* out; long long out_plane;
  double* stats;                                // forward: [n][32][2] (sum, sumsq) written for the tape; backward: read
  float* dgb;                                   // backward: [n][2][C] per-image (sum dy*xhat | sum dy)
};

// shared memory: float4 acc[2][kThreads] | float part[2][C] | float grp[32][2] | float mean[32], rstd[32]
template <int kThreads>
__device__ __forceinline__ void reduce_rows(float4* acc, float* part, int C, int c4n, int rstep, const float4& av,
                                            const float4& bv) {
  acc[threadIdx.x] = av;
  acc[kThreads + threadIdx.x] = bv;
  __syncthreads();
  if (threadIdx.x < c4n) {
    float4 a = av, b = bv;
    for (int k = 1; k < rstep; ++k) {
      const float4 pa = acc[threadIdx.x + k * c4n], pb = acc[kThreads + threadIdx.x + k * c4n];
      a.x += pa.x; a.y += pa.y; a.z += pa.z; a.w += pa.w;
      b.x += pb.x; b.y += pb.y; b.z += pb.z; b.w += pb.w;
    }
    *reinterpret_cast<float4*>(part + threadIdx.x * 4) = a;
    *reinterpret_cast<float4*>(part + C + threadIdx.x * 4) = b;
  }
}

// kThreads = 512, 2 CTAs per SM: the general shape.  kThreads = 128, up to 8 CTAs per SM: small maps whose whole batch fits
// the L2 — all clusters of the grid are resident at once, so the phases of different images overlap instead of running in
// waves (a CTA's life is a chain of ~8 memory / cluster-barrier latencies).
template <bool BWD, int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 128 ? 8 : 2) gn_cluster_kernel(const GnClusterParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = p.C, HW = p.HW, CS = p.CS, c4n = C >> 2, gsz = C / 32;
  float4* acc = reinterpret_cast<float4*>(smem_raw);
  float* part = reinterpret_cast<float*>(acc + 2 * kThreads);          // [2][C]
  float* grp = part + 2 * C;                                           // [32][2]: fwd (mean, rstd); bwd (m1, m2)
  float* s_mean = grp + 64;
  float* s_rstd = s_mean + 32;
  const int rank = (int)cluster.block_rank();
  const int n = blockIdx.x / CS;
  const int rpc = (HW + CS - 1) / CS;
  const int hw0 = rank * rpc, hw1 = min(HW, hw0 + rpc);
  const long long base = (long long)n * HW * C;
  const int c4 = threadIdx.x % c4n, rr = threadIdx.x / c4n, rstep = kThreads / c4n, c = c4 * 4;
  const double cnt = (double)HW * gsz;

  if (BWD) {
    if (threadIdx.x < 32) {
      const double m = p.stats[((long long)n * 32 + threadIdx.x) * 2] / cnt;
      double var = p.stats[((long long)n * 32 + threadIdx.x) * 2 + 1] / cnt - m * m;
      if (var < 0.0) var = 0.0;
      s_mean[threadIdx.x] = (float)m;
      s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    __syncthreads();
  }
  // ------------------------------------------------------------------------------------------------ phase 1
  float m[4], r[4], mg[4], mb[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    if (BWD) {
      m[e] = s_mean[(c + e) / gsz]; r[e] = s_rstd[(c + e) / gsz];
      mg[e] = p.beta ? p.gamma[c + e] : 0.f; mb[e] = p.beta ? p.beta[c + e] : 1.f;      // no ReLU: xhat * 0 + 1 > 0 always
    }
  }
  {
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int hw = hw0 + rr; hw < hw1; hw += rstep) {
      const long long off = base + (long long)hw * C + c;
      const float4 xv = *reinterpret_cast<const float4*>(p.x + off);
      const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
      if (BWD) {
        const float4 dv = *reinterpret_cast<const float4*>(p.dy + off);
        const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float xhat = (xx[e] - m[e]) * r[e];
          const float d = (xhat * mg[e] + mb[e] > 0.f) ? dd[e] : 0.f;
          a[e] += d * xhat;
          b[e] += d;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) { a[e] += xx[e]; b[e] += xx[e] * xx[e]; }
      }
    }
    reduce_rows<kThreads>(acc, part, C, c4n, rstep, make_float4(a[0], a[1], a[2], a[3]), make_float4(b[0], b[1], b[2], b[3]));
  }
  cluster.sync();
  // ----------------------------------------------------------------------------------------------- exchange
  // CTA `rank` owns the channels of groups [rank * 32 / CS, (rank + 1) * 32 / CS): totals over all peers, then group sums
  {
    const int cpc = C / CS, gpc = 32 / CS;                     // channels / groups per CTA
    float* tmp = reinterpret_cast<float*>(acc);                // [2][cpc] reused after the row reduction (synchronised above)
    if (threadIdx.x < cpc) {
      const int ch = rank * cpc + threadIdx.x;
      double ta = 0.0, tb = 0.0;
      for (int q = 0; q < CS; ++q) {
        const float* peer = cluster.map_shared_rank(part, q);
        ta += (double)peer[ch];
        tb += (double)peer[C + ch];
      }
      if (BWD) {
        p.dgb[(long long)n * 2 * C + ch] = (float)ta;
        p.dgb[(long long)n * 2 * C + C + ch] = (float)tb;
        const float g = p.gamma[ch];
        tmp[threadIdx.x] = (float)tb * g;                      // sum dy*gamma
        tmp[cpc + threadIdx.x] = (float)ta * g;                // sum dy*gamma*xhat
      } else {
        // doubles do not fit the float scratch: keep (sum, sumsq) as hi + lo float pairs
        tmp[threadIdx.x] = (float)ta; tmp[cpc + threadIdx.x] = (float)(ta - (double)(float)ta);
        tmp[2 * cpc + threadIdx.x] = (float)tb; tmp[3 * cpc + threadIdx.x] = (float)(tb - (double)(float)tb);
      }
    }
    __syncthreads();
    if (threadIdx.x < gpc) {
      const int g = rank * gpc + threadIdx.x;
      float v0, v1;
      if (BWD) {
        float s1 = 0.f, s2 = 0.f;
        for (int k = 0; k < gsz; ++k) { s1 += tmp[threadIdx.x * gsz + k]; s2 += tmp[cpc + threadIdx.x * gsz + k]; }
        v0 = (float)((double)s1 / cnt);
        v1 = (float)((double)s2 / cnt);
      } else {
        double s = 0.0, q = 0.0;
        for (int k = 0; k < gsz; ++k) {
          s += (double)tmp[threadIdx.x * gsz + k] + (double)tmp[cpc + threadIdx.x * gsz + k];
          q += (double)tmp[2 * cpc + threadIdx.x * gsz + k] + (double)tmp[3 * cpc + threadIdx.x * gsz + k];
        }
        p.stats[((long long)n * 32 + g) * 2] = s;
        p.stats[((long long)n * 32 + g) * 2 + 1] = q;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        v0 = (float)mean;
        v1 = (float)(1.0 / sqrt(var + (double)p.eps));
      }
      for (int q = 0; q < CS; ++q) {
        float* peer = cluster.map_shared_rank(grp, q);
        peer[g * 2] = v0;
        peer[g * 2 + 1] = v1;
      }
    }
  }
  cluster.sync();
  // ------------------------------------------------------------------------------------------------ phase 2
  float ga[4], be[4], g0[4], g1[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int g = (c + e) / gsz;
    ga[e] = p.gamma[c + e];
    be[e] = p.beta ? p.beta[c + e] : 0.f;
    g0[e] = grp[g * 2]; g1[e] = grp[g * 2 + 1];
  }
#pragma unroll 2
  for (int hw = hw0 + rr; hw < hw1; hw += rstep) {
    const long long off = base + (long long)hw * C + c;
    const float4 xv = *reinterpret_cast<const float4*>(p.x + off);
    const float xx[4] = {xv.x, xv.y, xv.z, xv.w};
    float o[4];
    if (BWD) {
      const float4 dv = *reinterpret_cast<const float4*>(p.dy + off);
      const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xhat = (xx[e] - m[e]) * r[e];
        const float d = (!p.beta || xhat * ga[e] + be[e] > 0.f) ? dd[e] : 0.f;
        o[e] = r[e] * (d * ga[e] - g0[e] - xhat * g1[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (xx[e] - g0[e]) * g1[e] * ga[e] + be[e];
      if (p.res) {
        const float4 rv = load_planes4(p.res + off, p.res_plane);
        o[0] += rv.x; o[1] += rv.y; o[2] += rv.z; o[3] += rv.w;
      }
      if (p.relu) {
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
      }
    }
    store_split4(p.out + off, p.out_plane, make_float4(o[0], o[1], o[2], o[3]));
  }
}

bool cluster_enabled() {
  static const bool on = !(getenv("MAED_B200_GN_CLUSTER") && atoi(getenv("MAED_B200_GN_CLUSTER")) == 0);
  return on;
}

template <bool BWD, int kThreads>
int launch_t(GnClusterParams& p, int n_img, size_t smem_floor, cudaStream_t st) {
  size_t smem = 2 * kThreads * sizeof(float4) + (size_t)(2 * p.C + 64 + 64) * sizeof(float);
  if (smem < smem_floor) smem = smem_floor;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gn_cluster_kernel<BWD, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(n_img * p.CS);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MAED_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gn_cluster_kernel<BWD, kThreads>, p));
  count_launch();
  return MAED_OK;
}

template <bool BWD>
int launch(GnClusterParams& p, int n_img, cudaStream_t st) {
  const int C = p.C, c4n = C / 4, CS = 8;
  // a thread owns 4 channels and a CTA row step covers whole rows; the exchange gives every CTA whole groups
  if (!cluster_enabled() || C % 32 != 0 || p.HW < CS) return MAED_ERR_UNSUPPORTED;
  p.CS = CS;
  // what the two passes share must stay in the L2 (126 MB, half of it as budget) between them
  const double per_image = (double)p.HW * C * 4.0 * (BWD ? 2.0 : 1.0);
  const double budget = 64e6;
  if (c4n <= 128 && 128 % c4n == 0 && C / CS <= 128 && per_image * n_img <= budget &&
      (long long)n_img * CS <= 7LL * sm_count())
    return launch_t<BWD, 128>(p, n_img, 0, st);                        // the whole batch in flight
  if (c4n > 512 || 512 % c4n != 0 || C / CS > 512) return MAED_ERR_UNSUPPORTED;
  if (per_image * (2.0 * sm_count() / CS) <= budget) return launch_t<BWD, 512>(p, n_img, 0, st);
  // larger images would need one CTA per SM to fit: measured slower than the multi-kernel path (574 vs 366 us on the
  // 56 x 56 x 256 maps), which streams with several times the threads in flight
  return MAED_ERR_UNSUPPORTED;
}

}  // namespace

int groupnorm_fwd_cluster(const float* x, const float* gamma, const float* beta, int n_img, int HW, int C, float eps, int relu,
                          const __half* res_hi, long long res_plane, __half* out_hi, long long out_plane, double* stats,
                          cudaStream_t st) {
  GnClusterParams p;
  memset(&p, 0, sizeof(p));
  p.HW = HW; p.C = C; p.relu = relu; p.eps = eps; p.x = x; p.gamma = gamma; p.beta = beta; p.res = res_hi; p.res_plane = res_plane;
  p.out = out_hi; p.out_plane = out_plane; p.stats = stats;
  return launch<false>(p, n_img, st);
}

int groupnorm_bwd_cluster(const float* dy, const float* x, const double* stats, const float* gamma, const float* relu_beta,
                          int n_img, int HW, int C, float eps, float* dgb_partial, __half* dx_hi, long long dx_plane,
                          cudaStream_t st) {
  GnClusterParams p;
  memset(&p, 0, sizeof(p));
  p.HW = HW; p.C = C; p.eps = eps; p.x = x; p.dy = dy; p.gamma = gamma; p.beta = relu_beta; p.out = dx_hi; p.out_plane = dx_plane;
  p.stats = const_cast<double*>(stats); p.dgb = dgb_partial;
  return launch<true>(p, n_img, st);
}

}  // namespace maed
