// Fused StdConv -> GroupNorm(32) -> (+shortcut) -> ReLU for the ResNetV2 backbone (reference resnetv2.py:189-204):
// the conv runs as a tcgen05 GEMM whose accumulators for ONE IMAGE x ONE CHANNEL BLOCK stay resident in tensor
// memory until the GroupNorm statistics of that image are known, so the fp32 conv output never goes to HBM.
//
//   item    = (image, channel block of BN output channels); BN is a multiple of the group size C/32
//   cluster = CS CTAs (1 or 4) share an item; CTA r owns the image's M tiles [r*TPC, (r+1)*TPC), TPC*BN <= 512 TMEM columns
//   main loop (per M tile): TMA -> smem -> tcgen05.mma (split fp16, 3 MMAs / K step) into TMEM columns [t*BN, (t+1)*BN)
//   pass 1 (epilogue warps, overlapped with the MMAs of later tiles): per-group sum / sum of squares of the valid rows
//   CTA reduce (smem) -> cluster reduce (DSMEM reads of the peers' partials) -> mean, rstd per group
//   pass 2: TMEM -> (x-mean)*rstd*gamma+beta (+ shortcut planes) -> ReLU -> fp16 hi/lo -> smem transpose -> coalesced stores
//
// Versus the unfused pipeline (GEMM writes fp32, gn_stats reads it, gn_apply reads it again and writes planes)
// this removes 12 of the 16 bytes of HBM traffic per conv-output element.
#include "gemm_host.h"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

static constexpr int kGnThreads = 256;
static constexpr int kGnMaxTiles = 8;
static constexpr int kGnMaxStages = 4;

struct GemmGnParams {
  int HW;                      // pixels per image
  int C;                       // total output channels (row pitch of out / residual), = N
  int K, num_k_blocks, nsplit, stages;
  int tiles_per_image, tpc;    // M tiles per image; tiles per CTA
  int n_blocks;                // C / BN
  int cluster;                 // CTAs per item
  uint32_t a_tx_bytes;
  // conv (tap) mode
  int conv, H, W, cin_blocks, KW, pad_h, pad_w, tile_h, tile_w, tiles_h, tiles_w;
  // epilogue
  const float* gamma; const float* beta; float eps; int relu;
  const __half* res; long long res_plane;
  __half* out; long long out_plane;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double* local_ptr, uint32_t rank) {
  uint32_t la = sm100::smem_u32(local_ptr), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  double v;
  asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
  return v;
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int BN, int GSZ>
__global__ void __launch_bounds__(kGnThreads, 1)
gemm_gn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmGnParams p) {
  using namespace sm100;
  constexpr int G = BN / GSZ;                         // groups in this channel block (<= 32)
  static_assert(G >= 1 && G <= 32, "bad group count");
  constexpr uint32_t kABytes = 128 * 64 * 2, kBBytes = BN * 64 * 2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int np = p.nsplit == 3 ? 2 : 1;
  const uint32_t stage_bytes = np * (kABytes + kBBytes);
  // transpose buffers of pass 2 alias the (by then drained) pipeline stages: [4 warps][2 planes][32 rows x 128 B]
  uint8_t* sStage = smem;
  double* s_warp_part = reinterpret_cast<double*>(smem + (size_t)p.stages * stage_bytes);   // [4][32][2]
  double* s_cta_part = s_warp_part + 4 * 32 * 2;                        // [G][2]  (read by cluster peers)
  float* s_mr = reinterpret_cast<float*>(s_cta_part + 32 * 2);          // [G][2] mean, rstd
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_mr + 64);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kGnMaxStages;
  uint64_t* tile_full = bars + 2 * kGnMaxStages;                        // [kGnMaxTiles]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kGnMaxStages + kGnMaxTiles);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0;
  const int item = blockIdx.x / p.cluster;
  const int img = item / p.n_blocks, nb = item % p.n_blocks;
  const int t_lo = crank * p.tpc;
  const int my_tiles = max(0, min(p.tpc, p.tiles_per_image - t_lo));

  if (warp == 0 && elect_one()) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int t = 0; t < kGnMaxTiles; ++t) mbar_init(&tile_full[t], 1);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_base_ptr, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int t = t_lo + tl;
        int h0 = 0, w0 = 0;
        if (p.conv) { h0 = (t / p.tiles_w) * p.tile_h; w0 = (t % p.tiles_w) * p.tile_w; }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + np * kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], np * (p.a_tx_bytes + kBBytes));
          for (int pl = 0; pl < np; ++pl) {
            if (p.conv) {
              const int tap = kb / p.cin_blocks, cb = kb % p.cin_blocks;
              tma_load_5d(sA + pl * kABytes, &tmA, &full_bar[stage], cb * 64, w0 + tap % p.KW - p.pad_w,
                          h0 + tap / p.KW - p.pad_h, img, pl);
            } else {
              tma_load_3d(sA + pl * kABytes, &tmA, &full_bar[stage], kb * 64, img * p.HW + t * 128, pl);
            }
            tma_load_3d(sB + pl * kBBytes, &tmB, &full_bar[stage], kb * 64, nb * BN, pl);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(128, BN, 0);
      int stage = 0; uint32_t phase = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        const uint32_t d = tmem_base + tl * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t aH = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t bH = aH + np * kABytes;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_desc_k_sw128(aH + k * 32), db = umma_desc_k_sw128(bH + k * 32);
            umma_f16(d, da, db, idesc, (kb | k) != 0);
            if (np == 2) {
              umma_f16(d, umma_desc_k_sw128(aH + kABytes + k * 32), db, idesc, 1);
              umma_f16(d, da, umma_desc_k_sw128(bH + kBBytes + k * 32), idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tile_full[tl]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- pass 1: group statistics of the valid rows
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
    float gs[G], gq[G];
#pragma unroll
    for (int g = 0; g < G; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    auto row_valid = [&](int t, long long* out_row) -> bool {
      if (p.conv) {
        const int lh = row_in_tile / p.tile_w, lw = row_in_tile % p.tile_w;
        const int h = (t / p.tiles_w) * p.tile_h + lh, w = (t % p.tiles_w) * p.tile_w + lw;
        *out_row = ((long long)img * p.H + h) * p.W + w;
        return lh < p.tile_h && h < p.H && w < p.W;
      }
      const int r = t * 128 + row_in_tile;
      *out_row = (long long)img * p.HW + r;
      return r < p.HW;
    };
    for (int tl = 0; tl < my_tiles; ++tl) {
      long long orow;
      const bool valid = row_valid(t_lo + tl, &orow);
      mbar_wait(&tile_full[tl], 0);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + tl * BN + lane_off + c0, r);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]);
            gs[(c0 + j) / GSZ] += v;
            gq[(c0 + j) / GSZ] += v * v;
          }
        }
      }
    }
    // ---- warp -> CTA -> cluster reduction
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float s = gs[g], q = gq[g];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
      if (lane == 0) { s_warp_part[(ew * 32 + g) * 2] = (double)s; s_warp_part[(ew * 32 + g) * 2 + 1] = (double)q; }
    }
    named_bar(1, 128);
    if (threadIdx.x - 128 < G) {
      const int g = threadIdx.x - 128;
      double s = 0.0, q = 0.0;
#pragma unroll
      for (int w4 = 0; w4 < 4; ++w4) { s += s_warp_part[(w4 * 32 + g) * 2]; q += s_warp_part[(w4 * 32 + g) * 2 + 1]; }
      s_cta_part[g * 2] = s; s_cta_part[g * 2 + 1] = q;
    }
  }

  // every thread of every CTA of the cluster: partials are published
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();

  if (warp >= 4) {
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
    if (threadIdx.x - 128 < G) {
      const int g = threadIdx.x - 128;
      double s = 0.0, q = 0.0;
      for (int r = 0; r < p.cluster; ++r) {
        if (p.cluster > 1) { s += ld_dsmem_f64(&s_cta_part[g * 2], r); q += ld_dsmem_f64(&s_cta_part[g * 2 + 1], r); }
        else { s += s_cta_part[g * 2]; q += s_cta_part[g * 2 + 1]; }
      }
      const double cnt = (double)p.HW * GSZ;
      const double m = s / cnt;
      double var = q / cnt - m * m;
      if (var < 0.0) var = 0.0;
      s_mr[g * 2] = (float)m;
      s_mr[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
    }
    named_bar(1, 128);
    // -------------------------------------------------------------- pass 2: normalise, shortcut, ReLU, split, store
    uint8_t* st_hi = sStage + ew * 8192;
    uint8_t* st_lo = st_hi + 4096;
    const int ch0 = nb * BN;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int t = t_lo + tl;
      long long orow = 0;
      bool valid;
      if (p.conv) {
        const int lh = row_in_tile / p.tile_w, lw = row_in_tile % p.tile_w;
        const int h = (t / p.tiles_w) * p.tile_h + lh, w = (t % p.tiles_w) * p.tile_w + lw;
        orow = ((long long)img * p.H + h) * p.W + w;
        valid = lh < p.tile_h && h < p.H && w < p.W;
      } else {
        const int r = t * 128 + row_in_tile;
        orow = (long long)img * p.HW + r;
        valid = r < p.HW;
      }
      // rows of this warp as seen by the coalesced copy-out (lane -> (row, 16-byte chunk))
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 64) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + tl * BN + lane_off + c0 + half * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const int c = c0 + half * 32 + j;                    // channel inside the block
            const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + ch0 + c));
            const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + ch0 + c));
            const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int g = (c + e) / GSZ;
              v[j + e] = (__uint_as_float(r[j + e]) - s_mr[g * 2]) * s_mr[g * 2 + 1] * gg[e] + bb[e];
            }
          }
          if (p.res && valid) {
            const __half* rp = p.res + orow * p.C + ch0 + c0 + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 H = *reinterpret_cast<const uint4*>(rp + j);
              const uint4 L = *reinterpret_cast<const uint4*>(rp + j + p.res_plane);
              const __half2* h2 = reinterpret_cast<const __half2*>(&H);
              const __half2* l2 = reinterpret_cast<const __half2*>(&L);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(h2[e]), b = __half22float2(l2[e]);
                v[j + 2 * e] += a.x + b.x;
                v[j + 2 * e + 1] += a.y + b.y;
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          // stage this row's 32 channels (64 B per plane) into the warp's transpose buffers (16-byte chunks XOR-swizzled)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const __half2 h2 = __floats2half2_rn(v[j + 2 * e], v[j + 2 * e + 1]);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(v[j + 2 * e] - hf.x, v[j + 2 * e + 1] - hf.y);
              hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            const int q = half * 4 + (j >> 3);                   // logical 16-byte chunk (0..7) of the 128-byte row segment
            const uint32_t off = lane * 128 + ((q ^ (lane & 7)) << 4);
            *reinterpret_cast<uint4*>(st_hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(st_lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
          }
        }
        __syncwarp();
        // coalesced copy-out: 8 lanes cover one row segment of 64 channels (128 B), 4 rows per instruction
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3), pq = lane & 7;
          const int lq = pq ^ (rr & 7);
          const long long row_g = __shfl_sync(0xffffffffu, orow, rr);
          const int ok = __shfl_sync(0xffffffffu, (int)valid, rr);
          if (ok) {
            __half* dst = p.out + row_g * p.C + ch0 + c0 + lq * 8;
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(st_hi + rr * 128 + pq * 16);
            *reinterpret_cast<uint4*>(dst + p.out_plane) = *reinterpret_cast<const uint4*>(st_lo + rr * 128 + pq * 16);
          }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();               // peers may still be reading this CTA's partials
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host
template <int BN, int GSZ>
static int launch_gn(const CUtensorMap& tmA, const CUtensorMap& tmB, GemmGnParams& p, int items, cudaStream_t st) {
  const int np = p.nsplit == 3 ? 2 : 1;
  const size_t stage_bytes = (size_t)np * (128 * 64 * 2 + BN * 64 * 2);
  const size_t fixed = 1024 + (4 * 32 * 2 + 32 * 2) * 8 + 64 * 4 + (2 * kGnMaxStages + kGnMaxTiles) * 8 + 64;
  int stages = (int)((232448 - fixed) / stage_bytes);
  if (stages > kGnMaxStages) stages = kGnMaxStages;
  if (stages < 2) { set_error("gemm_gn: tile too large for shared memory"); return MAED_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * stage_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gemm_gn_kernel<BN, GSZ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(items * p.cluster);
  cfg.blockDim = dim3(kGnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MAED_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_gn_kernel<BN, GSZ>, tmA, tmB, p));
  count_launch();
  return MAED_OK;
}

// Returns MAED_ERR_UNSUPPORTED (without setting up a launch) for shapes the fused kernel does not cover, so the
// caller can use the unfused gemm + gn_stats + gn_apply pipeline.
int conv_gn_fused(const ConvGnArgs& a, cudaStream_t st) {
  const int np = a.nsplit == 3 ? 2 : 1;
  MAED_CHECK_ARG(a.C % 32 == 0, "conv_gn_fused: C=%d", a.C);
  const int gsz = a.C / 32;
  GemmGnParams p;
  memset(&p, 0, sizeof(p));
  p.HW = a.H_out * a.W_out; p.C = a.C; p.nsplit = a.nsplit;
  p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.relu = a.relu; p.res = a.res; p.res_plane = a.res_plane;
  p.out = a.out; p.out_plane = a.out_plane;
  CUtensorMap tmA, tmB;
  const long long M = (long long)a.n_img * p.HW;
  if (a.conv) {
    if (a.Cin % 64 != 0) return MAED_ERR_UNSUPPORTED;
    p.conv = 1; p.H = a.H_out; p.W = a.W_out; p.cin_blocks = a.Cin / 64; p.KW = a.KW; p.pad_h = a.pad_h; p.pad_w = a.pad_w;
    conv_tile_shape(p.H, p.W, &p.tile_h, &p.tile_w);
    p.tiles_h = cdiv(p.H, p.tile_h); p.tiles_w = cdiv(p.W, p.tile_w);
    p.tiles_per_image = p.tiles_h * p.tiles_w;
    p.K = a.KH * a.KW * a.Cin; p.num_k_blocks = a.KH * a.KW * p.cin_blocks;
    p.a_tx_bytes = (uint32_t)(p.tile_h * p.tile_w * 128);
    const uint64_t dims[5] = {(uint64_t)a.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)a.n_img, (uint64_t)np};
    const uint64_t str[4] = {(uint64_t)a.Cin * 2, (uint64_t)p.W * a.Cin * 2, (uint64_t)p.H * p.W * a.Cin * 2,
                             (uint64_t)(np == 2 ? a.a_plane : M * a.Cin) * 2};
    const uint32_t box[5] = {64, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, a.A, 5, dims, str, box));
  } else {
    p.K = a.K; p.num_k_blocks = cdiv(a.K, 64);
    p.tiles_per_image = cdiv(p.HW, 128);
    p.a_tx_bytes = 128 * 64 * 2;
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)M, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)a.K * 2, (uint64_t)(np == 2 ? a.a_plane : M * a.K) * 2};
    const uint32_t box[3] = {64, 128, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, a.A, 3, dims, str, box));
  }
  // channel-block width and cluster size: all tiles of an image must fit the TMEM of the cluster (512 columns per CTA)
  int bn, cluster;
  if (p.tiles_per_image <= 2 && a.C % 256 == 0 && gsz <= 256) { bn = 256; cluster = 1; }
  else if (p.tiles_per_image <= 8) { bn = 64; cluster = 1; }
  else if (p.tiles_per_image <= 32) { bn = 64; cluster = 4; }
  else return MAED_ERR_UNSUPPORTED;
  if (bn % gsz != 0 || a.C % bn != 0) return MAED_ERR_UNSUPPORTED;
  p.cluster = cluster;
  p.tpc = cdiv(p.tiles_per_image, cluster);
  if (p.tpc * bn > 512) return MAED_ERR_UNSUPPORTED;
  p.n_blocks = a.C / bn;
  {
    const uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)a.C, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)p.K * 2, (uint64_t)(np == 2 ? a.b_plane : (long long)a.C * p.K) * 2};
    const uint32_t box[3] = {64, (uint32_t)bn, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmB, a.B, 3, dims, str, box));
  }
  const int items = a.n_img * p.n_blocks;
  if (bn == 256 && gsz == 8) return launch_gn<256, 8>(tmA, tmB, p, items, st);
  if (bn == 256 && gsz == 32) return launch_gn<256, 32>(tmA, tmB, p, items, st);
  if (bn == 64 && gsz == 2) return launch_gn<64, 2>(tmA, tmB, p, items, st);
  if (bn == 64 && gsz == 4) return launch_gn<64, 4>(tmA, tmB, p, items, st);
  if (bn == 64 && gsz == 8) return launch_gn<64, 8>(tmA, tmB, p, items, st);
  if (bn == 64 && gsz == 16) return launch_gn<64, 16>(tmA, tmB, p, items, st);
  if (bn == 64 && gsz == 32) return launch_gn<64, 32>(tmA, tmB, p, items, st);
  return MAED_ERR_UNSUPPORTED;
}

}  // namespace maed
