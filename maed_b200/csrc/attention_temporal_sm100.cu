// Temporal attention of the STE (reference lib/models/vision_transformer.py:216-228: T x T attention across the frames of a clip
// for every (clip, head, token)) on the 5th-gen tensor cores, forward and backward.
//
// The reference permutes (B*T, HW, C) -> (B*HW, T, C), makes contiguous copies of q, k, v and runs 18 912 batched 16 x 16 x 64
// GEMMs.  Here the reshape is a TMA box: for one (clip, head) and NPT = 128 / T consecutive token positions the box
// {64 head-dim columns, NPT tokens, T frames} of the [B*T*HW, 3C] qkv planes lands in shared memory as ONE 128-row K-major
// operand tile whose row r = t * NPT + n holds frame t of token n.  S = Q K^T is then a single 128 x 128 x 64 tcgen05 product
// of which only the entries with equal token (r mod NPT == c mod NPT) are kept: the softmax warps mask the rest to zero, write
// P back into TMEM as fp16 hi/lo pairs and O = P V runs with P as the TMEM operand — 8x redundant tensor work on a kernel whose
// tensor work is negligible, in exchange for fully coalesced (1.5 KB-run) global traffic and no shared-memory inner products.
// T in {4, 8, 16, 32}; other T use the CUDA-core kernels of attention.cu / attention_bwd.cu.
//
// Forward: two CTAs per SM (96 KB of operands, 256 TMEM columns each) run the simple sequential tile loop and overlap each
// other's load / MMA / softmax phases.  Backward: attention_bwd_sm100.cu's two-orientation scheme on the same tiles.
#include "bwd_kernels.h"

#include "device_utils.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

namespace {

constexpr int kTileBytes = 128 * 128;          // one plane of one operand tile: 128 rows x 128 B (K-major, 128-byte swizzle)
constexpr int kFwdThreads = 192;               // warp 0: TMA + MMA issue, warp 1: TMEM alloc, warps 2-5: softmax / epilogue

struct TemporalParams {
  int B, T, ntok, heads, groups;               // groups = ceil(ntok / NPT) token groups per (clip, head)
  float scale_log2e;
  float* out_f32; __half* out_hi; long long out_plane;
  float* lse;                                  // optional [B*T*ntok, heads]: log2-domain log-sum-exp of every row (training tape)
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int NPT>
__global__ void __launch_bounds__(kFwdThreads, 2)
attn_temporal_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const TemporalParams p) {
  using namespace sm100;
  constexpr uint32_t kS = 0, kO = 128;         // TMEM columns: S -> P | O
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem;                          // [hi | lo] each
  uint8_t* sK = sQ + 2 * kTileBytes;
  uint8_t* sV = sK + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * kTileBytes);
  uint64_t* ld_full = bars + 0;
  uint64_t* mma_done = bars + 1;
  uint64_t* ew_done = bars + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = p.B * p.heads * p.groups;
  const int C = p.heads * 64;

  // token groups at the end of a frame leave part of a tile unwritten by TMA: start from finite (zero) operands
  for (int i = threadIdx.x; i < 6 * kTileBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && elect_one()) prefetch_tmap(&tmQKV);
  if (warp == 1 && elect_one()) {
    mbar_init(ld_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(ew_done, 4);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_base_ptr, 256); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      constexpr uint32_t idesc_s = umma_idesc_f16(128, 128, 0, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 0, 1);          // B (= V) MN-major
      uint32_t ph_ld = 0, ph_ew = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int g = item % p.groups, h = (item / p.groups) % p.heads, b = item / (p.groups * p.heads);
        mbar_arrive_expect_tx(ld_full, 6 * kTileBytes);
        for (int pl = 0; pl < 2; ++pl) {
          tma_load_5d(sQ + pl * kTileBytes, &tmQKV, ld_full, h * 64, g * NPT, 0, b, pl);
          tma_load_5d(sK + pl * kTileBytes, &tmQKV, ld_full, C + h * 64, g * NPT, 0, b, pl);
          tma_load_5d(sV + pl * kTileBytes, &tmQKV, ld_full, 2 * C + h * 64, g * NPT, 0, b, pl);
        }
        mbar_wait(ld_full, ph_ld); ph_ld ^= 1;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                        // S = Q K^T, split precision
          const uint64_t qh = umma_desc_k_sw128(aQ + k * 32), ql = umma_desc_k_sw128(aQ + kTileBytes + k * 32);
          const uint64_t kh = umma_desc_k_sw128(aK + k * 32), kl = umma_desc_k_sw128(aK + kTileBytes + k * 32);
          umma_f16(tmem_base + kS, qh, kh, idesc_s, k != 0);
          umma_f16(tmem_base + kS, ql, kh, idesc_s, 1);
          umma_f16(tmem_base + kS, qh, kl, idesc_s, 1);
        }
        umma_commit(mma_done);
        mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                               // P (masked, unnormalised) in TMEM
        tc_fence_after();
#pragma unroll 1
        for (int kk = 0; kk < 8; ++kk) {                                     // O = P V
          const uint64_t vh = umma_desc_mn_sw128(aV + kk * 2048, 1024, 1024);
          const uint64_t vl = umma_desc_mn_sw128(aV + kTileBytes + kk * 2048, 1024, 1024);
          umma_f16_ts(tmem_base + kO, tmem_base + kS + kk * 16, vh, idesc_o, kk != 0);
          umma_f16_ts(tmem_base + kO, tmem_base + kS + kk * 16 + 8, vh, idesc_o, 1);
          umma_f16_ts(tmem_base + kO, tmem_base + kS + kk * 16, vl, idesc_o, 1);
        }
        umma_commit(mma_done);
        mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                               // O has left TMEM; the operand tiles are free
        tc_fence_after();
      }
    }
  } else if (warp >= 2) {
    const int wq = warp & 3;                               // TMEM lane quarter = warp id % 4
    const int row = wq * 32 + lane;                        // tile row = t * NPT + n
    const int t = row / NPT, nl = row % NPT;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t tS = tmem_base + kS + lane_off, tO = tmem_base + kO + lane_off;
    const float c2 = p.scale_log2e;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int g = item % p.groups, h = (item / p.groups) % p.heads, b = item / (p.groups * p.heads);
      mbar_wait(mma_done, ph); ph ^= 1;
      tc_fence_after();
      uint32_t r[32];
      // pass 1: max over the T entries of this row's token (columns c with c mod NPT == nl)
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        tmem_ld_32x32b_x32(tS + c0, r);
        tmem_ld_wait();
        const int cb = c0 % NPT;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (((cb + j) % NPT) == nl) mx = fmaxf(mx, __uint_as_float(r[j]));
      }
      const float mb = mx * c2;
      // pass 2: p = exp2(s c - max c) on the token's columns, 0 elsewhere; written back as fp16 hi/lo pairs
      float sum = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 16) {
        uint32_t r16[16], pk[16];
        tmem_ld_32x32b_x16(tS + c0, r16);
        tmem_ld_wait();
        const int cb = c0 % NPT;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float p0 = (((cb + j) % NPT) == nl) ? ex2f(__uint_as_float(r16[j]) * c2 - mb) : 0.f;
          const float p1 = (((cb + j + 1) % NPT) == nl) ? ex2f(__uint_as_float(r16[j + 1]) * c2 - mb) : 0.f;
          sum += p0 + p1;
          const __half2 h2 = __floats2half2_rn(p0, p1);
          const float2 hf = __half22float2(h2);
          const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
          pk[8 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        tmem_st_32x32b_x16(tS + c0, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ew_done);
      const float inv = 1.0f / sum;
      if (p.lse != nullptr && g * NPT + nl < p.ntok)
        p.lse[(((long long)b * p.T + t) * p.ntok + g * NPT + nl) * p.heads + h] = mb + log2f(sum);
      mbar_wait(mma_done, ph); ph ^= 1;
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld_32x32b_x32(tO, o0);
      tmem_ld_32x32b_x32(tO + 32, o1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ew_done);
      const int n = g * NPT + nl;
      if (n < p.ntok && t < p.T) {
        const long long grow = ((long long)b * p.T + t) * p.ntok + n;
        if (p.out_f32) {
          float4* o = reinterpret_cast<float4*>(p.out_f32 + grow * C + h * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = make_float4(__uint_as_float(o0[4 * j]) * inv, __uint_as_float(o0[4 * j + 1]) * inv,
                               __uint_as_float(o0[4 * j + 2]) * inv, __uint_as_float(o0[4 * j + 3]) * inv);
            o[8 + j] = make_float4(__uint_as_float(o1[4 * j]) * inv, __uint_as_float(o1[4 * j + 1]) * inv,
                                   __uint_as_float(o1[4 * j + 2]) * inv, __uint_as_float(o1[4 * j + 3]) * inv);
          }
        }
        if (p.out_hi) {
          __half* oh = p.out_hi + grow * C + h * 64;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t* src = half ? o1 : o0;
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float a = __uint_as_float(src[j]) * inv, c = __uint_as_float(src[j + 1]) * inv;
              const __half2 h2 = __floats2half2_rn(a, c);
              const float2 hf = __half22float2(h2);
              const __half2 l2 = __floats2half2_rn(a - hf.x, c - hf.y);
              hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
              lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              reinterpret_cast<uint4*>(oh + half * 32)[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              if (p.out_plane)
                reinterpret_cast<uint4*>(oh + half * 32 + p.out_plane)[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// qkv planes [B*T*ntok, 3C] as a 5-D tensor {column, token, frame, clip, plane} with a {64, NPT, T, 1, 1} box
int make_temporal_tmap(CUtensorMap* tm, const __half* base, long long plane, int B, int T, int ntok, int ld, int npt) {
  const uint64_t dims[5] = {(uint64_t)ld, (uint64_t)ntok, (uint64_t)T, (uint64_t)B, 2};
  const uint64_t str[4] = {(uint64_t)ld * 2, (uint64_t)ntok * ld * 2, (uint64_t)T * ntok * ld * 2, (uint64_t)plane * 2};
  const uint32_t box[5] = {64, (uint32_t)npt, (uint32_t)T, 1, 1};
  return make_tmap_f16(tm, base, 5, dims, str, box);
}

template <int NPT>
int launch_temporal_fwd(const CUtensorMap& tm, const TemporalParams& p, cudaStream_t st) {
  const size_t smem = 1024 + 6 * (size_t)kTileBytes + 64;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_temporal_tc_kernel<NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int items = p.B * p.heads * p.groups;
  const int grid = items < 2 * sm_count() ? items : 2 * sm_count();
  attn_temporal_tc_kernel<NPT><<<grid, kFwdThreads, smem, st>>>(tm, p);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace

bool attn_temporal_tc_supported(int T, long long qkv_plane) { return (T == 4 || T == 8 || T == 16 || T == 32) && qkv_plane != 0; }

int attn_temporal_tc(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale, float* out_f32,
                     __half* out_hi, long long out_plane, cudaStream_t st, float* lse) {
  MAED_CHECK_ARG(attn_temporal_tc_supported(T, qkv_plane), "attn_temporal_tc: T=%d unsupported (4, 8, 16, 32; split precision)", T);
  MAED_CHECK_ARG(qkv_hi && (out_f32 || out_hi), "attn_temporal_tc: null argument");
  const int npt = 128 / T, ld = 3 * heads * 64;
  MAED_CHECK_ARG(qkv_plane >= (long long)B * T * ntok * ld && qkv_plane % 8 == 0, "attn_temporal_tc: qkv planes overlap / misaligned");
  CUtensorMap tm;
  MAED_PROPAGATE(make_temporal_tmap(&tm, qkv_hi, qkv_plane, B, T, ntok, ld, npt));
  TemporalParams p;
  p.B = B; p.T = T; p.ntok = ntok; p.heads = heads; p.groups = (ntok + npt - 1) / npt;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out_f32 = out_f32; p.out_hi = out_hi; p.out_plane = out_plane; p.lse = lse;
  switch (npt) {
    case 32: return launch_temporal_fwd<32>(tm, p, st);
    case 16: return launch_temporal_fwd<16>(tm, p, st);
    case 8: return launch_temporal_fwd<8>(tm, p, st);
    default: return launch_temporal_fwd<4>(tm, p, st);
  }
}

}  // namespace maed
