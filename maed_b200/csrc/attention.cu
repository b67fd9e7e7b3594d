// STE attention kernels (reference lib/models/vision_transformer.py:206-228, 180-204).
//
//  * attn_spatial_tc_kernel — per (frame, head) softmax(Q K^T * scale) V over the 197 tokens of a frame, fused on
//    the 5th-gen tensor cores: Q/K/V tiles arrive by TMA straight from the qkv GEMM output (no permute /
//    contiguous copies), S = Q K^T accumulates in TMEM, the softmax warps read S from TMEM, write P back
//    IN PLACE as fp16 hi/lo pairs (tcgen05.st), and P V runs with P as the TMEM A-operand; the 197x197
//    score matrix never touches shared or global memory (the reference materialises 238 MB of it, 3x).
//  * attn_temporal_kernel — T x T attention across the frames of a clip for every (clip, head, token):
//    18 912 tiny problems; CUDA-core fp32, one warp each, rows gathered with 128-byte coalesced loads
//    straight from the (B*T, HW, C) qkv layout (this is the "(B*T,HW,C) <-> (B*HW,T,C) reshape", done by
//    addressing instead of by copies).
//  * attn_generic_kernel — fp32 CUDA-core flash-style attention over arbitrary sequence length: 'coupling'
//    mode (T*197 tokens) and the test cross-check of the tensor-core kernel.
#include <stdlib.h>

#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

#define LAUNCH_CHECK()                      \
  do {                                      \
    count_launch();                         \
    MAED_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

// =================================================================================== spatial (tcgen05)
static constexpr int kHeadDim = 64;
__device__ __forceinline__ float fast_exp2(float x) {   // one MUFU.EX2 (inputs are <= 0 here; flush-to-zero is fine)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t n) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
}
static constexpr int kQRows = 208;        // query rows fetched; the second M=128 tile reads past them into the next buffer
                                          // (rows >= ntok only feed accumulator rows that are never stored)
static constexpr int kKvRows = 208;       // keys padded to a multiple of 16 (UMMA N / K granularity)
static constexpr int kSpThreads = 640;    // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 / 12-19 softmax groups (2 threads per row)
static constexpr int kColSplit = 112;     // key columns [0,112) -> half 0, [112,208) -> half 1

struct SpatialParams {
  int BT, ntok, heads, nplanes;           // nplanes 2 (split precision) or 1
  float scale_log2e;                      // scale * log2(e)
  float* out_f32;
  __half* out_hi;
  long long out_plane;
  int ldo;                                // heads * 64
  int f32_tma;                            // fp32 output leaves by TMA store (out_f32 only, no planes requested)
  float* lse;                             // optional [BT*ntok, heads]: log2-domain log-sum-exp of every query row (training tape)
  int direct_store;                       // debug knob (MAED_B200_ATTN_DIRECT=1): per-thread global stores instead of TMA
  long long* dbg;                         // optional clock64 timeline of CTA 0 ([item][32] slots), MAED_B200_ATTN_DBG=1
};

#define ATTN_STAMP(slot)                                                                   \
  do {                                                                                     \
    if (p.dbg != nullptr && blockIdx.x == 0 && it < 16) p.dbg[it * 32 + (slot)] = clock64(); \
  } while (0)

__global__ void __launch_bounds__(kSpThreads, 1)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                       const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                       const SpatialParams p) {
  using namespace sm100;
  constexpr uint32_t kQBytes = kQRows * 128;     // per plane
  constexpr uint32_t kKVBytes = kKvRows * 128;   // per plane
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kOStage = 128 * 128;        // one plane of one query tile
  constexpr uint32_t kS0 = 0, kS1 = kKvRows, kO = 2 * kKvRows;   // TMEM column map: S0 | S1 | O

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic ST/LD)
  const int np = p.nplanes;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + np * kQBytes;
  uint8_t* sV = sK + np * kKVBytes;
  uint8_t* sO = sV + np * kKVBytes;                   // output staging [tile][plane] 128 rows x 128 B, 128-byte swizzle
  uint64_t* bars = reinterpret_cast<uint64_t*>(sO + 4 * kOStage);
  uint64_t* qk_full = bars + 0;
  uint64_t* v_full = bars + 1;
  uint64_t* qk_empty = bars + 2;
  uint64_t* v_empty = bars + 3;
  uint64_t* s_full = bars + 4;      // [2]
  uint64_t* p_full = bars + 6;      // [2]
  uint64_t* o_full = bars + 8;      // [2]
  uint64_t* o_empty = bars + 10;    // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 12);
  float* xch = reinterpret_cast<float*>(bars + 16);   // [tile][half][row]{max, sum} exchanged between the two threads of a row

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int items = p.BT * p.heads;
  // softmax warps that own at least one valid query row (rows >= ntok are never stored)
  const int warps_g0 = min(4, (p.ntok + 31) / 32);
  const int warps_g1 = max(0, min(4, (p.ntok - 128 + 31) / 32));

  if (warp == 0 && elect_one()) { prefetch_tmap(&tmQ); prefetch_tmap(&tmKV); }
  if (warp == 1 && elect_one()) {
    mbar_init(qk_full, 1); mbar_init(v_full, 1); mbar_init(qk_empty, 1); mbar_init(v_empty, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&o_full[g], 1);
      mbar_init(&p_full[g], 2 * (g == 0 ? warps_g0 : max(warps_g1, 1)));
      mbar_init(&o_empty[g], 2 * (g == 0 ? warps_g0 : max(warps_g1, 1)));
    }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_base_ptr, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  const bool two_tiles = warps_g1 > 0;

  if (warp == 0) {
    if (elect_one()) {
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const int bt = item / p.heads, h = item % p.heads;
        const int row0 = bt * p.ntok;
        const int ldq = p.heads * kHeadDim;                    // q | k | v column blocks
        mbar_wait(qk_empty, (it & 1) ^ 1);
        ATTN_STAMP(28);
        mbar_arrive_expect_tx(qk_full, np * (kQBytes + kKVBytes));
        for (int pl = 0; pl < np; ++pl) {
          tma_load_3d(sQ + pl * kQBytes, &tmQ, qk_full, h * kHeadDim, row0, pl);
          tma_load_3d(sK + pl * kKVBytes, &tmKV, qk_full, ldq + h * kHeadDim, row0, pl);
        }
        mbar_wait(v_empty, (it & 1) ^ 1);
        ATTN_STAMP(29);
        mbar_arrive_expect_tx(v_full, np * kKVBytes);
        for (int pl = 0; pl < np; ++pl)
          tma_load_3d(sV + pl * kKVBytes, &tmKV, v_full, 2 * ldq + h * kHeadDim, row0, pl);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, kKvRows, 0, 0, 0);   // Q K^T : both K-major
      constexpr uint32_t idesc_o = umma_idesc_f16(128, kHeadDim, 0, 0, 1);  // P V   : B (=V) MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      auto issue_qk = [&](int g) {                 // S_g = Q_g K^T
        const uint32_t d = tmem_base + (g ? kS1 : kS0);
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k) {
          const uint64_t dq = umma_desc_k_sw128(aQ + g * (128 * 128) + k * 32);
          const uint64_t dk = umma_desc_k_sw128(aK + k * 32);
          umma_f16(d, dq, dk, idesc_s, k != 0);
          if (np == 2) {
            const uint64_t dql = umma_desc_k_sw128(aQ + kQBytes + g * (128 * 128) + k * 32);
            const uint64_t dkl = umma_desc_k_sw128(aK + kKVBytes + k * 32);
            umma_f16(d, dql, dk, idesc_s, 1);
            umma_f16(d, dq, dkl, idesc_s, 1);
          }
        }
        umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g) {                 // O = P_g V, P read from TMEM in place of S_g
        const uint32_t a_p = tmem_base + (g ? kS1 : kS0);
        const uint32_t d = tmem_base + kO;
#pragma unroll 1
        for (int kk = 0; kk < kKvRows / 16; ++kk) {
          // V rows [16kk, 16kk+16): two 8-row swizzle atoms 1024 B apart; N = 64 fits one 128-byte atom row
          const uint64_t dv = umma_desc_mn_sw128(aV + kk * 2048, 1024, 1024);
          umma_f16_ts(d, a_p + kk * 16, dv, idesc_o, kk != 0);
          if (np == 2) {
            const uint64_t dvl = umma_desc_mn_sw128(aV + kKVBytes + kk * 2048, 1024, 1024);
            umma_f16_ts(d, a_p + kk * 16 + 8, dv, idesc_o, 1);       // P_lo * V_hi
            umma_f16_ts(d, a_p + kk * 16, dvl, idesc_o, 1);          // P_hi * V_lo
          }
        }
        umma_commit(&o_full[g]);
      };
      if (two_tiles) {
        // Software-pipelined issue order  PV0(i), QK0(i+1), PV1(i), QK1(i+1): the two softmax groups run half a period
        // apart, so the tensor pipe always has the other group's MMAs to execute while one group is in its softmax.
        uint32_t it = 0;
        int item = blockIdx.x;
        if (item < items) {
          mbar_wait(qk_full, 0);
          tc_fence_after();
          issue_qk(0);
          issue_qk(1);
          umma_commit(qk_empty);
        }
        for (; item < items; item += gridDim.x, ++it) {
          const uint32_t ph = it & 1;
          const bool has_next = item + (int)gridDim.x < items;
          mbar_wait(v_full, ph);
          ATTN_STAMP(0);
          mbar_wait(&p_full[0], ph);
          ATTN_STAMP(1);
          mbar_wait(&o_empty[1], ph ^ 1);
          ATTN_STAMP(2);
          tc_fence_after();
          issue_pv(0);
          ATTN_STAMP(3);
          if (has_next) {
            mbar_wait(qk_full, ph ^ 1);
            ATTN_STAMP(4);
            tc_fence_after();
            issue_qk(0);
          }
          ATTN_STAMP(5);
          mbar_wait(&p_full[1], ph);
          ATTN_STAMP(6);
          mbar_wait(&o_empty[0], ph);
          ATTN_STAMP(7);
          tc_fence_after();
          issue_pv(1);
          umma_commit(v_empty);
          if (has_next) {
            issue_qk(1);
            umma_commit(qk_empty);
          }
          ATTN_STAMP(8);
        }
      } else {
        uint32_t it = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
          const uint32_t ph = it & 1;
          mbar_wait(qk_full, ph);
          tc_fence_after();
          issue_qk(0);
          umma_commit(qk_empty);
          mbar_wait(v_full, ph);
          mbar_wait(&p_full[0], ph);
          mbar_wait(&o_empty[0], ph ^ 1);
          tc_fence_after();
          issue_pv(0);
          umma_commit(v_empty);
        }
      }
    }
  } else if (warp >= 4) {
    // ============================================================ softmax + epilogue warps
    // Two threads per query row: warps (g, half, wq) — half 0 owns key columns [0,112), half 1 owns [112,208) and,
    // in the epilogue, output columns [32*half, 32*half+32).  The halves exchange row max and row sum through
    // shared memory under a 64-thread named barrier, so the softmax and the epilogue each take half as long and
    // every SM sub-partition has 4 softmax warps to hide TMEM-load / MUFU latency behind.
    const int g = (warp - 4) >> 3;                 // query tile
    const int half = ((warp - 4) >> 2) & 1;        // key-column half
    const int wq = warp & 3;                       // TMEM lane quarter (hardware: warp id % 4)
    const bool active = g == 0 ? (wq < warps_g0) : (wq < warps_g1);
    if (active) {
      const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
      const uint32_t tS = tmem_base + (g ? kS1 : kS0) + lane_off;
      const uint32_t tO = tmem_base + kO + lane_off;
      const int qrow = g * 128 + wq * 32 + lane;   // token index of this thread's query row
      const int trow = wq * 32 + lane;
      float* x_mine = xch + ((g * 2 + half) * 128 + trow) * 2;          // {max, sum}
      const float* x_peer = xch + ((g * 2 + (half ^ 1)) * 128 + trow) * 2;
      const uint32_t pair_bar = 1 + g * 4 + wq;
      const int cbeg = half ? kColSplit : 0;
      const int nbatch = half ? (kKvRows - kColSplit) / 32 : (kColSplit + 31) / 32;   // 3 | 4 (last of half 0: 16 cols)
      uint32_t it = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++it) {
        const uint32_t ph = it & 1;
        const int bt = item / p.heads, h = item % p.heads;
        mbar_wait(&s_full[g], ph);
        tc_fence_after();
#define SM_STAMP(k) do { if (wq == 0 && half == 0 && lane == 0) ATTN_STAMP(10 + 8 * g + (k)); } while (0)
        SM_STAMP(0);
        uint32_t r[32];
        // pass 1: row max of the raw scores over the valid keys of this half
        float mx = -INFINITY;
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
          const int col0 = cbeg + b * 32;
          const int ncols = min(32, (half ? kKvRows : kColSplit) - col0);
          if (ncols == 32) tmem_ld_32x32b_x32(tS + col0, r);
          else tmem_ld_32x32b_x16(tS + col0, reinterpret_cast<uint32_t(&)[16]>(r[0]));
          tmem_ld_wait();
          if (col0 + 32 <= p.ntok && ncols == 32) {           // whole batch valid: no per-column predicates
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols && col0 + j < p.ntok) mx = fmaxf(mx, __uint_as_float(r[j]));
          }
        }
        SM_STAMP(1);
        x_mine[0] = mx;
        named_bar_sync(pair_bar, 64);
        SM_STAMP(2);
        mx = fmaxf(mx, x_peer[0]);
        const float mb = mx * p.scale_log2e;
        // pass 2: p = exp2(s*scale*log2e - max*scale*log2e); P written back in place, per 16 columns: 8 packed hi | 8 packed lo
        float sum = 0.f;
#pragma unroll 1
        for (int b = 0; b < nbatch; ++b) {
          const int col0 = cbeg + b * 32;
          const int ncols = min(32, (half ? kKvRows : kColSplit) - col0);
          if (ncols == 32) tmem_ld_32x32b_x32(tS + col0, r);
          else tmem_ld_32x32b_x16(tS + col0, reinterpret_cast<uint32_t(&)[16]>(r[0]));
          tmem_ld_wait();
          const bool all_valid = col0 + 32 <= p.ntok;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (c * 16 < ncols) {
              uint32_t pk[16];
#pragma unroll
              for (int jj = 0; jj < 16; jj += 2) {
                const int col = col0 + c * 16 + jj;
                float p0 = fast_exp2(__uint_as_float(r[c * 16 + jj]) * p.scale_log2e - mb);
                float p1 = fast_exp2(__uint_as_float(r[c * 16 + jj + 1]) * p.scale_log2e - mb);
                if (!all_valid) {
                  if (col >= p.ntok) p0 = 0.f;
                  if (col + 1 >= p.ntok) p1 = 0.f;
                }
                sum += p0 + p1;
                const __half2 h2 = __floats2half2_rn(p0, p1);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(p0 - hf.x, p1 - hf.y);
                pk[jj >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                pk[8 + (jj >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              tmem_st_32x32b_x16(tS + col0 + c * 16, pk);
            }
          }
        }
        x_mine[1] = sum;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[g]);
        SM_STAMP(3);
        named_bar_sync(pair_bar, 64);
        const float inv = 1.0f / (sum + x_peer[1]);
        if (p.lse != nullptr && half == 0 && qrow < p.ntok)      // the backward recomputes P = exp2(s c - lse) in one pass
          p.lse[((long long)bt * p.ntok + qrow) * p.heads + h] = mb + log2f(sum + x_peer[1]);
        // epilogue: O / sum.  O is pulled into registers and released at once (the other tile's P V may start);
        // the fp16 hi/lo planes go through a swizzled staging tile and leave by TMA store, so the softmax warps
        // never wait on global-memory back-pressure.
        mbar_wait(&o_full[g], ph);
        tc_fence_after();
        SM_STAMP(4);
        const long long orow = (long long)bt * p.ntok + qrow;
        tmem_ld_32x32b_x32(tO + half * 32, r);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_empty[g]);
        SM_STAMP(5);
        if (p.f32_tma) {
          // fp32 output: each (tile, half) staging buffer is one 32-float-wide box (128 B rows, 128-byte swizzle)
          const bool store_thread = wq == 0 && half == 0 && lane == 0;
          const uint32_t group_threads = 64u * (g == 0 ? warps_g0 : warps_g1);
          if (store_thread) tma_store_wait_read<0>();          // previous item's tile has left the staging buffers
          named_bar_sync(9 + 2 * g, group_threads);
          uint8_t* st_row = sO + (2 * g + half) * kOStage + trow * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(st_row + ((j ^ (trow & 7)) * 16)) =
                make_float4(__uint_as_float(r[4 * j]) * inv, __uint_as_float(r[4 * j + 1]) * inv,
                            __uint_as_float(r[4 * j + 2]) * inv, __uint_as_float(r[4 * j + 3]) * inv);
          fence_proxy_async();
          named_bar_sync(10 + 2 * g, group_threads);
          if (store_thread) {
            const CUtensorMap* tmO = g ? &tmO1 : &tmO0;       // fp32 matrix described as 2x as many 16-bit columns
            const int row_g = bt * p.ntok + g * 128;
            tma_store_3d(tmO, sO + (2 * g) * kOStage, 2 * (h * kHeadDim), row_g, 0);
            tma_store_3d(tmO, sO + (2 * g + 1) * kOStage, 2 * (h * kHeadDim + 32), row_g, 0);
            tma_store_commit();
          }
        } else if (p.out_f32 && qrow < p.ntok) {
          float4* op = reinterpret_cast<float4*>(p.out_f32 + orow * p.ldo + h * kHeadDim + half * 32);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            op[j >> 2] = make_float4(__uint_as_float(r[j]) * inv, __uint_as_float(r[j + 1]) * inv,
                                     __uint_as_float(r[j + 2]) * inv, __uint_as_float(r[j + 3]) * inv);
        }
        if (p.out_hi) {
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float a = __uint_as_float(r[j]) * inv, b = __uint_as_float(r[j + 1]) * inv;
            const __half2 h2 = __floats2half2_rn(a, b);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(a - hf.x, b - hf.y);
            hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          if (p.direct_store) {
            if (qrow < p.ntok) {
              __half* oh = p.out_hi + orow * p.ldo + h * kHeadDim + half * 32;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                reinterpret_cast<uint4*>(oh)[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
                reinterpret_cast<uint4*>(oh + p.out_plane)[j] =
                    make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
              }
            }
            SM_STAMP(6);
            continue;
          }
          const bool store_thread = wq == 0 && half == 0 && lane == 0;
          const uint32_t group_threads = 64u * (g == 0 ? warps_g0 : warps_g1);
          if (store_thread) tma_store_wait_read<0>();          // previous item's tile has left the staging buffer
          named_bar_sync(9 + 2 * g, group_threads);
          uint8_t* st_hi = sO + (2 * g) * kOStage + trow * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = ((half * 4 + j) ^ (trow & 7)) * 16;
            *reinterpret_cast<uint4*>(st_hi + chunk) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
            *reinterpret_cast<uint4*>(st_hi + kOStage + chunk) =
                make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
          }
          fence_proxy_async();
          named_bar_sync(10 + 2 * g, group_threads);
          if (store_thread) {
            const CUtensorMap* tmO = g ? &tmO1 : &tmO0;
            const int row_g = bt * p.ntok + g * 128;
            tma_store_3d(tmO, sO + (2 * g) * kOStage, h * kHeadDim, row_g, 0);
            tma_store_3d(tmO, sO + (2 * g + 1) * kOStage, h * kHeadDim, row_g, 1);
            tma_store_commit();
          }
        }
        SM_STAMP(6);
#undef SM_STAMP
      }
      if (wq == 0 && half == 0 && lane == 0) tma_store_wait_all();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

int attn_spatial(const __half* qkv_hi, long long qkv_plane, int BT, int ntok, int heads, float scale, int nsplit,
                 float* out_f32, __half* out_hi, long long out_plane, cudaStream_t st, float* lse) {
  MAED_CHECK_ARG(ntok >= 1 && ntok <= kKvRows, "attn_spatial: ntok=%d unsupported (1..%d)", ntok, kKvRows);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "attn_spatial: nsplit must be 1 or 3");
  const int np = nsplit == 3 ? 2 : 1;
  const int ld = 3 * heads * kHeadDim;
  const long long rows = (long long)BT * ntok;
  CUtensorMap tmQ, tmKV;
  const uint64_t dims[3] = {(uint64_t)ld, (uint64_t)rows, (uint64_t)np};
  const uint64_t str[2] = {(uint64_t)ld * 2, (uint64_t)(np == 2 ? qkv_plane : rows * ld) * 2};
  const uint32_t boxq[3] = {64, kQRows, 1};
  const uint32_t boxkv[3] = {64, kKvRows, 1};
  MAED_PROPAGATE(make_tmap_f16(&tmQ, qkv_hi, 3, dims, str, boxq));
  MAED_PROPAGATE(make_tmap_f16(&tmKV, qkv_hi, 3, dims, str, boxkv));
  // output planes leave by TMA store: one box per query tile (tile 1 holds only ntok - 128 rows of the frame)
  CUtensorMap tmO0 = tmKV, tmO1 = tmKV;
  if (out_hi) {
    MAED_CHECK_ARG(out_plane >= rows * heads * kHeadDim && (out_plane % 8) == 0 &&
                       (reinterpret_cast<uintptr_t>(out_hi) & 15) == 0,
                   "attn_spatial: output planes must be 16-byte aligned and at least rows*heads*64 apart");
    const uint64_t odims[3] = {(uint64_t)heads * kHeadDim, (uint64_t)rows, 2};
    const uint64_t ostr[2] = {(uint64_t)heads * kHeadDim * 2, (uint64_t)out_plane * 2};
    const uint32_t box0[3] = {64, (uint32_t)(ntok < 128 ? ntok : 128), 1};
    MAED_PROPAGATE(make_tmap_f16(&tmO0, out_hi, 3, odims, ostr, box0));
    if (ntok > 128) {
      const uint32_t box1[3] = {64, (uint32_t)(ntok - 128), 1};
      MAED_PROPAGATE(make_tmap_f16(&tmO1, out_hi, 3, odims, ostr, box1));
    }
  }
  static const bool f32_tma_on = getenv("MAED_B200_ATTN_DIRECT") == nullptr;   // debug knob: per-thread global stores
  const bool f32_tma = f32_tma_on && out_f32 && !out_hi && (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0;
  if (f32_tma) {
    const uint64_t odims[3] = {(uint64_t)heads * kHeadDim * 2, (uint64_t)rows, 1};
    const uint64_t ostr[2] = {(uint64_t)heads * kHeadDim * 4, (uint64_t)rows * heads * kHeadDim * 4};
    const uint32_t box0[3] = {64, (uint32_t)(ntok < 128 ? ntok : 128), 1};
    MAED_PROPAGATE(make_tmap_f16(&tmO0, out_f32, 3, odims, ostr, box0));
    if (ntok > 128) {
      const uint32_t box1[3] = {64, (uint32_t)(ntok - 128), 1};
      MAED_PROPAGATE(make_tmap_f16(&tmO1, out_f32, 3, odims, ostr, box1));
    }
  }
  SpatialParams p;
  p.f32_tma = f32_tma ? 1 : 0;
  p.lse = lse;
  p.BT = BT; p.ntok = ntok; p.heads = heads; p.nplanes = np;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out_f32 = out_f32; p.out_hi = out_hi; p.out_plane = out_plane; p.ldo = heads * kHeadDim;
  const size_t smem = 1024 + (size_t)np * (kQRows * 128 + 2 * kKvRows * 128) + 4 * 128 * 128 + 256 +
                      2 * 2 * 128 * 2 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_spatial_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int items = BT * heads;
  const int grid = items < sm_count() ? items : sm_count();
  static const bool direct = getenv("MAED_B200_ATTN_DIRECT") != nullptr;
  p.direct_store = direct ? 1 : 0;
  static const bool dbg_on = getenv("MAED_B200_ATTN_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  p.dbg = nullptr;
  if (dbg_on) {
    if (!dbg_buf) MAED_CUDA_CHECK(cudaMalloc(&dbg_buf, 16 * 32 * sizeof(long long)));
    MAED_CUDA_CHECK(cudaMemsetAsync(dbg_buf, 0, 16 * 32 * sizeof(long long), st));
    p.dbg = dbg_buf;
  }
  attn_spatial_tc_kernel<<<grid, kSpThreads, smem, st>>>(tmQ, tmKV, tmO0, tmO1, p);
  LAUNCH_CHECK();
  if (dbg_on) {   // debug only: per-item event times of CTA 0, cycles relative to the first stamp
    static long long h[16 * 32];
    MAED_CUDA_CHECK(cudaStreamSynchronize(st));
    MAED_CUDA_CHECK(cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (int i = 0; i < 16 * 32; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    for (int i = 0; i < 16; ++i) {
      fprintf(stderr, "ATTN_DBG item %2d:", i);
      for (int sl = 0; sl < 32; ++sl) fprintf(stderr, " %lld", h[i * 32 + sl] ? h[i * 32 + sl] - t0 : -1LL);
      fprintf(stderr, "\n");
    }
  }
  return MAED_OK;
}

// ======================================================================================== temporal
// One warp per (clip, head, token).  The T K-rows and V-rows (64 halfs = 128 B each, hi + lo planes) are fetched
// with 16-byte loads (8 lanes per row, 4 rows per instruction), summed to fp32 and staged in shared memory; lane
// (t, seg) then owns query frame t and every SEGS-th float4 of the head dimension, so the inner products read
// shared memory with conflict-free 128-bit loads (one LDS per 4 FMAs).
__device__ __forceinline__ void cvt8(const uint4& hi, const uint4& lo, bool has_lo, float* out) {
  const __half2* h = reinterpret_cast<const __half2*>(&hi);
  const __half2* l = reinterpret_cast<const __half2*>(&lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    if (has_lo) { const float2 g = __half22float2(l[i]); f.x += g.x; f.y += g.y; }
    out[2 * i] = f.x; out[2 * i + 1] = f.y;
  }
}

template <int SEGS, int TMAX>
__global__ void attn_temporal_kernel(const __half* __restrict__ qkv, long long plane, int B, int T, int ntok, int heads,
                                     float scale, float* __restrict__ out_f32, __half* __restrict__ out_hi,
                                     long long out_plane, long long total) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NJ = 16 / SEGS;                           // float4 chunks of the head dim owned by a lane
  const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sK = sm + (size_t)w * 2 * TMAX * kHeadDim;
  float* sV = sK + TMAX * kHeadDim;
  const int ld = 3 * heads * kHeadDim, C = heads * kHeadDim;
  const bool has_lo = plane != 0;
  const int lrow = lane >> 3, lchunk = lane & 7;
  for (long long item = (long long)blockIdx.x * warps + w; item < total; item += (long long)gridDim.x * warps) {
    // heads fastest: the warps of a block read ADJACENT 128-byte head slices of the same token rows, i.e. contiguous
    // 1.5 KB runs of the qkv matrix (round 2; with tokens fastest the runs were 128 B every 4.6 KB: 30 % of HBM peak)
    const int h = (int)(item % heads);
    const int n = (int)((item / heads) % ntok);
    const int b = (int)(item / ((long long)ntok * heads));
    __syncwarp();
    // ---- stage K, V (all loads of the item issued before the first use)
    uint4 kh[TMAX / 4], kl[TMAX / 4], vh[TMAX / 4], vl[TMAX / 4];
#pragma unroll
    for (int i = 0; i < TMAX / 4; ++i) {
      const int r = lrow + 4 * i;
      kh[i] = kl[i] = vh[i] = vl[i] = make_uint4(0, 0, 0, 0);
      if (r < T) {
        const __half* kp = qkv + (((long long)b * T + r) * ntok + n) * ld + C + h * kHeadDim + lchunk * 8;
        kh[i] = *reinterpret_cast<const uint4*>(kp);
        vh[i] = *reinterpret_cast<const uint4*>(kp + C);
        if (has_lo) {
          kl[i] = *reinterpret_cast<const uint4*>(kp + plane);
          vl[i] = *reinterpret_cast<const uint4*>(kp + C + plane);
        }
      }
    }
    const int t = (SEGS == 1) ? lane : (lane % T);
    const int seg = (SEGS == 1) ? 0 : (lane / T);
    const bool act = t < T;
    const long long qrow = ((long long)b * T + (act ? t : 0)) * ntok + n;
    float q[NJ * 4];
    {
      const __half* qp = qkv + qrow * ld + h * kHeadDim;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const int d = 4 * (i * SEGS + seg);
        const uint2 a = *reinterpret_cast<const uint2*>(qp + d);
        float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
        float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
        if (has_lo) {
          const uint2 c = *reinterpret_cast<const uint2*>(qp + d + plane);
          const float2 g0 = __half22float2(*reinterpret_cast<const __half2*>(&c.x));
          const float2 g1 = __half22float2(*reinterpret_cast<const __half2*>(&c.y));
          f0.x += g0.x; f0.y += g0.y; f1.x += g1.x; f1.y += g1.y;
        }
        q[4 * i] = f0.x * scale; q[4 * i + 1] = f0.y * scale; q[4 * i + 2] = f1.x * scale; q[4 * i + 3] = f1.y * scale;
      }
    }
#pragma unroll
    for (int i = 0; i < TMAX / 4; ++i) {
      const int r = lrow + 4 * i;
      if (r < T) {
        float f[8];
        cvt8(kh[i], kl[i], has_lo, f);
        *reinterpret_cast<float4*>(&sK[r * kHeadDim + lchunk * 8]) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(&sK[r * kHeadDim + lchunk * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
        cvt8(vh[i], vl[i], has_lo, f);
        *reinterpret_cast<float4*>(&sV[r * kHeadDim + lchunk * 8]) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(&sV[r * kHeadDim + lchunk * 8 + 4]) = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
    __syncwarp();
    // ---- scores (q already carries the softmax scale)
    float s[TMAX];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < TMAX; ++u) {
      float acc = 0.f;
      if (u < T) {
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          const float4 k4 = *reinterpret_cast<const float4*>(&sK[u * kHeadDim + 4 * (i * SEGS + seg)]);
          acc += q[4 * i] * k4.x + q[4 * i + 1] * k4.y + q[4 * i + 2] * k4.z + q[4 * i + 3] * k4.w;
        }
      }
      if (SEGS >= 2) acc += __shfl_xor_sync(0xffffffffu, acc, T);
      if (SEGS >= 4) acc += __shfl_xor_sync(0xffffffffu, acc, 2 * T);
      s[u] = (u < T) ? acc : -INFINITY;
      mx = fmaxf(mx, s[u]);
    }
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < TMAX; ++u) { s[u] = (u < T) ? expf(s[u] - mx) : 0.f; sum += s[u]; }
    const float inv = 1.0f / sum;
    float o[NJ * 4];
#pragma unroll
    for (int d = 0; d < NJ * 4; ++d) o[d] = 0.f;
#pragma unroll
    for (int u = 0; u < TMAX; ++u) {
      if (u < T) {
        const float pu = s[u] * inv;
#pragma unroll
        for (int i = 0; i < NJ; ++i) {
          const float4 v4 = *reinterpret_cast<const float4*>(&sV[u * kHeadDim + 4 * (i * SEGS + seg)]);
          o[4 * i] += pu * v4.x; o[4 * i + 1] += pu * v4.y; o[4 * i + 2] += pu * v4.z; o[4 * i + 3] += pu * v4.w;
        }
      }
    }
    if (act) {
      const long long off = qrow * C + h * kHeadDim;
#pragma unroll
      for (int i = 0; i < NJ; ++i) {
        const int d = 4 * (i * SEGS + seg);
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + off + d) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        if (out_hi) {
          const __half2 h0 = __floats2half2_rn(o[4 * i], o[4 * i + 1]), h1 = __floats2half2_rn(o[4 * i + 2], o[4 * i + 3]);
          const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
          const __half2 l0 = __floats2half2_rn(o[4 * i] - f0.x, o[4 * i + 1] - f0.y);
          const __half2 l1 = __floats2half2_rn(o[4 * i + 2] - f1.x, o[4 * i + 3] - f1.y);
          uint2 H, L;
          H.x = *reinterpret_cast<const uint32_t*>(&h0); H.y = *reinterpret_cast<const uint32_t*>(&h1);
          L.x = *reinterpret_cast<const uint32_t*>(&l0); L.y = *reinterpret_cast<const uint32_t*>(&l1);
          *reinterpret_cast<uint2*>(out_hi + off + d) = H;
          *reinterpret_cast<uint2*>(out_hi + off + d + out_plane) = L;
        }
      }
    }
  }
}

int attn_temporal(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale, float* out_f32,
                  __half* out_hi, long long out_plane, cudaStream_t st, float* lse) {
  MAED_CHECK_ARG(T >= 1 && T <= 32, "attn_temporal: T=%d unsupported (1..32)", T);
  MAED_CHECK_ARG(qkv_plane % 8 == 0 && out_plane % 4 == 0, "attn_temporal: plane strides must be 16-byte aligned");
  // tensor-core version (TMA-gathered 128-row tiles); MAED_B200_TEMPORAL_TC=0 selects the CUDA-core kernel below
  static const bool tc_on = [] { const char* v = getenv("MAED_B200_TEMPORAL_TC"); return !(v && v[0] == '0'); }();
  if (tc_on && attn_temporal_tc_supported(T, qkv_plane) && (!out_hi || out_plane % 8 == 0))
    return attn_temporal_tc(qkv_hi, qkv_plane, B, T, ntok, heads, scale, out_f32, out_hi, out_plane, st, lse);
  MAED_CHECK_ARG(lse == nullptr, "attn_temporal: the row statistics are only produced by the tensor-core kernel (T=%d)", T);
  const long long total = (long long)B * heads * ntok;
  const long long cap = (long long)sm_count() * 32;
#define TEMPORAL_LAUNCH(SEGS, TMAX, WARPS)                                                                          \
  do {                                                                                                              \
    const size_t smem = (size_t)(WARPS) * 2 * (TMAX) * kHeadDim * sizeof(float);                                    \
    static bool attr_set = false;                                                                                   \
    if (!attr_set) {                                                                                                \
      MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_temporal_kernel<SEGS, TMAX>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));                    \
      attr_set = true;                                                                                              \
    }                                                                                                               \
    long long blocks = (total + (WARPS) - 1) / (WARPS);                                                             \
    if (blocks > cap) blocks = cap;                                                                                 \
    attn_temporal_kernel<SEGS, TMAX><<<(int)blocks, (WARPS) * 32, smem, st>>>(                                      \
        qkv_hi, qkv_plane, B, T, ntok, heads, scale, out_f32, out_hi, out_plane, total);                            \
  } while (0)
  if (T == 16) TEMPORAL_LAUNCH(2, 16, 4);
  else if (T == 8) TEMPORAL_LAUNCH(4, 8, 4);
  else if (T <= 16) TEMPORAL_LAUNCH(1, 16, 4);
  else TEMPORAL_LAUNCH(1, 32, 4);
#undef TEMPORAL_LAUNCH
  LAUNCH_CHECK();
  return MAED_OK;
}

// ========================================================================================= generic
// grid (ceil(seq/128), heads, batch); thread = one query row; K/V streamed through shared memory in tiles of
// 32 keys; online softmax in fp32.
__global__ void __launch_bounds__(128)
attn_generic_kernel(const __half* __restrict__ qkv, long long plane, int seq, int heads, float scale,
                    float* __restrict__ out_f32, __half* __restrict__ out_hi, long long out_plane) {
  __shared__ float sK[32][kHeadDim];
  __shared__ float sV[32][kHeadDim];
  const int b = blockIdx.z, h = blockIdx.y;
  const int qi = blockIdx.x * 128 + threadIdx.x;
  const int ld = 3 * heads * kHeadDim, C = heads * kHeadDim;
  const long long base = (long long)b * seq;
  float q[kHeadDim], o[kHeadDim];
  const bool act = qi < seq;
  {
    const __half* qp = qkv + (base + (act ? qi : 0)) * ld + h * kHeadDim;
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 2) {
      float2 f = __half22float2(*reinterpret_cast<const __half2*>(qp + d));
      if (plane) { const float2 l = __half22float2(*reinterpret_cast<const __half2*>(qp + d + plane)); f.x += l.x; f.y += l.y; }
      q[d] = f.x * scale; q[d + 1] = f.y * scale;
      o[d] = 0.f; o[d + 1] = 0.f;
    }
  }
  float mx = -INFINITY, sum = 0.f;
  for (int k0 = 0; k0 < seq; k0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 32 * 32; idx += 128) {
      const int r = idx >> 5, c2 = (idx & 31) * 2;
      float2 kf = make_float2(0.f, 0.f), vf = make_float2(0.f, 0.f);
      if (k0 + r < seq) {
        const __half* kp = qkv + (base + k0 + r) * ld + C + h * kHeadDim + c2;
        kf = __half22float2(*reinterpret_cast<const __half2*>(kp));
        vf = __half22float2(*reinterpret_cast<const __half2*>(kp + C));
        if (plane) {
          const float2 kl = __half22float2(*reinterpret_cast<const __half2*>(kp + plane));
          const float2 vl = __half22float2(*reinterpret_cast<const __half2*>(kp + C + plane));
          kf.x += kl.x; kf.y += kl.y; vf.x += vl.x; vf.y += vl.y;
        }
      }
      sK[r][c2] = kf.x; sK[r][c2 + 1] = kf.y; sV[r][c2] = vf.x; sV[r][c2 + 1] = vf.y;
    }
    __syncthreads();
    const int kn = min(32, seq - k0);
    for (int r = 0; r < kn; ++r) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < kHeadDim; ++d) s += q[d] * sK[r][d];
      const float nm = fmaxf(mx, s);
      const float corr = expf(mx - nm), pr = expf(s - nm);
      sum = sum * corr + pr;
#pragma unroll
      for (int d = 0; d < kHeadDim; ++d) o[d] = o[d] * corr + pr * sV[r][d];
      mx = nm;
    }
  }
  if (act) {
    const float inv = 1.0f / sum;
    const long long off = (base + qi) * C + h * kHeadDim;
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 2) {
      const float a = o[d] * inv, c = o[d + 1] * inv;
      if (out_f32) { out_f32[off + d] = a; out_f32[off + d + 1] = c; }
      if (out_hi) {
        const __half2 h2 = __floats2half2_rn(a, c);
        const float2 hf = __half22float2(h2);
        *reinterpret_cast<__half2*>(out_hi + off + d) = h2;
        *reinterpret_cast<__half2*>(out_hi + off + d + out_plane) = __floats2half2_rn(a - hf.x, c - hf.y);
      }
    }
  }
}

int attn_generic(const __half* qkv_hi, long long qkv_plane, int batch, int seq, int heads, float scale, int tokens_per_frame,
                 int frames_per_batch, float* out_f32, __half* out_hi, long long out_plane, cudaStream_t st) {
  (void)tokens_per_frame; (void)frames_per_batch;       // rows of one batch element are contiguous: row = b*seq + i
  attn_generic_kernel<<<dim3(cdiv(seq, 128), heads, batch), 128, 0, st>>>(qkv_hi, qkv_plane, seq, heads, scale, out_f32,
                                                                         out_hi, out_plane);
  LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
