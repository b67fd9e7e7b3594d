// tcgen05 / TMA / TMEM GEMM for sm_100a — the workhorse of the MAED hot path.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )          A, B: fp16, K-major; fp32 accumulation in TMEM
//
// * "split" precision (nsplit = 3): every operand is stored as two fp16 planes (hi = rn(x), lo = rn(x - hi))
//   and each K step issues three MMAs  Ah*Bh + Al*Bh + Ah*Bl  into the same TMEM accumulator (~2^-22 operand
//   precision).  DESIGN.md §precision explains why plain fp16/bf16/tf32 operands cannot meet the reference's
//   1e-3 parity gate (the weight-standardised backbone amplifies rounding error ~80x).  nsplit = 1 is the
//   plain single-MMA fp16 mode.
// * A is either a plain [M,K] matrix (linears, 1x1 convs, explicit im2col) or, in conv mode, an NHWC
//   activation tensor walked tap by tap with 5-D TMA boxes (implicit GEMM for stride-1 kxk convs; zero
//   padding comes from TMA out-of-bounds fill).
// * Warp roles (256 threads): warp 0 TMA producer, warp 1 MMA issuer (one elected thread), warp 2 TMEM
//   allocator, warps 4..7 epilogue (TMEM -> registers -> bias / GELU / residual / fp16-split -> global).
//   Persistent CTAs walk tiles round-robin; TMEM accumulators are double buffered so the epilogue of tile i
//   overlaps the main loop of tile i+1.
#pragma once
#include "sm100_ptx.cuh"

namespace maed {

enum : int { OUT_F32 = 0, OUT_F16 = 1, OUT_F16_SPLIT = 2 };
enum : int { ACT_NONE = 0, ACT_GELU = 1, ACT_RELU = 2, ACT_TANH = 3 };

struct GemmParams {
  int M, N, K;                 // logical sizes (K per plane)
  int num_k_blocks;            // K blocks of 64 (conv: taps * cin_blocks)
  int nsplit;                  // 1 or 3
  int stages;                  // smem pipeline depth
  int m_tiles, n_tiles;
  // epilogue
  const float* bias;           // [N] or nullptr
  const float* residual;       // fp32 [M, ldc] or nullptr
  const __half* res_hi;        // residual given as fp16 planes [M, ldc] (hi; lo at + res_plane when res_plane != 0) or nullptr
  long long res_plane;
  int act;
  int act_post;                // ACT_NONE or ACT_RELU, applied AFTER the residuals (ResNet bottleneck: relu(conv + identity))
  int out_mode;
  void* out;
  long long out_plane_stride;  // elements between the hi and lo output planes (OUT_F16_SPLIT)
  int ldc;
  // conv (implicit GEMM) mode
  int conv;                    // 0 plain, 1 tap mode
  int H, W, cin_blocks, KW, pad_h, pad_w, tile_h, tile_w, tiles_h, tiles_w;
  uint32_t a_tx_bytes;         // bytes one A-plane TMA box delivers (conv boxes cover tile_h*tile_w <= 128 rows)
};

static constexpr int kBlockM = 128;
static constexpr int kBlockK = 64;
static constexpr int kGemmThreads = 256;
static constexpr int kMaxStages = 8;

// out = act_post(act(acc + bias) + residual + residual_planes) for one row x 32 columns
__device__ __forceinline__ void epilogue_math(float (&v)[32], const uint32_t (&r)[32], const GemmParams& p, int col0,
                                              long long out_row) {
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  if (p.bias) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + erff(v[j] * 0.70710678118654752440f));
  } else if (p.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  } else if (p.act == ACT_TANH) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
  }
  if (p.residual) {
    const float4* rp = reinterpret_cast<const float4*>(p.residual + out_row * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      const float4 b = rp[j >> 2];
      v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
    }
  }
  if (p.res_hi) {
    const uint4* rh = reinterpret_cast<const uint4*>(p.res_hi + out_row * p.ldc + col0);
    const uint4* rl = reinterpret_cast<const uint4*>(p.res_hi + p.res_plane + out_row * p.ldc + col0);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint4 h = rh[q];
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[k]));
        v[q * 8 + 2 * k] += f.x; v[q * 8 + 2 * k + 1] += f.y;
      }
      if (p.res_plane) {
        const uint4 l = rl[q];
        const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&lw[k]));
          v[q * 8 + 2 * k] += f.x; v[q * 8 + 2 * k + 1] += f.y;
        }
      }
    }
  }
  if (p.act_post == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
}

// TMA_EPI = true (opt-in: MAED_B200_GEMM_TMA_EPI=1, plain mode only): the epilogue stages every 128 x 32 output chunk in
// shared memory (swizzled like the output tensor map) and one thread writes it with a TMA store, instead of 16-byte
// row-per-thread global stores.  Same arithmetic.  Added (not yet measured) after the spatial-attention kernel gained 20 %
// from the same change (profiles/README.md); the default instantiation keeps the original epilogue.
static constexpr uint32_t kEpiStageBytes = 2 * 16384;       // two 16 KB staging buffers (TMA_EPI only)

template <int BLOCK_N, bool TMA_EPI = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const GemmParams p) {
  using namespace sm100;
  constexpr int kAccStages = (2 * BLOCK_N <= 512) ? 2 : 1;
  constexpr int kTmemCols = (kAccStages * BLOCK_N <= 32) ? 32 : (kAccStages * BLOCK_N <= 64) ? 64
                          : (kAccStages * BLOCK_N <= 128) ? 128 : (kAccStages * BLOCK_N <= 256) ? 256 : 512;
  constexpr uint32_t kABytes = kBlockM * kBlockK * 2;     // 16 KB
  constexpr uint32_t kBBytes = BLOCK_N * kBlockK * 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem0 = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic ST/LD)
  uint8_t* smem = smem0 + (TMA_EPI ? kEpiStageBytes : 0u);                          // pipeline stages stay 1024-byte aligned
  const int nplanes = (p.nsplit == 3) ? 2 : 1;
  const uint32_t stage_bytes = nplanes * (kABytes + kBBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && elect_one()) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < kAccStages; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ===================================================================== TMA producer (one thread)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {
          const int tw = m_tile % p.tiles_w;
          const int th = (m_tile / p.tiles_w) % p.tiles_h;
          img = m_tile / (p.tiles_w * p.tiles_h);
          h0 = th * p.tile_h;
          w0 = tw * p.tile_w;
        }
        // K block -> (tap row r, tap column s, channel block cb) by counting: this one thread feeds a stage every 384 tensor
        // cycles on the 64-wide tiles, integer divisions per stage would make it the bottleneck (measured in gemm_gn_sm100.cu)
        int cb = 0, r = 0, s = 0;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + nplanes * kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], nplanes * (p.a_tx_bytes + kBBytes));
          for (int pl = 0; pl < nplanes; ++pl) {
            if (p.conv) {
              tma_load_5d(sA + pl * kABytes, &tmA, &full_bar[stage], cb * kBlockK, w0 + s - p.pad_w,
                          h0 + r - p.pad_h, img, pl);
            } else {
              tma_load_3d(sA + pl * kABytes, &tmA, &full_bar[stage], kb * kBlockK, m_tile * kBlockM, pl);
            }
            tma_load_3d(sB + pl * kBBytes, &tmB, &full_bar[stage], kb * kBlockK, n_tile * BLOCK_N, pl);
          }
          if (++cb == p.cin_blocks) { cb = 0; if (++s == p.KW) { s = 0; ++r; } }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer (one thread)
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBlockM, BLOCK_N, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t aH = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t bH = aH + nplanes * kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t da = umma_desc_k_sw128(aH + k * 32);
            const uint64_t db = umma_desc_k_sw128(bH + k * 32);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0);
            if (p.nsplit == 3) {
              const uint64_t dal = umma_desc_k_sw128(aH + kABytes + k * 32);
              const uint64_t dbl = umma_desc_k_sw128(bH + kBBytes + k * 32);
              umma_f16(d_tmem, dal, db, idesc, 1);
              umma_f16(d_tmem, da, dbl, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);                  // frees this smem stage when the MMAs retire
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================================================================ epilogue warps
    const int ew = warp & 3;                               // TMEM lane quarter this warp may access
    const int row_in_tile = ew * 32 + lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t epi_chunk = 0;                                 // TMA_EPI: staging buffer = chunk counter & 1
    (void)epi_chunk;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_tile = tile / p.n_tiles, n_tile = tile % p.n_tiles;
      long long out_row;
      bool row_ok;
      if (p.conv) {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int img = m_tile / (p.tiles_w * p.tiles_h);
        const int lh = row_in_tile / p.tile_w, lw = row_in_tile % p.tile_w;
        const int h = th * p.tile_h + lh, w = tw * p.tile_w + lw;
        row_ok = (lh < p.tile_h) && (h < p.H) && (w < p.W);
        out_row = ((long long)img * p.H + h) * p.W + w;
      } else {
        out_row = (long long)m_tile * kBlockM + row_in_tile;
        row_ok = out_row < p.M;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c0, r);
        tmem_ld_wait();
        const int col0 = n_tile * BLOCK_N + c0;
        if constexpr (TMA_EPI) {
          // every epilogue thread takes part in the barriers; rows / columns outside the matrix are clipped by the TMA store
          float v[32];
          if (row_ok && col0 < p.N) {
            epilogue_math(v, r, p, col0, out_row);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0.f;
          }
          const int epi_tid = threadIdx.x - 128;
          uint8_t* stg = smem0 + (epi_chunk & 1) * 16384;
          ++epi_chunk;
          if (epi_tid == 0) tma_store_wait_read<1>();        // the store issued two chunks ago has left this buffer
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (p.out_mode == OUT_F32) {                       // 128 rows x 128 B, 128-byte swizzle
            uint8_t* rowp = stg + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(rowp + ((j ^ (row_in_tile & 7)) << 4)) =
                  make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {                                           // hi (and lo at +8192): 128 rows x 64 B, 64-byte swizzle
            uint8_t* rowp = stg + row_in_tile * 64;
            const uint32_t sw = (row_in_tile >> 1) & 3;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint32_t hi[4], lo[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const __half2 h2 = __floats2half2_rn(v[q * 8 + 2 * k], v[q * 8 + 2 * k + 1]);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(v[q * 8 + 2 * k] - hf.x, v[q * 8 + 2 * k + 1] - hf.y);
                hi[k] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[k] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              *reinterpret_cast<uint4*>(rowp + ((q ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(rowp + 8192 + ((q ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          fence_proxy_async();
          asm volatile("bar.sync 2, 128;" ::: "memory");
          if (epi_tid == 0 && col0 < p.N) {
            const int row0 = m_tile * kBlockM;
            if (p.out_mode == OUT_F32) {
              tma_store_3d(&tmO, stg, 2 * col0, row0, 0);     // fp32 matrix described as 2x as many 16-bit columns
            } else {
              tma_store_3d(&tmO, stg, col0, row0, 0);
              if (p.out_mode == OUT_F16_SPLIT) tma_store_3d(&tmO, stg + 8192, col0, row0, 1);
            }
            tma_store_commit();
          }
        } else {
          if (row_ok && col0 < p.N) {                        // N is a multiple of 32 for every MAED layer
            float v[32];
            epilogue_math(v, r, p, col0, out_row);
            if (p.out_mode == OUT_F32) {
              float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + out_row * p.ldc + col0);
#pragma unroll
              for (int j = 0; j < 32; j += 4) op[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              __half* oh = static_cast<__half*>(p.out) + out_row * p.ldc + col0;
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const __half2 h2 = __floats2half2_rn(v[j], v[j + 1]);
                const float2 hf = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
                hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
              }
              uint4* o4 = reinterpret_cast<uint4*>(oh);
#pragma unroll
              for (int j = 0; j < 4; ++j) o4[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
              if (p.out_mode == OUT_F16_SPLIT) {
                uint4* l4 = reinterpret_cast<uint4*>(oh + p.out_plane_stride);
#pragma unroll
                for (int j = 0; j < 4; ++j) l4[j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
    if constexpr (TMA_EPI) {
      if (threadIdx.x == 128) tma_store_wait_all();       // staging must outlive the last bulk store
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace maed
