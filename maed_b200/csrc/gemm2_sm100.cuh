// CTA-pair variant of the tcgen05 GEMM of gemm_sm100.cuh (opt-in: MAED_B200_GEMM_2CTA=1; plain and implicit-conv A operands;
// parity green on a B200 in round 2: tests/test_gemm_pair.py; fc1 414 vs 396 TFLOP/s for the single-CTA kernel with the same
// direct-store epilogue, 430 for the single-CTA kernel with the TMA-store epilogue, which therefore stayed the default).
//
// Why: the 128 x 256 split-precision tile of gemm_tc_kernel is shared-memory-bandwidth bound (profiles/README.md: MMA operand
// reads 96 B/clk + TMA fills 62 B/clk against 128 B/clk; tensor pipe 65-73 %).  With tcgen05.mma.cta_group::2 two CTAs of a
// cluster (one TPC) issue ONE M = 256 instruction: each CTA keeps its own 128 rows of A and only HALF of the B tile, so the B
// fill and the B operand reads per SM halve and a stage shrinks from 96 KB to 64 KB (3 stages instead of 2 at BLOCK_N = 256).
//
// Protocol (every barrier exists at the same shared-memory offset in both CTAs; rank 0 = leader):
//   full[s]        leader's only.  Leader's producer: arrive.expect_tx(2 x bytes per CTA); BOTH producers issue their TMA
//                  loads with .cta_group::2 and the leader's barrier as completion target.
//   empty[s]       local in each CTA (count 1), signalled in both CTAs by the leader's multicast tcgen05.commit.
//   tmem_full[a]   local in each CTA (count 1), multicast commit after the last K block of a tile.
//   tmem_empty[a]  leader's only, count 8: the four epilogue warps of each CTA arrive remotely (mapa + mbarrier.arrive).
// The accumulator of rows [0,128) of a 256-row tile lives in the leader's TMEM, rows [128,256) in the peer's, at the same
// TMEM address; each CTA's epilogue drains its own half exactly like the single-CTA kernel.
#pragma once
#include "gemm_sm100.cuh"

namespace maed {

// registers -> global for one row x 32 columns (same arithmetic and layout as the direct-store epilogue of gemm_tc_kernel)
__device__ __forceinline__ void epilogue_store_row32(const float (&v)[32], const GemmParams& p, int col0, long long out_row) {
  if (p.out_mode == OUT_F32) {
    float4* op = reinterpret_cast<float4*>(static_cast<float*>(p.out) + out_row * p.ldc + col0);
#pragma unroll
    for (int j = 0; j < 32; j += 4) op[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    return;
  }
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

// p.m_tiles counts 128-row tiles as in gemm_tc_kernel (conv mode: one (image, tile_h x tile_w) patch each); a pair works on
// tiles 2i (leader) and 2i + 1 (peer) — an odd last tile leaves the peer with zero-filled loads and no stores.  tmB's box is
// {64, BLOCK_N / 2, 1}.
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using namespace sm100;
  static_assert(BLOCK_N == 128 || BLOCK_N == 256, "pair tile width");
  constexpr int kAccStages = 2;
  constexpr int kTmemCols = kAccStages * BLOCK_N;                 // 256 or 512 columns
  constexpr uint32_t kABytes = kBlockM * kBlockK * 2;             // 16 KB: this CTA's 128 rows of A
  constexpr uint32_t kBBytes = (BLOCK_N / 2) * kBlockK * 2;       // this CTA's half of the B tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nplanes = (p.nsplit == 3) ? 2 : 1;
  const uint32_t stage_bytes = nplanes * (kABytes + kBBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kMaxStages;
  uint64_t* tmem_full = bars + 2 * kMaxStages;
  uint64_t* tmem_empty = bars + 2 * kMaxStages + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5;
  const uint32_t rank = cluster_ctarank();                        // 0 = leader
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int m_pairs = (p.m_tiles + 1) >> 1;
  const int num_tiles = m_pairs * p.n_tiles;

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
      mbar_init(&tmem_empty[a], 8);                               // 4 epilogue warps x 2 CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {                                                // one warp of EACH CTA of the pair
    tmem_alloc_pair(tmem_base_ptr, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                                             // barriers of both CTAs initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ============================================================ TMA producer (one thread in each CTA)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m_tile = 2 * (tile / p.n_tiles) + (int)rank, n_tile = tile % p.n_tiles;   // this CTA's 128-row tile
        const int row0 = m_tile * kBlockM;
        const int col0 = n_tile * BLOCK_N + (int)rank * (BLOCK_N / 2);
        int img = 0, h0 = 0, w0 = 0;
        if (p.conv) {                                               // m_tile == p.m_tiles (odd tail): img == n_img, all zero fill
          const int tw = m_tile % p.tiles_w;
          const int th = (m_tile / p.tiles_w) % p.tiles_h;
          img = m_tile / (p.tiles_w * p.tiles_h);
          h0 = th * p.tile_h;
          w0 = tw * p.tile_w;
        }
        int cb = 0, r = 0, s = 0;                                   // K block -> (tap row, tap column, channel block) by counting
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + nplanes * kABytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * nplanes * (p.a_tx_bytes + kBBytes));
          const uint32_t leader_full = map_to_cta(smem_u32(&full_bar[stage]), 0);
          for (int pl = 0; pl < nplanes; ++pl) {
            if (p.conv) {
              tma_load_5d_pair(sA + pl * kABytes, &tmA, leader_full, cb * kBlockK, w0 + s - p.pad_w, h0 + r - p.pad_h, img, pl);
            } else {
              tma_load_3d_pair(sA + pl * kABytes, &tmA, leader_full, kb * kBlockK, row0, pl);
            }
            tma_load_3d_pair(sB + pl * kBBytes, &tmB, leader_full, kb * kBlockK, col0, pl);
          }
          if (++cb == p.cin_blocks) { cb = 0; if (++s == p.KW) { s = 0; ++r; } }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (one thread, leader only)
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(2 * kBlockM, BLOCK_N, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
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
            umma_f16_pair(d_tmem, da, db, idesc, (kb | k) != 0);
            if (p.nsplit == 3) {
              const uint64_t dal = umma_desc_k_sw128(aH + kABytes + k * 32);
              const uint64_t dbl = umma_desc_k_sw128(bH + kBBytes + k * 32);
              umma_f16_pair(d_tmem, dal, db, idesc, 1);
              umma_f16_pair(d_tmem, da, dbl, idesc, 1);
            }
          }
          umma_commit_pair(&empty_bar[stage], 3);                  // frees this stage in BOTH CTAs
          if (kb == p.num_k_blocks - 1) umma_commit_pair(&tmem_full[acc], 3);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ==================================================================== epilogue warps (both CTAs)
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int m_tile = 2 * (tile / p.n_tiles) + (int)rank, n_tile = tile % p.n_tiles;
      long long out_row;
      bool row_ok;
      if (p.conv) {
        const int tw = m_tile % p.tiles_w;
        const int th = (m_tile / p.tiles_w) % p.tiles_h;
        const int img = m_tile / (p.tiles_w * p.tiles_h);
        const int lh = row_in_tile / p.tile_w, lw = row_in_tile % p.tile_w;
        const int h = th * p.tile_h + lh, w = tw * p.tile_w + lw;
        row_ok = (m_tile < p.m_tiles) && (lh < p.tile_h) && (h < p.H) && (w < p.W);
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
        if (row_ok && col0 < p.N) {
          float v[32];
          epilogue_math(v, r, p, col0, out_row);
          epilogue_store_row32(v, p, col0, out_row);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[acc]), 0));
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();                                             // nobody leaves while the peer may still touch its memory
  if (warp == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
}

}  // namespace maed
