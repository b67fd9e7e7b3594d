// Fused StdConv -> GroupNorm(32) -> (+shortcut) -> ReLU for the ResNetV2 backbone (reference resnetv2.py:189-204):
// the conv runs as a tcgen05 GEMM whose accumulators for ONE IMAGE x ONE CHANNEL BLOCK stay resident in tensor
// memory until the GroupNorm statistics of that image are known, so the fp32 conv output never goes to HBM.
//
//   item    = (image, channel block of BN output channels); BN is a multiple of the group size C/32
//   cluster = CS CTAs (1, 2, 4 or 8) share an item; CTA r owns the image's M tiles [r*TPC, (r+1)*TPC), TPC*BN <= 256 columns
//   Persistent, warp specialised (384 threads), TMEM double buffered (2 x 256 columns = 2 items in flight):
//     warp 0      TMA producer (A/B operand stages; also TMA-prefetches the shortcut boxes of the item into L2)
//     warp 1      MMA issuer (split fp16: 3 tcgen05.mma per K step into TMEM columns half*256 + t*BN)
//     warps 2, 3  epilogue I/O streams, one per epilogue group: TMA loads of shortcut half-boxes, TMA stores of results
//     warps 4-7 / 8-11  two epilogue groups; group g owns TMEM half g and the items of parity g
//   epilogue of an item:
//     pass 1   per-group sum / sum of squares of the valid rows (as soon as a tile's MMAs retire)
//     reduce   warp shuffles -> smem (CTA) -> every CTA pushes its partials into all peers' shared memory (DSMEM stores
//              + remote mbarrier arrive); no cluster-wide barrier, so producers never stall on the epilogue
//     pass 2   TMEM -> v*a_c + b_c (folded mean/rstd/gamma/beta) (+ shortcut, read from its smem slot) -> ReLU -> fp16
//              hi/lo written IN PLACE into the slot (64-byte-swizzled half-box) -> mbarrier arrive; the I/O stream
//              stores the slot with TMA and refills it.  No named barrier, no per-thread global access.
//
// Versus the unfused pipeline (GEMM writes fp32, gn_stats reads it, gn_apply reads it again and writes planes)
// this removes 12 of the 16 bytes of HBM traffic per conv-output element.  In-kernel timeline: scripts/dbg_gn_timeline.py.
#include <cstdlib>

#include "gemm_host.h"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

static constexpr int kGnThreads = 384;       // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue
static constexpr int kGnMaxTpc = 4;
static constexpr int kGnMaxStages = 4;
static constexpr int kGnMaxCluster = 8;

struct GemmGnParams {
  int HW, C, K, num_k_blocks, nsplit, stages;   // stages: depth of the A ring
  int b_stages, b_shared;      // depth of the B ring; b_shared: one B load per K block serves all tiles of the CTA
  int tiles_per_image, tpc, n_blocks, cluster, items;
  uint32_t a_tx_bytes;
  uint32_t o_tx_bytes;         // bytes of one 32-channel half-box plane (rows x 64 B)
  uint32_t box_bytes;          // staging bytes after the pipeline stages: 2 groups x res_slots x 16 KB
  int res_slots;               // staging half-box slots per epilogue group (2 or 3)
  int conv, H, W, cin_blocks, KW, pad_h, pad_w, tile_h, tile_w, tiles_h, tiles_w;
  const float* gamma; const float* beta; float eps; int relu;
  const __half* res; long long res_plane;
  __half* out; long long out_plane;
  long long* dbg;              // optional [items][8] clock64 timestamps of (cluster 0, rank 0, group 0); nullptr = off
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_dsmem_f64(uint32_t remote_addr, double v) {
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(remote_addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t remote_bar_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  // Poll with the default (CTA-scope) try_wait: a cluster-scope acquire poll makes ptxas emit an L1 invalidate
  // (CCTL.IVALL) per iteration.  One cluster-scope fence after success orders the peers' DSMEM stores before our reads.
  sm100::mbar_wait(bar, parity);
  asm volatile("fence.acq_rel.cluster;" ::: "memory");
}
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// PAIR: the two CTAs of a cluster (tiles 0 and 1 of a 14 x 14 image) issue ONE tcgen05.mma.cta_group::2 per step (M = 256):
// each CTA stages its own 128 rows of A and only HALF of the B block, so a BN = 256 block fits (64 KB per stage) and the
// shared-memory traffic per MMA cycle drops from 128 + 85 B/clk (128 x 128 tiles) to 64 + 42 B/clk.  Protocol as in
// gemm2_sm100.cuh: operand-full barriers live in the leader (rank 0) and collect the bytes of both producers, empty / tile-full
// barriers are signalled in both CTAs by multicast commits, the TMEM-half release collects all eight epilogue warps of the pair
// in the leader.  Everything behind the accumulator (statistics exchange over DSMEM, epilogue, I/O streams) is per CTA as before.
template <int BN, int GSZ, bool PAIR>
__global__ void __launch_bounds__(kGnThreads, 1)
gemm_gn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
               const __grid_constant__ CUtensorMap tmOw, const GemmGnParams p) {
  using namespace sm100;
  constexpr int G = BN / GSZ;                 // groups in the channel block
  static_assert(G >= 1 && G <= 32 && BN % 32 == 0 && BN <= 256 && (BN <= 128 || PAIR), "unsupported block / group shape");
  constexpr uint32_t kABytes = 128 * 64 * 2, kBBytes = (PAIR ? BN / 2 : BN) * 64 * 2;   // PAIR: this CTA's half of the B block
  constexpr int kCoef = BN > 128 ? BN : 128;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic ST/LD)
  const int np = p.nsplit == 3 ? 2 : 1;
  // operand rings: A slots [stages][np][128 x 64] then B slots [b_stages][np][BN x 64].  They are separate because a B
  // block (a K slice of the weights) is the same for every M tile of the item: with b_shared the loops run K-block-outer /
  // tile-inner and each B block crosses the L2 -> SM port once per item instead of once per tile.  These kernels are bound
  // by that port (A + B bytes per MMA cycle), not by the tensor pipe.
  const uint32_t a_slot = np * kABytes, b_slot = np * kBBytes;
  uint8_t* sBring = smem + (size_t)p.stages * a_slot;
  // staging after the pipeline stages.  Shortcut layers: [2 groups][kResSlots][hi | lo][128 rows x 64 B] half-boxes (the
  // shortcut lands here by TMA, is updated IN PLACE and stored by TMA); other layers: [2 groups][hi | lo][128 x 128 B] boxes.
  uint8_t* sOut = sBring + (size_t)p.b_stages * b_slot;
  uint8_t* sRes = sOut;
  double* s_warp_part = reinterpret_cast<double*>(sOut + p.box_bytes);     // [2 groups][4 warps][32 groups][2]
  double* s_parts = s_warp_part + 2 * 4 * 32 * 2;                          // [2 groups][8 ranks][32 groups][2]
  float* s_mr = reinterpret_cast<float*>(s_parts + 2 * kGnMaxCluster * 32 * 2);   // [2 groups][32][2] mean, rstd
  float* s_coef = s_mr + 2 * 64;                                           // [2 groups][a_c[kCoef] | b_c[kCoef]]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_coef + 2 * 2 * kCoef);
  uint64_t* full_bar = bars;                                               // [stages]
  uint64_t* empty_bar = bars + kGnMaxStages;
  uint64_t* tile_full = bars + 2 * kGnMaxStages;                           // [2 halves][kGnMaxTpc]
  uint64_t* half_empty = tile_full + 2 * kGnMaxTpc;                        // [2]
  uint64_t* parts_full = half_empty + 2;                                   // [2 buf]
  uint64_t* res_full = parts_full + 2;                                     // [2 groups][4 slots]: slot may be used by the epilogue
  uint64_t* box_ready = res_full + 8;                                      // [2 groups][4 slots]: slot holds a finished result
  uint64_t* b_full = box_ready + 8;                                        // [kGnMaxStages] B ring
  uint64_t* b_empty = b_full + kGnMaxStages;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(b_empty + kGnMaxStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CS = p.cluster;
  const uint32_t crank = CS > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS, n_clusters = gridDim.x / CS;
  const int t_lo = crank * p.tpc;
  const int my_tiles = max(0, min(p.tpc, p.tiles_per_image - t_lo));

  if (warp == 0 && elect_one()) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); prefetch_tmap(&tmO); prefetch_tmap(&tmOw); if (p.res) prefetch_tmap(&tmR); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < p.b_stages; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
    for (int t = 0; t < 2 * kGnMaxTpc; ++t) mbar_init(&tile_full[t], 1);
    for (int h = 0; h < 2; ++h) { mbar_init(&half_empty[h], PAIR ? 8 : 4); mbar_init(&parts_full[h], CS * G); }
    for (int h = 0; h < 8; ++h) { mbar_init(&res_full[h], 1); mbar_init(&box_ready[h], 128); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) { tmem_alloc_pair(tmem_base_ptr, 512); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_base_ptr, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (CS > 1) cluster_sync_all();             // peers' mbarriers are initialised before anyone arrives on them
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    // ================================================================================ TMA producer
    if (elect_one()) {
      int stage = 0; uint32_t phase = 0;
      int sb = 0; uint32_t bphase = 0;
      // this one thread must hand out a stage every few hundred cycles (384 tensor cycles per stage on the 64-wide blocks): no
      // integer divisions inside the loop — tile origins once per kernel, the K block -> (tap row, tap column, channel block)
      // decomposition by counting
      const int h00 = p.conv ? (t_lo / p.tiles_w) * p.tile_h : 0, w00 = p.conv ? (t_lo % p.tiles_w) * p.tile_w : 0;
      const int wlim = p.tiles_w * p.tile_w;
      int h0 = h00, w0 = w00;
      int cb = 0, tr = 0, ts = 0;
      for (int item = cluster_id; item < p.items; item += n_clusters) {
        const int img = item / p.n_blocks, nb = item % p.n_blocks;
        if (p.res) {
          // pull the shortcut boxes of this item into L2 now: the epilogue reads them ~2 items later
          for (int tl = 0; tl < my_tiles; ++tl) {
            const int t = t_lo + tl;
            for (int cb = 0; cb < BN; cb += 32)
              for (int pl = 0; pl < 2; ++pl) {
                if (p.conv) tma_prefetch_l2_5d(&tmR, nb * BN + cb, (t % p.tiles_w) * p.tile_w, (t / p.tiles_w) * p.tile_h, img, pl);
                else tma_prefetch_l2_4d(&tmR, nb * BN + cb, t * 128, img, pl);
              }
          }
        }
        const int n_outer = p.b_shared ? p.num_k_blocks : my_tiles, n_inner = p.b_shared ? my_tiles : p.num_k_blocks;
        for (int o = 0; o < n_outer; ++o) {
          for (int i = 0; i < n_inner; ++i) {
            const int kb = p.b_shared ? o : i, tl = p.b_shared ? i : o;
            if (!p.b_shared || i == 0) {                   // a new K block
              if (kb == 0) { cb = 0; tr = 0; ts = 0; }
              else if (++cb == p.cin_blocks) { cb = 0; if (++ts == p.KW) { ts = 0; ++tr; } }
            }
            if (p.b_shared || i == 0) {                    // a new tile (row-major over the image)
              if (tl == 0) { h0 = h00; w0 = w00; }
              else { w0 += p.tile_w; if (w0 >= wlim) { w0 = 0; h0 += p.tile_h; } }
            }
            if (p.b_shared && i == 0) {                    // shared B block: its own ring and barriers (never with PAIR)
              mbar_wait(&b_empty[sb], bphase ^ 1);
              uint8_t* sB = sBring + (size_t)sb * b_slot;
              mbar_arrive_expect_tx(&b_full[sb], np * kBBytes);
              for (int pl = 0; pl < np; ++pl) tma_load_3d(sB + pl * kBBytes, &tmB, &b_full[sb], kb * 64, nb * BN, pl);
              if (++sb == p.b_stages) { sb = 0; bphase ^= 1; }
            }
            const int t = t_lo + tl;
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sA = smem + (size_t)stage * a_slot;
            // otherwise B slot `stage` travels with A slot `stage` under the same pair of barriers
            const uint32_t tx = np * (p.a_tx_bytes + (p.b_shared ? 0u : kBBytes));
            if (!PAIR || crank == 0) mbar_arrive_expect_tx(&full_bar[stage], (PAIR ? 2 : 1) * tx);
            const uint32_t lead = PAIR ? map_to_cta(smem_u32(&full_bar[stage]), 0) : 0;
            if (!p.b_shared) {
              uint8_t* sB = sBring + (size_t)stage * b_slot;
              for (int pl = 0; pl < np; ++pl) {
                if (PAIR) tma_load_3d_pair(sB + pl * kBBytes, &tmB, lead, kb * 64, nb * BN + (int)crank * (BN / 2), pl);
                else tma_load_3d(sB + pl * kBBytes, &tmB, &full_bar[stage], kb * 64, nb * BN, pl);
              }
            }
            for (int pl = 0; pl < np; ++pl) {
              if (p.conv) {
                if (PAIR)
                  tma_load_5d_pair(sA + pl * kABytes, &tmA, lead, cb * 64, w0 + ts - p.pad_w, h0 + tr - p.pad_h, img, pl);
                else
                  tma_load_5d(sA + pl * kABytes, &tmA, &full_bar[stage], cb * 64, w0 + ts - p.pad_w, h0 + tr - p.pad_h,
                              img, pl);
              } else if (PAIR) {
                tma_load_3d_pair(sA + pl * kABytes, &tmA, lead, kb * 64, img * p.HW + t * 128, pl);
              } else {
                tma_load_3d(sA + pl * kABytes, &tmA, &full_bar[stage], kb * 64, img * p.HW + t * 128, pl);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================================== MMA issuer
    if ((!PAIR || crank == 0) && elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(PAIR ? 256 : 128, BN, 0);
      int stage = 0; uint32_t phase = 0;
      int sb = 0; uint32_t bphase = 0;
      uint32_t j = 0;
      for (int item = cluster_id; item < p.items; item += n_clusters, ++j) {
        const uint32_t half = j & 1, hphase = (j >> 1) & 1;      // epilogue group `half` handles this item
        mbar_wait(&half_empty[half], hphase ^ 1);
        tc_fence_after();
        const int n_outer = p.b_shared ? p.num_k_blocks : my_tiles, n_inner = p.b_shared ? my_tiles : p.num_k_blocks;
        for (int o = 0; o < n_outer; ++o) {
          for (int i = 0; i < n_inner; ++i) {
            const int kb = p.b_shared ? o : i, tl = p.b_shared ? i : o;
            const uint32_t d = tmem_base + half * 256 + tl * BN;
            if (p.b_shared && i == 0) mbar_wait(&b_full[sb], bphase);
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t aH = smem_u32(smem + (size_t)stage * a_slot);
            const uint32_t bH = smem_u32(sBring + (size_t)(p.b_shared ? sb : stage) * b_slot);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t da = umma_desc_k_sw128(aH + k * 32), db = umma_desc_k_sw128(bH + k * 32);
              if (PAIR) {
                umma_f16_pair(d, da, db, idesc, (kb | k) != 0);
                if (np == 2) {
                  umma_f16_pair(d, umma_desc_k_sw128(aH + kABytes + k * 32), db, idesc, 1);
                  umma_f16_pair(d, da, umma_desc_k_sw128(bH + kBBytes + k * 32), idesc, 1);
                }
              } else {
                umma_f16(d, da, db, idesc, (kb | k) != 0);
                if (np == 2) {
                  umma_f16(d, umma_desc_k_sw128(aH + kABytes + k * 32), db, idesc, 1);
                  umma_f16(d, da, umma_desc_k_sw128(bH + kBBytes + k * 32), idesc, 1);
                }
              }
            }
            if (PAIR) umma_commit_pair(&empty_bar[stage], 3); else umma_commit(&empty_bar[stage]);
            if (p.b_shared && i == n_inner - 1) {
              umma_commit(&b_empty[sb]);
              if (++sb == p.b_stages) { sb = 0; bphase ^= 1; }
            }
            if (kb == p.num_k_blocks - 1) {
              if (PAIR) umma_commit_pair(&tile_full[half * kGnMaxTpc + tl], 3); else umma_commit(&tile_full[half * kGnMaxTpc + tl]);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ============================================================ epilogue I/O streams (one thread per epilogue group)
    // All global traffic of pass 2 is bulk + asynchronous and issued from here, so the epilogue warps never block on a
    // named barrier: they wait for a slot (res_full), update it in place and signal box_ready.
    //   slot life cycle:  [TMA load of the shortcut half-box | plain arrive]  -> res_full -> epilogue -> box_ready
    //                     -> TMA store -> (store has read the slot) -> refill
    if (elect_one()) {
      const int grp = warp - 2;
      const int RS = p.res_slots;
      constexpr int NCH = BN / 32;
      uint8_t* rbase = sRes + (size_t)grp * (RS * 16384);
      uint64_t* rfull = res_full + grp * 4;
      uint64_t* bready = box_ready + grp * 4;
      uint32_t n_fill = 0, n_store = 0;                   // running slot counters of this stream
      for (int item = cluster_id + grp * n_clusters; item < p.items; item += 2 * n_clusters) {
        const int img = item / p.n_blocks, nb = item % p.n_blocks;
        const int n_it = my_tiles * NCH;
        auto fill = [&](int it) {                         // make slot (n_fill % RS) usable for iteration `it`
          const uint32_t slot = n_fill % RS;
          if (p.res) {
            const int tl = it / NCH, c0 = (it % NCH) * 32, t = t_lo + tl;
            uint8_t* dst = rbase + slot * 16384;
            mbar_arrive_expect_tx(&rfull[slot], 2 * p.o_tx_bytes);
            for (int pl = 0; pl < 2; ++pl) {
              if (p.conv) tma_load_5d(dst + pl * 8192, &tmR, &rfull[slot], nb * BN + c0, (t % p.tiles_w) * p.tile_w, (t / p.tiles_w) * p.tile_h, img, pl);
              else tma_load_4d(dst + pl * 8192, &tmR, &rfull[slot], nb * BN + c0, t * 128, img, pl);
            }
          } else {
            mbar_arrive(&rfull[slot]);
          }
          ++n_fill;
        };
        tma_store_wait_read<0>();                         // every slot of the previous item has been stored
        for (int i = 0; i < RS - 1 && i < n_it; ++i) fill(i);
        for (int it = 0; it < n_it; ++it) {
          const uint32_t slot = n_store % RS;
          mbar_wait(&bready[slot], (n_store / RS) & 1);
          ++n_store;
          const int tl = it / NCH, c0 = (it % NCH) * 32, t = t_lo + tl, ch = nb * BN + c0;
          const uint8_t* sb = rbase + slot * 16384;
          if (p.conv) {
            const int h0 = (t / p.tiles_w) * p.tile_h, w0 = (t % p.tiles_w) * p.tile_w;
            tma_store_5d(&tmO, sb, ch, w0, h0, img, 0);
            tma_store_5d(&tmO, sb + 8192, ch, w0, h0, img, 1);
          } else {
            tma_store_4d(&tmO, sb, ch, t * 128, img, 0);
            tma_store_4d(&tmO, sb + 8192, ch, t * 128, img, 1);
          }
          tma_store_commit();
          if (it + RS - 1 < n_it) {
            tma_store_wait_read<1>();                     // every store but the one just issued has released its slot
            fill(it + RS - 1);
          }
        }
      }
      tma_store_wait_all();
    }
  } else if (warp >= 4) {
    // =============================================================================== epilogue warps
    // Two independent groups of 4 warps: group g owns TMEM half g and the items of parity g, so the statistics
    // exchange / global-memory latency of one item overlaps the arithmetic of the other.
    const int e = warp - 4;                    // 0..7
    const int qw = e & 3;                      // TMEM lane quarter (== warp % 4)
    const int grp = e >> 2;                    // item parity / TMEM half owned by this warp
    const int gt = threadIdx.x - 128 - grp * 128;   // 0..127 inside the group
    const int row_in_tile = qw * 32 + lane;
    const uint32_t lane_off = (uint32_t)(qw * 32) << 16;
    double* wpart = s_warp_part + grp * (4 * 32 * 2);     // [4 warps][32 groups][2]
    float* mr = s_mr + grp * 64;
    float* coef = s_coef + grp * 2 * kCoef;
    const int bar_a = 1 + grp * 2, bar_b = 2 + grp * 2;
    const uint32_t t_half = tmem_base + grp * 256 + lane_off;
    uint32_t jj = 0;                           // per-group item counter
    uint32_t res_use = 0;                      // running slot counter of this group (slot = n % RS, parity = (n / RS) & 1)
    for (int item = cluster_id + grp * n_clusters; item < p.items; item += 2 * n_clusters, ++jj) {
      const int img = item / p.n_blocks, nb = item % p.n_blocks;
      const uint32_t hphase = jj & 1;
      // output row (and validity) of this thread in tile tl of the item
      auto row_info = [&](int tl, long long* orow) -> bool {
        const int t = t_lo + tl;
        if (p.conv) {
          const int lh = row_in_tile / p.tile_w, lw = row_in_tile % p.tile_w;
          const int h = (t / p.tiles_w) * p.tile_h + lh, w = (t % p.tiles_w) * p.tile_w + lw;
          *orow = ((long long)img * p.H + h) * p.W + w;
          return lh < p.tile_h && h < p.H && w < p.W;
        }
        const int r = t * 128 + row_in_tile;
        *orow = (long long)img * p.HW + r;
        return r < p.HW;
      };
      const bool dbg_on = p.dbg && blockIdx.x == 0 && gt == 0 && grp == 0;
      if (dbg_on) p.dbg[jj * 8 + 0] = clock64();
      // ---------------------------------------------------------------- pass 1: group statistics
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        constexpr int GC = (32 / GSZ) > 0 ? (32 / GSZ) : 1;        // groups touched by a 32-column chunk
        float gs[GC], gq[GC];
#pragma unroll
        for (int g = 0; g < GC; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
        for (int tl = 0; tl < my_tiles; ++tl) {
          if (c0 == 0) { mbar_wait(&tile_full[grp * kGnMaxTpc + tl], hphase); tc_fence_after(); if (dbg_on && tl == 0) p.dbg[jj * 8 + 6] = clock64(); }
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_half + tl * BN + c0, r);
          tmem_ld_wait();
          long long orow_unused;
          if (row_info(tl, &orow_unused)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float v = __uint_as_float(r[i]);
              gs[GSZ >= 32 ? 0 : i / GSZ] += v;
              gq[GSZ >= 32 ? 0 : i / GSZ] += v * v;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < GC; ++g) {
          float sv = gs[g], qv = gq[g];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, o); qv += __shfl_xor_sync(0xffffffffu, qv, o); }
          if (lane == 0) {
            // for GSZ = 32 a chunk is exactly one group; partial sums of the same group from different chunks never occur
            const int gidx = (GSZ >= 32) ? c0 / 32 : c0 / GSZ + g;
            wpart[(qw * 32 + gidx) * 2] = (double)sv;
            wpart[(qw * 32 + gidx) * 2 + 1] = (double)qv;
          }
        }
      }
      if (dbg_on) p.dbg[jj * 8 + 1] = clock64();
      named_bar(bar_a, 128);
      if (dbg_on) p.dbg[jj * 8 + 2] = clock64();
      // ---- CTA partial of every group, pushed to every CTA of the cluster (including this one).  One thread per (group, peer):
      // a remote release-arrive costs ~1.2 k cycles and a thread's arrives serialise — with one thread per group walking the 8
      // peers the exchange took 10-13 k cycles per item (in-kernel timeline, round 2), i.e. a third of the item period.
      if (CS > 1) {
        const uint32_t lb = smem_u32(&parts_full[grp]);
        for (int idx = gt; idx < G * CS; idx += 128) {
          const int g = idx % G, r = idx / G;
          double sv = 0.0, qv = 0.0;
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) { sv += wpart[(w4 * 32 + g) * 2]; qv += wpart[(w4 * 32 + g) * 2 + 1]; }
          const uint32_t ra = mapa_u32(smem_u32(s_parts + ((grp * kGnMaxCluster + crank) * 32 + g) * 2), r);
          st_dsmem_f64(ra, sv);
          st_dsmem_f64(ra + 8, qv);
          mbar_arrive_remote(mapa_u32(lb, r));
        }
      }
      if (gt < G) {
        const int g = gt;
        double sv = 0.0, qv = 0.0;
        if (CS > 1) {
          mbar_wait_cluster(&parts_full[grp], hphase);
        } else {
#pragma unroll
          for (int w4 = 0; w4 < 4; ++w4) { sv += wpart[(w4 * 32 + g) * 2]; qv += wpart[(w4 * 32 + g) * 2 + 1]; }
          double* slot = s_parts + ((grp * kGnMaxCluster + crank) * 32 + g) * 2;
          slot[0] = sv; slot[1] = qv;
        }
        sv = 0.0; qv = 0.0;
        for (int r = 0; r < CS; ++r) {
          const double* ps = s_parts + ((grp * kGnMaxCluster + r) * 32 + g) * 2;
          sv += ps[0]; qv += ps[1];
        }
        const double cnt = (double)p.HW * GSZ;
        const double m = sv / cnt;
        double var = qv / cnt - m * m;
        if (var < 0.0) var = 0.0;
        mr[g * 2] = (float)m;
        mr[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)p.eps));
      }
      if (dbg_on) p.dbg[jj * 8 + 3] = clock64();
      named_bar(bar_b, 128);
      for (int c = gt; c < BN; c += 128) {     // y = v * a_c + b_c
        const int g = c / GSZ;
        const float ga = __ldg(p.gamma + nb * BN + c), be = __ldg(p.beta + nb * BN + c);
        const float a = mr[g * 2 + 1] * ga;
        coef[c] = a;
        coef[kCoef + c] = be - mr[g * 2] * a;
      }
      named_bar(bar_a, 128);
      // ------------------------------------------- pass 2: normalise, shortcut, ReLU, split -> smem boxes -> TMA stores
      // Unit of work = (tile, 64-channel block): every thread writes its row's 64 channels (128 B per plane) into the
      // group's hi / lo staging boxes in the 128-byte-swizzled layout of the output tensor map; one thread then issues
      // two bulk tensor stores (rows outside the image are clipped by the tensor map).  The shortcut of the next
      // 32-column chunk is loaded while the current one is processed.
      if (dbg_on) p.dbg[jj * 8 + 4] = clock64();
      // Iteration = (tile, 32-channel chunk).  All global traffic of the epilogue is bulk and asynchronous:
      //   * the shortcut half-box (128 rows x 32 channels, hi + lo plane) of iteration it + kResSlots is fetched by TMA into
      //     a shared-memory slot while earlier iterations compute (row-per-thread loads touched 32 cache lines per
      //     instruction and made the LSU the bottleneck);
      //   * results are staged in a 64-byte-swizzled half-box and written by TMA stores (clipped at the image edge).
      constexpr int NCH = BN / 32;
      const int n_it = my_tiles * NCH;
      const int RS = p.res_slots;
      uint8_t* rbase = sRes + (size_t)grp * (RS * 16384);
      uint64_t* rfull = res_full + grp * 4;
      uint64_t* bready = box_ready + grp * 4;
      const uint32_t sw = (row_in_tile >> 1) & 3;                          // 64-byte swizzle: chunk ^= (row >> 1) & 3
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        const int tl = it / NCH, c0 = (it % NCH) * 32;
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_half + tl * BN + c0, r);
        tmem_ld_wait();
        float v[32];
        const float* ca = coef + c0;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 a4 = *reinterpret_cast<const float4*>(ca + i);
          const float4 b4 = *reinterpret_cast<const float4*>(ca + kCoef + i);
          v[i] = __uint_as_float(r[i]) * a4.x + b4.x;
          v[i + 1] = __uint_as_float(r[i + 1]) * a4.y + b4.y;
          v[i + 2] = __uint_as_float(r[i + 2]) * a4.z + b4.z;
          v[i + 3] = __uint_as_float(r[i + 3]) * a4.w + b4.w;
        }
        const uint32_t slot = res_use % RS;
        mbar_wait(&rfull[slot], (res_use / RS) & 1);                       // slot free (and, with a shortcut, loaded)
        ++res_use;
        uint8_t* box = rbase + slot * 16384 + row_in_tile * 64;            // this thread's row: hi at +0, lo at +8192
        if (p.res) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 H = *reinterpret_cast<const uint4*>(box + ((q ^ sw) << 4));
            const uint4 L = *reinterpret_cast<const uint4*>(box + 8192 + ((q ^ sw) << 4));
            const __half2* h2 = reinterpret_cast<const __half2*>(&H);
            const __half2* l2 = reinterpret_cast<const __half2*>(&L);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 a = __half22float2(h2[k]), b = __half22float2(l2[k]);
              v[q * 8 + 2 * k] += a.x + b.x;
              v[q * 8 + 2 * k + 1] += a.y + b.y;
            }
          }
        }
        if (p.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                      // in place: each thread touches only its own row
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __half2 h2 = __floats2half2_rn(v[q * 8 + 2 * k], v[q * 8 + 2 * k + 1]);
            const float2 hf = __half22float2(h2);
            const __half2 l2 = __floats2half2_rn(v[q * 8 + 2 * k] - hf.x, v[q * 8 + 2 * k + 1] - hf.y);
            hi[k] = *reinterpret_cast<const uint32_t*>(&h2);
            lo[k] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          *reinterpret_cast<uint4*>(box + ((q ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(box + 8192 + ((q ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async();                                               // generic-proxy writes -> visible to the TMA store
        mbar_arrive(&bready[slot]);
      }
      if (dbg_on) p.dbg[jj * 8 + 5] = clock64();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(map_to_cta(smem_u32(&half_empty[grp]), 0));     // the leader's MMA thread waits for both CTAs
        else mbar_arrive(&half_empty[grp]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CS > 1) cluster_sync_all();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host
template <int BN, int GSZ, bool PAIR = false>
static int launch_gn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, const CUtensorMap& tmR,
                     const CUtensorMap& tmOw, GemmGnParams& p, cudaStream_t st) {
  const int np = p.nsplit == 3 ? 2 : 1;
  const size_t a_slot = (size_t)np * 128 * 64 * 2, b_slot = (size_t)np * (PAIR ? BN / 2 : BN) * 64 * 2;
  // shortcut layers: 3 staging slots per group (2 shortcut loads in flight: 1.5 k cycles per 32-column iteration instead of 2.8 k)
  // whenever the operand rings still get two stages
  size_t fixed = 0;
  int stages = 0;
  for (p.res_slots = p.res ? 3 : 2; p.res_slots >= 2; --p.res_slots) {
    p.box_bytes = 2 * p.res_slots * 16384;
    fixed = 1024 + p.box_bytes + (2 * 4 * 32 * 2 + 2 * kGnMaxCluster * 32 * 2) * 8 + (128 + 4 * (BN > 128 ? BN : 128)) * 4 +
            (4 * kGnMaxStages + 2 * kGnMaxTpc + 20) * 8 + 64;
    // shared B blocks turn over once per K block: two slots; otherwise the B ring is as deep as the A ring (one B per A tile)
    stages = p.b_shared ? (int)(((long long)232448 - (long long)fixed - 2 * (long long)b_slot) / (long long)a_slot)
                        : (int)(((long long)232448 - (long long)fixed) / (long long)(a_slot + b_slot));
    if (stages >= 2) break;
  }
  if (p.res_slots < 2) p.res_slots = 2;
  if (stages > kGnMaxStages) stages = kGnMaxStages;
  p.b_stages = p.b_shared ? 2 : stages;
  if (stages < 2) { set_error("gemm_gn: tile too large for shared memory"); return MAED_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = fixed + (size_t)stages * a_slot + (size_t)p.b_stages * b_slot;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gemm_gn_kernel<BN, GSZ, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.blockDim = dim3(kGnThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // persistent: exactly as many clusters as can be co-resident (a cluster must fit inside one GPC, so this is
  // fewer than sm_count / cluster for large clusters); a second wave would serialise whole item queues
  static int max_clusters[kGnMaxCluster + 1] = {0};
  if (!max_clusters[p.cluster]) {
    cfg.gridDim = dim3(sm_count() / p.cluster * p.cluster);
    int n = 0;
    MAED_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, gemm_gn_kernel<BN, GSZ, PAIR>, &cfg));
    if (n < 1) { set_error("gemm_gn: cluster of %d CTAs cannot be scheduled", p.cluster); return MAED_ERR_UNSUPPORTED; }
    max_clusters[p.cluster] = n;
  }
  int n_clusters = max_clusters[p.cluster];
  if (n_clusters > p.items) n_clusters = p.items;
  cfg.gridDim = dim3(n_clusters * p.cluster);
  MAED_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_gn_kernel<BN, GSZ, PAIR>, tmA, tmB, tmO, tmR, tmOw, p));
  count_launch();
  return MAED_OK;
}

// Returns MAED_ERR_UNSUPPORTED (without setting up a launch) for shapes the fused kernel does not cover, so the
// caller can use the unfused gemm + gn_stats + gn_apply kernels.
int conv_gn_fused(const ConvGnArgs& a, cudaStream_t st) {
  const int np = a.nsplit == 3 ? 2 : 1;
  MAED_CHECK_ARG(a.C % 32 == 0, "conv_gn_fused: C=%d", a.C);
  const int gsz = a.C / 32;
  GemmGnParams p;
  memset(&p, 0, sizeof(p));
  p.HW = a.H_out * a.W_out; p.C = a.C; p.nsplit = a.nsplit;
  p.gamma = a.gamma; p.beta = a.beta; p.eps = a.eps; p.relu = a.relu; p.res = a.res; p.res_plane = a.res_plane;
  p.out = a.out; p.out_plane = a.out_plane;
  p.dbg = a.dbg;
  CUtensorMap tmA, tmB, tmO, tmR, tmOw;
  const long long M = (long long)a.n_img * p.HW;
  if (a.conv) {
    if (a.Cin % 64 != 0) return MAED_ERR_UNSUPPORTED;
    p.conv = 1; p.H = a.H_out; p.W = a.W_out; p.cin_blocks = a.Cin / 64; p.KW = a.KW; p.pad_h = a.pad_h; p.pad_w = a.pad_w;
    conv_tile_shape(p.H, p.W, &p.tile_h, &p.tile_w);
    p.tiles_h = cdiv(p.H, p.tile_h); p.tiles_w = cdiv(p.W, p.tile_w);
    p.tiles_per_image = p.tiles_h * p.tiles_w;
    p.K = a.KH * a.KW * a.Cin; p.num_k_blocks = a.KH * a.KW * p.cin_blocks;
    p.a_tx_bytes = (uint32_t)(p.tile_h * p.tile_w * 128);
  } else {
    p.K = a.K; p.num_k_blocks = cdiv(a.K, 64);
    p.tiles_per_image = cdiv(p.HW, 128);
    p.a_tx_bytes = 128 * 64 * 2;
  }
  // channel-block width, tiles per CTA and cluster size: an image's tiles x BN columns must fit half of the TMEM
  // (256 columns) of the CTAs of one cluster
  int bn, cluster, tpc;
  bool pair = false;
  // 14 x 14 maps (exactly two M tiles per image): the two CTAs of a cluster work as a cta_group::2 pair on 256-wide blocks
  // (MAED_B200_GN_PAIR=0: the single-CTA plans below)
  static const bool pair_on = !(getenv("MAED_B200_GN_PAIR") && atoi(getenv("MAED_B200_GN_PAIR")) == 0);
  // measured: 3x3 126 -> 103 us (tensor pipe 53 -> 66 %), 1x1 63 -> 58 us; the 1024-channel shortcut layers are epilogue-bound
  // and lose with 256-wide blocks (97 -> 108 us), so they keep the single-CTA plan
  // shortcut layers of those maps (1024 channels) as pairs on 128-wide blocks (a stage is 48 KB, which leaves room for the third
  // staging slot per group): measured 100 vs 98 us for the single-CTA plan — opt-in only (MAED_B200_GN_PAIR_RES=1)
  static const bool pair_res = getenv("MAED_B200_GN_PAIR_RES") && atoi(getenv("MAED_B200_GN_PAIR_RES")) == 1;
  if (pair_on && p.tiles_per_image == 2 && !a.res && a.C % 256 == 0 && gsz == 8) { bn = 256; tpc = 1; cluster = 2; pair = true; }
  else if (pair_on && pair_res && p.tiles_per_image == 2 && a.res && a.C % 128 == 0 && gsz == 32) { bn = 128; tpc = 1; cluster = 2; pair = true; }
  else if (a.res) {
    // shortcut layers are epilogue-bound: 64-wide blocks leave shared memory for 3 shortcut slots (2 TMA loads in flight)
    // ... except with at most 2 tiles per image (stage 2: 14 x 14): there the L2 -> SM port is the bound and a 128-wide
    // block halves the number of times the image's A tiles are fetched (MAED_B200_GN_RES_BN128=0: the 64-wide plan)
    static const bool res128 = !(getenv("MAED_B200_GN_RES_BN128") && atoi(getenv("MAED_B200_GN_RES_BN128")) == 0);
    if (p.tiles_per_image <= 2 && res128 && a.C % 128 == 0 && 128 % gsz == 0) { bn = 128; tpc = 2; cluster = 1; }
    else if (p.tiles_per_image <= 2) { bn = 64; tpc = 2; cluster = 1; }
    else if (p.tiles_per_image <= 8) { bn = 64; tpc = 4; cluster = 2; }
    else if (p.tiles_per_image <= 32) { bn = 64; tpc = 4; cluster = 8; }
    else return MAED_ERR_UNSUPPORTED;
  }
  else if (p.tiles_per_image <= 2 && a.C % 128 == 0) { bn = 128; tpc = 2; cluster = 1; }
  else if (p.tiles_per_image <= 8 && a.C % 128 == 0) { bn = 128; tpc = 2; cluster = 4; }
  else if (p.tiles_per_image <= 8) { bn = 64; tpc = 4; cluster = 2; }
  else if (p.tiles_per_image <= 32) { bn = 64; tpc = 4; cluster = 8; }
  else return MAED_ERR_UNSUPPORTED;
  if (bn % gsz != 0 || (bn / 2) % gsz != 0 || a.C % bn != 0 || tpc * bn > 256 || tpc * cluster < p.tiles_per_image)
    return MAED_ERR_UNSUPPORTED;
  p.cluster = cluster; p.tpc = tpc;
  // one B load per K block for all tiles of the CTA: measured (profiles/r02_fwd_step_per_launch_v3_l2.txt) -25 % bytes through
  // the L2 -> SM port everywhere, but faster only with 2 tiles per CTA (stage 1/2: -3 .. -8 %); with 4 tiles per CTA all tiles
  // of an item finish together and the epilogue loses its head start (stage 0: +4 .. +12 %)
  static const bool b_shared = !(getenv("MAED_B200_GN_BSHARED") && atoi(getenv("MAED_B200_GN_BSHARED")) == 0);
  p.b_shared = (b_shared && tpc == 2 && !pair) ? 1 : 0;
  p.n_blocks = a.C / bn;
  p.items = a.n_img * p.n_blocks;
  if (a.conv) {
    const uint64_t dims[5] = {(uint64_t)a.Cin, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)a.n_img, (uint64_t)np};
    const uint64_t str[4] = {(uint64_t)a.Cin * 2, (uint64_t)p.W * a.Cin * 2, (uint64_t)p.H * p.W * a.Cin * 2,
                             (uint64_t)(np == 2 ? a.a_plane : M * a.Cin) * 2};
    const uint32_t box[5] = {64, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, a.A, 5, dims, str, box));
  } else {
    const uint64_t dims[3] = {(uint64_t)a.K, (uint64_t)M, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)a.K * 2, (uint64_t)(np == 2 ? a.a_plane : M * a.K) * 2};
    const uint32_t box[3] = {64, 128, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, a.A, 3, dims, str, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)a.C, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)p.K * 2, (uint64_t)(np == 2 ? a.b_plane : (long long)a.C * p.K) * 2};
    const uint32_t box[3] = {64, (uint32_t)(pair ? bn / 2 : bn), 1};      // a pair CTA stages its half of the block
    MAED_PROPAGATE(make_tmap_f16(&tmB, a.B, 3, dims, str, box));
  }
  // output planes [n_img, HW (or H, W), C] x 2 planes; 64-channel x one-tile boxes, 128-byte swizzle
  MAED_CHECK_ARG(a.out_plane > 0, "conv_gn_fused: output planes required");
  if (a.conv) {
    const uint64_t dims[5] = {(uint64_t)a.C, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)a.n_img, 2};
    const uint64_t str[4] = {(uint64_t)a.C * 2, (uint64_t)p.W * a.C * 2, (uint64_t)p.H * p.W * a.C * 2, (uint64_t)a.out_plane * 2};
    const uint32_t box[5] = {32, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1};
    p.o_tx_bytes = (uint32_t)(p.tile_w * p.tile_h * 64);
    MAED_PROPAGATE(make_tmap_f16(&tmO, a.out, 5, dims, str, box, 64));
    { const uint32_t wb[5] = {64, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1}; MAED_PROPAGATE(make_tmap_f16(&tmOw, a.out, 5, dims, str, wb, 128)); }
    if (a.res) {
      const uint64_t rstr[4] = {str[0], str[1], str[2], (uint64_t)a.res_plane * 2};
      MAED_PROPAGATE(make_tmap_f16(&tmR, a.res, 5, dims, rstr, box, 64));
    }
  } else {
    const uint64_t dims[4] = {(uint64_t)a.C, (uint64_t)p.HW, (uint64_t)a.n_img, 2};
    const uint64_t str[3] = {(uint64_t)a.C * 2, (uint64_t)p.HW * a.C * 2, (uint64_t)a.out_plane * 2};
    const uint32_t box[4] = {32, 128, 1, 1};
    p.o_tx_bytes = 128 * 64;
    MAED_PROPAGATE(make_tmap_f16(&tmO, a.out, 4, dims, str, box, 64));
    { const uint32_t wb[4] = {64, 128, 1, 1}; MAED_PROPAGATE(make_tmap_f16(&tmOw, a.out, 4, dims, str, wb, 128)); }
    if (a.res) {
      const uint64_t rstr[3] = {str[0], str[1], (uint64_t)a.res_plane * 2};
      MAED_PROPAGATE(make_tmap_f16(&tmR, a.res, 4, dims, rstr, box, 64));
    }
  }
  if (!a.res) tmR = tmO;
  if (pair && gsz == 8) return launch_gn<256, 8, true>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (pair && gsz == 32) return launch_gn<128, 32, true>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 128 && gsz == 4) return launch_gn<128, 4>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 128 && gsz == 8) return launch_gn<128, 8>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 128 && gsz == 16) return launch_gn<128, 16>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 128 && gsz == 32) return launch_gn<128, 32>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 64 && gsz == 2) return launch_gn<64, 2>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 64 && gsz == 4) return launch_gn<64, 4>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 64 && gsz == 8) return launch_gn<64, 8>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 64 && gsz == 16) return launch_gn<64, 16>(tmA, tmB, tmO, tmR, tmOw, p, st);
  if (bn == 64 && gsz == 32) return launch_gn<64, 32>(tmA, tmB, tmO, tmR, tmOw, p, st);
  return MAED_ERR_UNSUPPORTED;
}

}  // namespace maed
