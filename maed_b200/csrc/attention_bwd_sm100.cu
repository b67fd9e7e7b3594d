// Backward of the per-frame spatial attention (reference lib/models/vision_transformer.py:206-214 under autograd) on the
// 5th-gen tensor cores.  One CTA per (frame, head) item, persistent over the items; everything between the q/k/v/dO tiles
// and the three gradient tiles stays in shared memory / TMEM:
//
//     S = Q K^T * scale        P = softmax_j(S)        O = P V                                  (forward, recomputed)
//     dP = dO V^T              D_i = sum_j P_ij dP_ij  dS = P o (dP - D_i) * scale
//     dQ = dS K                dK = dS^T Q             dV = P^T dO
//
// tcgen05.mma takes A from shared memory or TMEM with M on the TMEM lanes, so products that contract over the queries
// (dK, dV) need P^T / dS^T with the KEYS on the lanes.  Instead of transposing through shared memory the kernel computes both
// orientations with the tensor cores (their FLOPs are negligible here):
//
//   orientation A (queries on lanes, per 128-query tile):  S, dP -> TMEM; the element-wise warps take the row max / sum /
//       D_i (two threads per row, like the forward kernel), leave lse_i and D_i in shared memory and write dS back over dP
//       as fp16 hi/lo pairs; dQ = dS (TMEM operand) x K (MN-major shared-memory operand).
//   orientation B (keys on lanes, per 128-key tile, queries in two column chunks):  S^T = K Q^T, dP^T = V dO^T -> TMEM;
//       P^T = exp2(S^T c - lse_i), dS^T = P^T o (dP^T - D_i) scale — purely element-wise with the per-column statistics of
//       orientation A — written back in place; dV += P^T x dO, dK += dS^T x Q (TMEM operand x MN-major operand).
//
// Split precision as everywhere on the path: every operand is an fp16 hi/lo pair, three MMAs per product (hi*hi + lo*hi +
// hi*lo), fp32 accumulation in TMEM, so the gradients are fp32-class (tests: 2e-5 against float64 autograd).
// The same kernel serves the TEMPORAL attention (vision_transformer.py:216-228; template TEMPORAL): an item is then (clip, head,
// group of NPT = 128 / T tokens), its operand tiles are TMA boxes {64, NPT, T} of the qkv / dO planes (row r = t * NPT + n, see
// attention_temporal_sm100.cu) and "valid key of query row i" means "same token" (c mod NPT == i mod NPT) instead of c < ntok.
// First version: the phases of an item run one after the other (one thread issues the TMA loads and the MMAs, eight warps
// do the element-wise work); round-2 measurement decides whether the phases need software pipelining like the forward kernel.
#include "bwd_kernels.h"

#include "device_utils.cuh"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

namespace {

constexpr int kD = 64;                 // head dim
constexpr int kRows = 208;             // token rows fetched per operand (ntok <= 208), multiple of 16
constexpr int kBufBytes = kRows * 128; // one plane of one operand: 208 rows x 128 B (K-major, 128-byte swizzle)
constexpr int kThreadsB = 384;         // warp 0: TMA + MMA issue, warp 1: TMEM alloc, warps 4-11: element-wise (2 threads / row)
constexpr int kSplit = 112;            // column split: [0,112) | [112,208)  (orientation A halves, orientation B chunks)
// TMEM column maps
constexpr uint32_t kS = 0, kP = 208, kDQ = 416;            // orientation A: S | dP -> dS | dQ  (temporal: 128 of the 208 columns)
constexpr uint32_t kST = 0, kPT = 112, kDV = 224, kDK = 288;   // orientation B: S^T -> P^T | dP^T -> dS^T | dV | dK

struct BwdParams {
  int BT, ntok, heads;
  int T, groups;                       // temporal: frames per clip, token groups per (clip, head); BT = clips
  float scale, scale_log2e;
  float* d_qkv;                        // fp32 [BT*ntok, 3*heads*64]
  int accumulate;
  // optional, both [rows, heads]: the forward's log2-domain row log-sum-exp and D = rowsum(dO o O).  When given, orientation A is
  // ONE element-wise pass (dS = exp2(S c - lse) (dP - D) scale) instead of three (row max, row sum + D, dS) with two exchanges
  const float* lse_in; const float* d_in;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// 16 consecutive fp32 values -> 8 packed fp16 hi words | 8 packed fp16 lo words (the TMEM A-operand layout of a 16-deep K step)
__device__ __forceinline__ void pack16(const float (&v)[16], uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 2) {
    const __half2 h2 = __floats2half2_rn(v[j], v[j + 1]);
    const float2 hf = __half22float2(h2);
    const __half2 l2 = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
    pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
    pk[8 + (j >> 1)] = *reinterpret_cast<const uint32_t*>(&l2);
  }
}

template <bool TEMPORAL, int NPT>
__global__ void __launch_bounds__(kThreadsB, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO, const BwdParams p) {
  using namespace sm100;
  constexpr int ROWS = TEMPORAL ? 128 : kRows;            // key rows / score columns of an item
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // operand buffers [hi | lo]; an M = 128 tile starting at row 128 reads 48 rows past its buffer into the next one (finite fp16
  // data or the zero pad): those rows only feed accumulator lanes that are never stored
  uint8_t* sDO = smem;
  uint8_t* sQ = sDO + 2 * kBufBytes;
  uint8_t* sK = sQ + 2 * kBufBytes;
  uint8_t* sV = sK + 2 * kBufBytes;
  uint8_t* sPad = sV + 2 * kBufBytes;                     // 6 KB of zeros behind V_lo
  float* lse = reinterpret_cast<float*>(sPad + 6144);     // [256] log2-domain log-sum-exp of every query row
  float* Dr = lse + 256;                                  // [256] D_i
  float* xch = Dr + 256;                                  // [2 halves][128 rows][4]: {max, sum, dot}
  uint64_t* bars = reinterpret_cast<uint64_t*>(xch + 2 * 128 * 4);
  uint64_t* ld_full = bars + 0;
  uint64_t* mma_done = bars + 1;
  uint64_t* ew_done = bars + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = TEMPORAL ? p.BT * p.heads * p.groups : p.BT * p.heads;
  const int ntok = p.ntok;
  const int tiles = TEMPORAL ? 1 : (ntok + 127) / 128;    // 128-row tiles of queries (orientation A) / keys (orientation B)
  const int chunks = (TEMPORAL || ntok > kSplit) ? 2 : 1; // query column chunks of orientation B
  const int ld3 = 3 * p.heads * kD;

  for (int i = threadIdx.x; i < 6144 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sPad)[i] = 0u;
  if (TEMPORAL)   // token groups at the end of a frame leave part of a tile unwritten by TMA: start from finite (zero) operands
    for (int i = threadIdx.x; i < 8 * kBufBytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0 && elect_one()) { prefetch_tmap(&tmQKV); prefetch_tmap(&tmDO); }
  if (warp == 1 && elect_one()) {
    mbar_init(ld_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(ew_done, 8);
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_base_ptr, 512); tmem_relinquish(); }
  fence_proxy_async();                                    // the zero pad is read by the tensor cores (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (elect_one()) {
      const uint32_t aDO = smem_u32(sDO), aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      uint32_t ph_ld = 0, ph_ew = 0;
      // three MMAs of one split-precision product, both operands K-major in shared memory
      auto mma_ss = [&](uint32_t d, uint32_t a, uint32_t b, uint32_t idesc) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t ah = umma_desc_k_sw128(a + k * 32), al = umma_desc_k_sw128(a + kBufBytes + k * 32);
          const uint64_t bh = umma_desc_k_sw128(b + k * 32), bl = umma_desc_k_sw128(b + kBufBytes + k * 32);
          umma_f16(d, ah, bh, idesc, k != 0);
          umma_f16(d, al, bh, idesc, 1);
          umma_f16(d, ah, bl, idesc, 1);
        }
      };
      // D (+)= A[tmem: fp16 hi/lo pairs, K = 16 * ksteps] x B[smem rows b.., MN-major: N = 64 head-dim columns]
      auto mma_ts = [&](uint32_t d, uint32_t a_tmem, uint32_t b, int ksteps, bool acc, uint32_t idesc) {
#pragma unroll 1
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t bh = umma_desc_mn_sw128(b + kk * 2048, 1024, 1024);
          const uint64_t bl = umma_desc_mn_sw128(b + kBufBytes + kk * 2048, 1024, 1024);
          umma_f16_ts(d, a_tmem + kk * 16, bh, idesc, (acc || kk != 0) ? 1u : 0u);
          umma_f16_ts(d, a_tmem + kk * 16 + 8, bh, idesc, 1);
          umma_f16_ts(d, a_tmem + kk * 16, bl, idesc, 1);
        }
      };
      constexpr uint32_t id208 = umma_idesc_f16(128, ROWS, 0, 0, 0);
      constexpr uint32_t id112 = umma_idesc_f16(128, kSplit, 0, 0, 0);
      constexpr uint32_t id96 = umma_idesc_f16(128, ROWS - kSplit, 0, 0, 0);
      constexpr uint32_t id64 = umma_idesc_f16(128, kD, 0, 0, 1);          // B MN-major
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        // ---- operands of the item (the previous item's MMAs have all retired: its last ew_done was waited for below)
        if (TEMPORAL) {
          const int g = item % p.groups, h = (item / p.groups) % p.heads, b = item / (p.groups * p.heads);
          mbar_arrive_expect_tx(ld_full, 8 * 128 * 128);
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_5d(sDO + pl * kBufBytes, &tmDO, ld_full, h * kD, g * NPT, 0, b, pl);
            tma_load_5d(sQ + pl * kBufBytes, &tmQKV, ld_full, h * kD, g * NPT, 0, b, pl);
            tma_load_5d(sK + pl * kBufBytes, &tmQKV, ld_full, p.heads * kD + h * kD, g * NPT, 0, b, pl);
            tma_load_5d(sV + pl * kBufBytes, &tmQKV, ld_full, 2 * p.heads * kD + h * kD, g * NPT, 0, b, pl);
          }
        } else {
          const int bt = item / p.heads, h = item % p.heads;
          const int row0 = bt * ntok;
          mbar_arrive_expect_tx(ld_full, 8 * kBufBytes);
          for (int pl = 0; pl < 2; ++pl) {
            tma_load_3d(sDO + pl * kBufBytes, &tmDO, ld_full, h * kD, row0, pl);
            tma_load_3d(sQ + pl * kBufBytes, &tmQKV, ld_full, h * kD, row0, pl);
            tma_load_3d(sK + pl * kBufBytes, &tmQKV, ld_full, p.heads * kD + h * kD, row0, pl);
            tma_load_3d(sV + pl * kBufBytes, &tmQKV, ld_full, 2 * p.heads * kD + h * kD, row0, pl);
          }
        }
        mbar_wait(ld_full, ph_ld);
        ph_ld ^= 1;
        tc_fence_after();
        // ---- orientation A: queries on the TMEM lanes
        for (int g = 0; g < tiles; ++g) {
          mma_ss(tmem_base + kS, aQ + g * 16384, aK, id208);               // S_g  = Q_g K^T
          mma_ss(tmem_base + kP, aDO + g * 16384, aV, id208);              // dP_g = dO_g V^T
          umma_commit(mma_done);
          mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                           // dS_g in TMEM (over dP), lse / D in shared memory
          tc_fence_after();
          mma_ts(tmem_base + kDQ, tmem_base + kP, aK, ROWS / 16, false, id64);    // dQ_g = dS_g K
          umma_commit(mma_done);
          mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                           // dQ_g has left TMEM
          tc_fence_after();
        }
        // ---- orientation B: keys on the TMEM lanes, queries in column chunks
        for (int t = 0; t < tiles; ++t) {
          for (int c = 0; c < chunks; ++c) {
            const int i0 = c * kSplit;
            const uint32_t idn = c == 0 ? id112 : id96;
            mma_ss(tmem_base + kST, aK + t * 16384, aQ + i0 * 128, idn);   // S^T  = K_t Q_c^T
            mma_ss(tmem_base + kPT, aV + t * 16384, aDO + i0 * 128, idn);  // dP^T = V_t dO_c^T
            umma_commit(mma_done);
            mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                         // P^T, dS^T written back in place
            tc_fence_after();
            const int ks = (c == 0 ? kSplit : ROWS - kSplit) / 16;
            mma_ts(tmem_base + kDV, tmem_base + kST, aDO + i0 * 128, ks, c != 0, id64);   // dV_t += P^T dO_c
            mma_ts(tmem_base + kDK, tmem_base + kPT, aQ + i0 * 128, ks, c != 0, id64);    // dK_t += dS^T Q_c
            // (tcgen05.mma executes in issue order: the next chunk's S^T / dP^T may overwrite these operands behind them)
          }
          umma_commit(mma_done);
          mbar_wait(ew_done, ph_ew); ph_ew ^= 1;                           // dV_t, dK_t have left TMEM
          tc_fence_after();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ==================================================================== element-wise warps, two threads per TMEM lane
    const int half = (warp - 4) >> 2;
    const int wq = warp & 3;                              // TMEM lane quarter = warp id % 4
    const int trow = wq * 32 + lane;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    float* x_mine = xch + (half * 128 + trow) * 4;
    const float* x_peer = xch + ((half ^ 1) * 128 + trow) * 4;
    const uint32_t pair_bar = 1 + wq;
    const float c2 = p.scale_log2e;
    uint32_t ph_mma = 0;
    const int my_nl = trow % NPT;                         // temporal: token of this thread's tile row within the group
    // is score column `col` (key index in orientation A, query index in orientation B) paired with this thread's row?
    auto valid = [&](int col) -> bool { return TEMPORAL ? ((col % NPT) == my_nl) : (col < ntok); };
    auto arrive_ew = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ew_done);
    };
    // 32 fp32 values of this thread's accumulator row -> global (row-contiguous 128 bytes)
    auto store32 = [&](const uint32_t (&r)[32], float* dst) {
      float4* o = reinterpret_cast<float4*>(dst);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                               __uint_as_float(r[4 * j + 3]));
        if (p.accumulate) { const float4 w = o[j]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
        o[j] = v;
      }
    };
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      // global row of this thread's tile row `r` (query row in orientation A, key row in orientation B), or -1
      int h, tg = 0, tb = 0;
      long long row_base;
      if (TEMPORAL) {
        tg = item % p.groups; h = (item / p.groups) % p.heads; tb = item / (p.groups * p.heads);
        row_base = 0;
      } else {
        const int bt = item / p.heads;
        h = item % p.heads;
        row_base = (long long)bt * ntok;
      }
      auto grow = [&](int r) -> long long {
        if (TEMPORAL) {
          const int n = tg * NPT + r % NPT, t = r / NPT;
          return (n < ntok && t < p.T) ? ((long long)tb * p.T + t) * ntok + n : -1;
        }
        return r < ntok ? row_base + r : -1;
      };
      float* out_item = p.d_qkv + h * kD + half * 32;
      // ------------------------------------------------------------------------------------------ orientation A
      for (int g = 0; g < tiles; ++g) {
        const uint32_t tS = tmem_base + kS + lane_off, tP = tmem_base + kP + lane_off, tQ = tmem_base + kDQ + lane_off;
        const int cbeg = half ? kSplit : 0, cend = half ? ROWS : kSplit;
        mbar_wait(mma_done, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        uint32_t r[32], q[32];
        float mb, Di, ps;
        if (p.lse_in != nullptr) {
          // statistics saved by the forward kernel (lse) and precomputed from dO and O (D): no reduction passes, no exchange
          const long long qrow_s = grow(g * 128 + trow);
          mb = qrow_s >= 0 ? p.lse_in[qrow_s * p.heads + h] : INFINITY;       // rows outside the item: exp2(-inf) = 0
          Di = qrow_s >= 0 ? p.d_in[qrow_s * p.heads + h] : 0.f;
          ps = p.scale;
          if (half == 0) {
            lse[g * 128 + trow] = mb;
            Dr[g * 128 + trow] = Di;
          }
        } else {
        // pass 1: row max of the raw scores over the valid keys of this half
        float mx = -INFINITY;
#pragma unroll 1
        for (int col0 = cbeg; col0 < cend; col0 += 32) {
          const int ncols = min(32, cend - col0);
          if (ncols == 32) tmem_ld_32x32b_x32(tS + col0, r);
          else tmem_ld_32x32b_x16(tS + col0, reinterpret_cast<uint32_t(&)[16]>(r[0]));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols && valid(col0 + j)) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        x_mine[0] = mx;
        bar_sync(pair_bar, 64);
        mx = fmaxf(mx, x_peer[0]);
        mb = mx * c2;
        // pass 2: row sum of e = exp2(s c - max c) and sum of e * dP
        float sum = 0.f, dot = 0.f;
#pragma unroll 1
        for (int col0 = cbeg; col0 < cend; col0 += 32) {
          const int ncols = min(32, cend - col0);
          if (ncols == 32) { tmem_ld_32x32b_x32(tS + col0, r); tmem_ld_32x32b_x32(tP + col0, q); }
          else {
            tmem_ld_32x32b_x16(tS + col0, reinterpret_cast<uint32_t(&)[16]>(r[0]));
            tmem_ld_32x32b_x16(tP + col0, reinterpret_cast<uint32_t(&)[16]>(q[0]));
          }
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols && valid(col0 + j)) {
              const float e = ex2(__uint_as_float(r[j]) * c2 - mb);
              sum += e;
              dot = fmaf(e, __uint_as_float(q[j]), dot);
            }
        }
        x_mine[1] = sum;
        x_mine[2] = dot;
        bar_sync(pair_bar, 64);
        sum += x_peer[1];
        dot += x_peer[2];
        const float inv = 1.0f / sum;
        Di = dot * inv;
        if (half == 0) {
          lse[g * 128 + trow] = mb + log2f(sum);
          Dr[g * 128 + trow] = Di;
        }
        ps = inv * p.scale;
        }
        // pass 3: dS = (e / sum) (dP - D_i) scale, written over dP as fp16 hi/lo pairs (zero for padded keys); 16 columns a time
#pragma unroll 1
        for (int col0 = cbeg; col0 < cend; col0 += 16) {
          uint32_t r16[16], q16[16], pk[16];
          float v[16];
          tmem_ld_32x32b_x16(tS + col0, r16);
          tmem_ld_32x32b_x16(tP + col0, q16);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float e = ex2(__uint_as_float(r16[j]) * c2 - mb);
            v[j] = valid(col0 + j) ? e * ps * (__uint_as_float(q16[j]) - Di) : 0.f;
          }
          pack16(v, pk);
          tmem_st_32x32b_x16(tP + col0, pk);
        }
        tmem_st_wait();
        __threadfence_block();                            // lse / D visible to every element-wise warp after the next barrier round
        arrive_ew();
        // dQ_g: this thread's query row, head-dim columns [32 half, 32 half + 32)
        mbar_wait(mma_done, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        tmem_ld_32x32b_x32(tQ + half * 32, r);
        tmem_ld_wait();
        arrive_ew();
        const long long qrow = grow(g * 128 + trow);
        if (qrow >= 0) store32(r, out_item + qrow * ld3);
      }
      // ------------------------------------------------------------------------------------------ orientation B
      bar_sync(5, 256);                                   // lse / D of every query row written (all element-wise warps)
      for (int t = 0; t < tiles; ++t) {
        const uint32_t tST = tmem_base + kST + lane_off, tPT = tmem_base + kPT + lane_off;
        for (int c = 0; c < chunks; ++c) {
          const int i0 = c * kSplit;
          const int ni = c == 0 ? kSplit : ROWS - kSplit;             // 112 | 96 (temporal: 16) query columns
          const int cbeg = half ? 64 : 0, cend = half ? ni : min(64, ni);
          mbar_wait(mma_done, ph_mma); ph_mma ^= 1;
          tc_fence_after();
#pragma unroll 1
          for (int col0 = cbeg; col0 < cend; col0 += 16) {
            uint32_t r16[16], q16[16], pk[16];
            float pv[16], dv[16];
            tmem_ld_32x32b_x16(tST + col0, r16);
            tmem_ld_32x32b_x16(tPT + col0, q16);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
              const int i = i0 + col0 + j4;                           // query index of the first of 4 columns (multiple of 4)
              const float4 l4 = *reinterpret_cast<const float4*>(lse + i);
              const float4 d4 = *reinterpret_cast<const float4*>(Dr + i);
              const float ls[4] = {l4.x, l4.y, l4.z, l4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const bool ok = valid(i + u);
                const float pr = ok ? ex2(__uint_as_float(r16[j4 + u]) * c2 - ls[u]) : 0.f;
                pv[j4 + u] = pr;
                dv[j4 + u] = ok ? pr * p.scale * (__uint_as_float(q16[j4 + u]) - dd[u]) : 0.f;
              }
            }
            pack16(pv, pk);
            tmem_st_32x32b_x16(tST + col0, pk);
            pack16(dv, pk);
            tmem_st_32x32b_x16(tPT + col0, pk);
          }
          tmem_st_wait();
          arrive_ew();
        }
        // dV_t / dK_t: this thread's key row
        mbar_wait(mma_done, ph_mma); ph_mma ^= 1;
        tc_fence_after();
        uint32_t rv[32], rk[32];
        tmem_ld_32x32b_x32(tmem_base + kDV + lane_off + half * 32, rv);
        tmem_ld_32x32b_x32(tmem_base + kDK + lane_off + half * 32, rk);
        tmem_ld_wait();
        arrive_ew();
        const long long krow = grow(t * 128 + trow);
        if (krow >= 0) {
          store32(rk, out_item + krow * ld3 + p.heads * kD);
          store32(rv, out_item + krow * ld3 + 2 * p.heads * kD);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

constexpr size_t kBwdSmem = 1024 + 8 * (size_t)kBufBytes + 6144 + (256 + 256 + 2 * 128 * 4) * sizeof(float) + 64;

}  // namespace

int attn_spatial_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int BT, int ntok,
                        int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse, const float* Dv) {
  MAED_CHECK_ARG((lse == nullptr) == (Dv == nullptr), "attn_spatial_bwd_tc: lse and D come together");
  MAED_CHECK_ARG(qkv_hi && dout_hi && d_qkv, "attn_spatial_bwd_tc: null argument");
  MAED_CHECK_ARG(ntok >= 1 && ntok <= kRows, "attn_spatial_bwd_tc: ntok=%d unsupported (1..%d)", ntok, kRows);
  MAED_CHECK_ARG(BT >= 1 && heads >= 1, "attn_spatial_bwd_tc: bad batch");
  const long long rows = (long long)BT * ntok;
  const int ld3 = 3 * heads * kD, ldo = heads * kD;
  MAED_CHECK_ARG(qkv_plane >= rows * ld3 && dout_plane >= rows * ldo && qkv_plane % 8 == 0 && dout_plane % 8 == 0,
                 "attn_spatial_bwd_tc: operand planes overlap or are misaligned");
  CUtensorMap tmQKV, tmDO;
  {
    const uint64_t dims[3] = {(uint64_t)ld3, (uint64_t)rows, 2};
    const uint64_t str[2] = {(uint64_t)ld3 * 2, (uint64_t)qkv_plane * 2};
    const uint32_t box[3] = {64, kRows, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmQKV, qkv_hi, 3, dims, str, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)ldo, (uint64_t)rows, 2};
    const uint64_t str[2] = {(uint64_t)ldo * 2, (uint64_t)dout_plane * 2};
    const uint32_t box[3] = {64, kRows, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmDO, dout_hi, 3, dims, str, box));
  }
  BwdParams p;
  p.BT = BT; p.ntok = ntok; p.heads = heads; p.T = 1; p.groups = 1; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.d_qkv = d_qkv; p.accumulate = accumulate; p.lse_in = lse; p.d_in = Dv;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    attr_set = true;
  }
  const int items = BT * heads;
  const int grid = items < sm_count() ? items : sm_count();
  attn_bwd_tc_kernel<false, 1><<<grid, kThreadsB, kBwdSmem, st>>>(tmQKV, tmDO, p);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

template <int NPT>
static int launch_temporal_bwd(const CUtensorMap& tmQKV, const CUtensorMap& tmDO, const BwdParams& p, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_tc_kernel<true, NPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwdSmem));
    attr_set = true;
  }
  const int items = p.BT * p.heads * p.groups;
  const int grid = items < sm_count() ? items : sm_count();
  attn_bwd_tc_kernel<true, NPT><<<grid, kThreadsB, kBwdSmem, st>>>(tmQKV, tmDO, p);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// temporal attention backward on the same kernel: B clips, T frames (4, 8, 16, 32), d_out as fp16 hi/lo planes [B*T*ntok, heads*64]
int attn_temporal_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int B, int T,
                         int ntok, int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse,
                         const float* Dv) {
  MAED_CHECK_ARG((lse == nullptr) == (Dv == nullptr), "attn_temporal_bwd_tc: lse and D come together");
  MAED_CHECK_ARG(qkv_hi && dout_hi && d_qkv, "attn_temporal_bwd_tc: null argument");
  MAED_CHECK_ARG(T == 4 || T == 8 || T == 16 || T == 32, "attn_temporal_bwd_tc: T=%d unsupported (4, 8, 16, 32)", T);
  const long long rows = (long long)B * T * ntok;
  const int ld3 = 3 * heads * kD, ldo = heads * kD, npt = 128 / T;
  MAED_CHECK_ARG(qkv_plane >= rows * ld3 && dout_plane >= rows * ldo && qkv_plane % 8 == 0 && dout_plane % 8 == 0,
                 "attn_temporal_bwd_tc: operand planes overlap or are misaligned");
  CUtensorMap tmQKV, tmDO;
  auto tmap5 = [&](CUtensorMap* tm, const __half* base, long long plane, int ld) -> int {
    const uint64_t dims[5] = {(uint64_t)ld, (uint64_t)ntok, (uint64_t)T, (uint64_t)B, 2};
    const uint64_t str[4] = {(uint64_t)ld * 2, (uint64_t)ntok * ld * 2, (uint64_t)T * ntok * ld * 2, (uint64_t)plane * 2};
    const uint32_t box[5] = {64, (uint32_t)npt, (uint32_t)T, 1, 1};
    return make_tmap_f16(tm, base, 5, dims, str, box);
  };
  MAED_PROPAGATE(tmap5(&tmQKV, qkv_hi, qkv_plane, ld3));
  MAED_PROPAGATE(tmap5(&tmDO, dout_hi, dout_plane, ldo));
  BwdParams p;
  p.BT = B; p.ntok = ntok; p.heads = heads; p.T = T; p.groups = (ntok + npt - 1) / npt; p.scale = scale;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.d_qkv = d_qkv; p.accumulate = accumulate; p.lse_in = lse; p.d_in = Dv;
  switch (npt) {
    case 32: return launch_temporal_bwd<32>(tmQKV, tmDO, p, st);
    case 16: return launch_temporal_bwd<16>(tmQKV, tmDO, p, st);
    case 8: return launch_temporal_bwd<8>(tmQKV, tmDO, p, st);
    default: return launch_temporal_bwd<4>(tmQKV, tmDO, p, st);
  }
}

}  // namespace maed
