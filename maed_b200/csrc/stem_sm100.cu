// Stem convolution of the ResNetV2 backbone (reference resnetv2.py:245-274: StdConv2dSame 7x7 / stride 2,
// 3 -> 64 channels, 224x224 -> 112x112) as a tcgen05 implicit GEMM with the im2col tile staged in shared
// memory by the CTA itself:
//   * one tile = one output row (112 pixels, M = 128 with 16 dead rows) x 64 channels, K = 147 (+13 zeros);
//   * the 7 input rows x 3 channels the tile needs are fetched once with cp.async (coalesced 16-byte loads of
//     the fp32 NCHW frame, double buffered), then 128 builder threads write the A operand straight into the
//     128-byte-swizzled K-major layout the UMMA descriptor expects, as fp16 hi/lo planes;
//   * the standardised weights (64 x 160, hi/lo) stay resident in shared memory for the whole kernel (one TMA load);
//   * the epilogue transposes through shared memory for fully coalesced fp32 stores and accumulates the
//     GroupNorm statistics (32 groups of 2 channels) in registers, flushed with fp64 atomics per image.
// This replaces a 976 MB explicit im2col matrix (write + read) of the first version.
#include "gemm_host.h"
#include "kernels.h"
#include "sm100_ptx.cuh"

namespace maed {

static constexpr int kStemThreads = 384;      // warps 0-7 builders (+MMA issue by thread 0), warps 8-11 epilogue
static constexpr int kStemBuilders = 256;
static constexpr int kOW = 112, kOH = 112, kIW = 224, kIH = 224;
static constexpr int kPatchW = 232;            // 2 zero columns left, 224 data, 6 right (3 needed)
static constexpr int kPatchFloats = 3 * 7 * kPatchW;
static constexpr int kKReal = 147, kKSteps = 10;   // 10 x 16 = 160 = 147 + 13 zeros

struct StemParams {
  const float* x;        // [n_img, 3, 224, 224]
  float* out;            // [n_img * 112 * 112, 64] fp32
  double* stats;         // [n_img][32][2] (sum, sumsq), pre-zeroed
  int n_img;
  int nsplit;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm100::smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// patch offset of im2col column k (k = (r*7 + s)*3 + c), or -1 for the zero padding columns 147..159
__host__ __device__ constexpr int stem_patch_off(int k) {
  return k < kKReal ? ((k % 3) * 7 + (k / 3) / 7) * kPatchW + (k / 3) % 7 : -1;
}
// Builds chunks [CH0, CH0 + 10) of A row `row` (pixel ow = row): all gather offsets are compile-time immediates.
template <int CH0>
__device__ __forceinline__ void stem_build_row(const float* __restrict__ pb, bool valid, int row, uint8_t* sA, uint32_t a_plane,
                                               int np) {
#pragma unroll
  for (int ch = CH0; ch < CH0 + 10; ++ch) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      constexpr int dummy = 0; (void)dummy;
      const int off = stem_patch_off(ch * 8 + j);
      v[j] = (valid && off >= 0) ? pb[off >= 0 ? off : 0] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      const __half2 h2 = __floats2half2_rn(v[j], v[j + 1]);
      const float2 hf = __half22float2(h2);
      const __half2 l2 = __floats2half2_rn(v[j] - hf.x, v[j + 1] - hf.y);
      hi[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      lo[j >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
    }
    const int kb = ch >> 3, cin = ch & 7;                 // K block, 16-byte chunk inside the 128-byte row
    const uint32_t off = kb * (128 * 128) + row * 128 + ((cin ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(sA + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (np == 2) *reinterpret_cast<uint4*>(sA + a_plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

__global__ void __launch_bounds__(kStemThreads, 1)
stem_conv_tc_kernel(const __grid_constant__ CUtensorMap tmW, const StemParams p) {
  using namespace sm100;
  constexpr uint32_t kAKb = 128 * 128;        // one K block (64 halfs) of the A tile
  constexpr uint32_t kBKb = 64 * 128;
  constexpr uint32_t kAPlane = 3 * kAKb, kBPlane = 3 * kBKb;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (STS/LDS, not generic ST/LD)
  uint8_t* sA = smem;                                   // [2 planes][3 kb][128 x 128 B]
  uint8_t* sB = sA + 2 * kAPlane;                       // [2 planes][3 kb][64 x 128 B]
  float* sPatch = reinterpret_cast<float*>(sB + 2 * kBPlane);   // [2][3][7][232]
  float* sStage = sPatch + 2 * kPatchFloats;            // [4 warps][32 rows x 64 floats]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + 4 * 32 * 64);
  uint64_t* w_full = bars + 0;
  uint64_t* mma_done = bars + 1;
  uint64_t* tmem_full = bars + 2;    // [2]
  uint64_t* tmem_empty = bars + 4;   // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = p.nsplit == 3 ? 2 : 1;
  const int total_tiles = p.n_img * kOH;
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;       // contiguous tile ranges: few image changes
  const int tile_lo = blockIdx.x * per_cta, tile_hi = min(total_tiles, tile_lo + per_cta);

  if (warp == 1 && elect_one()) {
    mbar_init(w_full, 1);
    mbar_init(mma_done, 1);
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_base_ptr, 128); tmem_relinquish(); }
  // zero the patch buffers once (pad columns stay zero forever)
  for (int i = threadIdx.x; i < 2 * kPatchFloats; i += kStemThreads) sPatch[i] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;
  if (tile_lo >= tile_hi) {                    // idle CTA (more CTAs than tiles)
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
    return;
  }

  if (warp == 0 && elect_one()) {              // resident weights: 3 K blocks x planes
    prefetch_tmap(&tmW);
    mbar_arrive_expect_tx(w_full, np * kBPlane);
    for (int pl = 0; pl < np; ++pl)
      for (int kb = 0; kb < 3; ++kb) tma_load_3d(sB + pl * kBPlane + kb * kBKb, &tmW, w_full, kb * 64, 0, pl);
  }

  if (warp < 8) {
    // =================================================================== builders (+ MMA issue by thread 0)
    const int tid = threadIdx.x;               // 0..255: row = tid & 127 (output pixel ow; >= 112 -> zero rows), half = tid >> 7
    const int brow = tid & 127, bhalf = tid >> 7;
    auto issue_patch = [&](int tile, int buf) {
      const int img = tile / kOH, oh = tile % kOH;
      float* dst = sPatch + buf * kPatchFloats;
      // 21 (c, r) rows x 56 float4
      for (int i = tid; i < 21 * 56; i += kStemBuilders) {
        const int row = i / 56, q = i % 56;
        const int c = row / 7, r = row % 7;
        const int ih = 2 * oh + r - 2;
        float* d = dst + row * kPatchW + 2 + q * 4;          // data starts at column 2 -> 8-byte aligned only
        if (ih >= 0 && ih < kIH) {
          const float* src = p.x + (((long long)img * 3 + c) * kIH + ih) * kIW + q * 4;
          // destination is only 8-byte aligned (2-float left pad): use two 8-byte async copies
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(d)), "l"(src) : "memory");
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(d + 2)), "l"(src + 2) : "memory");
        } else {
          d[0] = 0.f; d[1] = 0.f; d[2] = 0.f; d[3] = 0.f;
        }
      }
      cp_async_commit();
    };
    issue_patch(tile_lo, 0);
    uint32_t done_phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
      const int buf = it & 1;
      cp_async_wait_all();
      named_bar_sync(1, kStemBuilders);                        // patch(tile) visible to all builders
      if (tile + 1 < tile_hi) issue_patch(tile + 1, buf ^ 1);
      if (it > 0) { mbar_wait(mma_done, done_phase); done_phase ^= 1; }   // A tile free again
      // ---- build A row `brow` (pixel ow): this thread's 10 of the 20 chunks of 8 k-values, swizzled 16-byte stores
      const float* pb = sPatch + buf * kPatchFloats + 2 * brow;
      const bool valid = brow < kOW;
      if (bhalf == 0) stem_build_row<0>(pb, valid, brow, sA, kAPlane, np);
      else stem_build_row<10>(pb, valid, brow, sA, kAPlane, np);
      fence_proxy_async();                                     // generic-proxy writes -> visible to the tensor core
      named_bar_sync(2, kStemBuilders);
      if (tid == 0) {
        if (it == 0) mbar_wait(w_full, 0);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        constexpr uint32_t idesc = umma_idesc_f16(128, 64, 0);
        const uint32_t d = tmem_base + acc * 64;
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
        for (int ks = 0; ks < kKSteps; ++ks) {
          const uint32_t ao = (ks >> 2) * kAKb + (ks & 3) * 32, bo = (ks >> 2) * kBKb + (ks & 3) * 32;
          const uint64_t da = umma_desc_k_sw128(a0 + ao), db = umma_desc_k_sw128(b0 + bo);
          umma_f16(d, da, db, idesc, ks != 0);
          if (np == 2) {
            umma_f16(d, umma_desc_k_sw128(a0 + kAPlane + ao), db, idesc, 1);
            umma_f16(d, da, umma_desc_k_sw128(b0 + kBPlane + bo), idesc, 1);
          }
        }
        umma_commit(mma_done);
        umma_commit(&tmem_full[acc]);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // =========================================================================================== epilogue
    const int ew = warp & 3;
    float* stage = sStage + ew * 32 * 64;
    float gs[32], gq[32];
#pragma unroll
    for (int g = 0; g < 32; ++g) { gs[g] = 0.f; gq[g] = 0.f; }
    int cur_img = tile_lo / kOH;
    auto flush = [&](int img) {
#pragma unroll
      for (int g = 0; g < 32; ++g) {
        float s = gs[g], q = gq[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); q += __shfl_xor_sync(0xffffffffu, q, o); }
        if (lane == 0) {
          atomicAdd(&p.stats[((long long)img * 32 + g) * 2 + 0], (double)s);
          atomicAdd(&p.stats[((long long)img * 32 + g) * 2 + 1], (double)q);
        }
        gs[g] = 0.f; gq[g] = 0.f;
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = tile_lo; tile < tile_hi; ++tile) {
      const int img = tile / kOH, oh = tile % kOH;
      if (img != cur_img) { flush(cur_img); cur_img = img; }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = ew * 32 + lane;                          // pixel ow
      const bool valid = row < kOW;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem_base + acc * 64 + ((uint32_t)(ew * 32) << 16) + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          const float a = __uint_as_float(r[j]), b = __uint_as_float(r[j + 1]);
          if (valid) { gs[(c0 + j) >> 1] += a + b; gq[(c0 + j) >> 1] += a * a + b * b; }
        }
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const int q = (c0 + j) >> 2;                         // logical 16-byte chunk (0..15) of this row
          *reinterpret_cast<uint4*>(stage + lane * 64 + ((q ^ (lane & 7)) << 2)) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      // coalesced copy of this warp's 32 rows x 256 B (contiguous in the NHWC output)
      float* obase = p.out + (((long long)img * kOH + oh) * kOW + ew * 32) * 64;
      const int rows_valid = min(32, kOW - ew * 32);           // warp 3 owns pixels 96..111 only
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int idx = j * 32 + lane;                         // physical 16-byte chunk index in the staging tile
        const int rr = idx >> 4, pq = idx & 15;
        const int lq = pq ^ (rr & 7);
        if (rr < rows_valid)
          *reinterpret_cast<uint4*>(obase + rr * 64 + lq * 4) = *reinterpret_cast<const uint4*>(stage + rr * 64 + pq * 4);
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    flush(cur_img);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

int stem_conv(const float* x, int n_img, const __half* w_hi, long long w_plane, int k_pad, int nsplit, float* out,
              double* stats, cudaStream_t st) {
  MAED_CHECK_ARG(k_pad % 8 == 0 && k_pad >= kKReal, "stem_conv: bad k_pad %d", k_pad);
  CUtensorMap tmW;
  const int np = nsplit == 3 ? 2 : 1;
  const uint64_t dims[3] = {(uint64_t)k_pad, 64, (uint64_t)np};
  const uint64_t str[2] = {(uint64_t)k_pad * 2, (uint64_t)(np == 2 ? w_plane : 64LL * k_pad) * 2};
  const uint32_t box[3] = {64, 64, 1};
  MAED_PROPAGATE(make_tmap_f16(&tmW, w_hi, 3, dims, str, box));
  StemParams p;
  p.x = x; p.out = out; p.stats = stats; p.n_img = n_img; p.nsplit = nsplit;
  const size_t smem = 1024 + 2 * 3 * 128 * 128 + 2 * 3 * 64 * 128 + 2 * kPatchFloats * 4 + 4 * 32 * 64 * 4 + 128;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(stem_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const int tiles = n_img * kOH;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  stem_conv_tc_kernel<<<grid, kStemThreads, smem, st>>>(tmW, p);
  count_launch();
  MAED_CUDA_CHECK(cudaGetLastError());
  return MAED_OK;
}

}  // namespace maed
