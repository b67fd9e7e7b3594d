// Split-K tcgen05 GEMM for the weight gradients of the MAED training path.
//
//   D[Mo, No] (+)= scale * sum_r A[Mo, r] * B[No, r]          A = dY^T, B = X^T (planes)
//
// Two operand layouts: `gemm_wgrad_splitk` takes A and B with r contiguous (K-major UMMA operands, the forward GEMM's layout);
// `gemm_wgrad_rows` (MN = true) takes dY [R, Mo] and X [R, No] as the training tape holds them — rows = r — and feeds them to
// tcgen05.mma as MN-major operands (instruction-descriptor transpose bits, 64 x 64 TMA boxes whose 128-byte swizzled rows hold
// 64 consecutive output rows / columns of one r), so the weight gradients need no transposed copies of dY and X
// (round 2: `transpose_planes` was 11 % of the train step).
//
// The reduction dimension r is the number of activation rows (25 216 tokens ... 1.6 M stem pixels) while Mo x No is a
// weight matrix of at most a few hundred tiles, so the K loop is cut into `S` slices: tile index = (slice, m, n), every
// CTA accumulates its slice in TMEM and writes an fp32 slab; splitk_reduce_kernel sums the slabs in a fixed order
// (deterministic, no atomics).  Same pipeline as gemm_sm100.cuh (TMA producer warp, single-thread MMA issuer,
// double-buffered TMEM accumulators, 4 epilogue warps, split-precision 3-MMA scheme); kept as a separate kernel so the
// validated forward GEMM is untouched.
#include "bwd_kernels.h"

#include "device_utils.cuh"
#include "gemm_host.h"
#include "sm100_ptx.cuh"

namespace maed {

namespace {

struct SplitKParams {
  int M, N;                    // output rows / cols (Mo, No)
  int num_k_blocks;            // ceil(R / 64); conv mode: n_img * patches
  // conv mode (implicit im2col): a K block = one tile_h x tile_w (= 64 pixel) patch of one image
  int conv, patches, tiles_w, tile_h, tile_w, Cin, KW, pad;
  int kb_per_slice, slices;
  int nsplit, stages;
  int m_tiles, n_tiles;
  float* slabs;                // [slices][M][N]
};

constexpr int kBM = 128, kBK = 64, kThreads = 256, kStagesMax = 8;

template <int BLOCK_N, bool MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const SplitKParams p) {
  using namespace sm100;
  constexpr int kAccStages = (2 * BLOCK_N <= 512) ? 2 : 1;
  constexpr int kTmemCols = (kAccStages * BLOCK_N <= 32) ? 32 : (kAccStages * BLOCK_N <= 64) ? 64
                          : (kAccStages * BLOCK_N <= 128) ? 128 : (kAccStages * BLOCK_N <= 256) ? 256 : 512;
  constexpr uint32_t kABytes = kBM * kBK * 2;
  constexpr uint32_t kBBytes = BLOCK_N * kBK * 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nplanes = (p.nsplit == 3) ? 2 : 1;
  const uint32_t stage_bytes = nplanes * (kABytes + kBBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kStagesMax;
  uint64_t* tmem_full = bars + 2 * kStagesMax;
  uint64_t* tmem_empty = bars + 2 * kStagesMax + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStagesMax + 4);

  const int warp = threadIdx.x >> 5;
  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int num_tiles = tiles_mn * p.slices;

  if (warp == 0 && elect_one()) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < kAccStages; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 2) { tmem_alloc(tmem_base_ptr, kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int slice = tile / tiles_mn, mn = tile % tiles_mn;
        const int m_tile = mn / p.n_tiles, n_tile = mn % p.n_tiles;
        const int kb0 = slice * p.kb_per_slice;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_slice);
        // conv mode: everything that needs an integer division is formed once per tile — the tap / channel offset of each
        // 64-column block of the weight row, and the first patch of the slice; later patches follow by counting (this thread
        // must hand out a stage every few hundred tensor cycles)
        int b_cin0[BLOCK_N / 64], b_dw[BLOCK_N / 64], b_dh[BLOCK_N / 64];
        int img = 0, ph = 0, pw = 0;
        if (MN && p.conv) {
#pragma unroll
          for (int nb = 0; nb < BLOCK_N / 64; ++nb) {
            const int col = n_tile * BLOCK_N + nb * 64;
            const int tap = col / p.Cin;
            b_cin0[nb] = col % p.Cin; b_dw[nb] = tap % p.KW - p.pad; b_dh[nb] = tap / p.KW - p.pad;
          }
          const int rem = kb0 % p.patches;
          img = kb0 / p.patches; ph = rem / p.tiles_w; pw = rem % p.tiles_w;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + nplanes * kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], nplanes * (kABytes + kBBytes));
          for (int pl = 0; pl < nplanes; ++pl) {
            if (MN && p.conv) {
              // implicit im2col: dY patch {64 channels, tile_w, tile_h} and, per 64-column block of the [kh][kw][Cin] weight
              // row, the SAME patch of x shifted by the tap (TMA zero-fills the padding and ragged patch borders)
              const int h0 = ph * p.tile_h, w0 = pw * p.tile_w;
#pragma unroll
              for (int mb = 0; mb < kBM / 64; ++mb)
                tma_load_5d(sA + pl * kABytes + mb * 8192, &tmA, &full_bar[stage], m_tile * kBM + mb * 64, w0, h0, img, pl);
#pragma unroll
              for (int nb = 0; nb < BLOCK_N / 64; ++nb) {
                tma_load_5d(sB + pl * kBBytes + nb * 8192, &tmB, &full_bar[stage], b_cin0[nb], w0 + b_dw[nb], h0 + b_dh[nb], img, pl);
              }
            } else if (MN) {       // [64-wide MN block][64 rows of r][128 B]: 8 KB per block, the layout umma_desc_mn_sw128 describes
#pragma unroll
              for (int mb = 0; mb < kBM / 64; ++mb)
                tma_load_3d(sA + pl * kABytes + mb * 8192, &tmA, &full_bar[stage], m_tile * kBM + mb * 64, kb * kBK, pl);
#pragma unroll
              for (int nb = 0; nb < BLOCK_N / 64; ++nb)
                tma_load_3d(sB + pl * kBBytes + nb * 8192, &tmB, &full_bar[stage], n_tile * BLOCK_N + nb * 64, kb * kBK, pl);
            } else {
              tma_load_3d(sA + pl * kABytes, &tmA, &full_bar[stage], kb * kBK, m_tile * kBM, pl);
              tma_load_3d(sB + pl * kBBytes, &tmB, &full_bar[stage], kb * kBK, n_tile * BLOCK_N, pl);
            }
          }
          if (MN && p.conv) {                              // next patch of the image, row-major; then the next image
            if (++pw == p.tiles_w) { pw = 0; if (++ph * p.tiles_w >= p.patches) { ph = 0; ++img; } }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_f16(kBM, BLOCK_N, 0, MN ? 1 : 0, MN ? 1 : 0);
      // operand descriptor of the k-th 16-deep K step of a stage: K-major = 32 bytes along the 128-byte rows; MN-major = 16
      // rows of r (two 8-row swizzle atoms, SBO = 1024 B), 64-wide MN blocks 8 KB apart (LBO)
      auto desc = [](uint32_t base, int k) -> uint64_t {
        return MN ? umma_desc_mn_sw128(base + k * 2048, 8192, 1024) : umma_desc_k_sw128(base + k * 32);
      };
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int slice = tile / tiles_mn;
        const int kb0 = slice * p.kb_per_slice;
        const int kb1 = min(p.num_k_blocks, kb0 + p.kb_per_slice);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t aH = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t bH = aH + nplanes * kABytes;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint64_t da = desc(aH, k);
            const uint64_t db = desc(bH, k);
            umma_f16(d_tmem, da, db, idesc, (kb != kb0) || (k != 0));
            if (p.nsplit == 3) {
              const uint64_t dal = desc(aH + kABytes, k);
              const uint64_t dbl = desc(bH + kBBytes, k);
              umma_f16(d_tmem, dal, db, idesc, 1);
              umma_f16(d_tmem, da, dbl, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tmem_full[acc]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int row_in_tile = ew * 32 + lane_id();
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int slice = tile / tiles_mn, mn = tile % tiles_mn;
      const int m_tile = mn / p.n_tiles, n_tile = mn % p.n_tiles;
      const long long out_row = (long long)m_tile * kBM + row_in_tile;
      const bool row_ok = out_row < p.M;
      float* slab = p.slabs + (long long)slice * p.M * p.N;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * BLOCK_N + ((uint32_t)(ew * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + c0, r);
        tmem_ld_wait();
        const int col0 = n_tile * BLOCK_N + c0;
        if (row_ok && col0 < p.N) {                        // N is a multiple of 32
          float4* op = reinterpret_cast<float4*>(slab + out_row * p.N + col0);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            op[j >> 2] = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                     __uint_as_float(r[j + 3]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

__global__ void splitk_reduce_kernel(const float* __restrict__ slabs, int slices, long long mn, int N, float scale,
                                     int accumulate, float* __restrict__ D, int ldd) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < mn; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < slices; ++k) s += slabs[(long long)k * mn + i];
    float* d = D + (i / N) * ldd + (i % N);
    *d = (accumulate ? *d : 0.f) + scale * s;
  }
}

struct SliceChoice { int block_n, m_tiles, n_tiles, nkb, kb_per, slices; };
SliceChoice choose_slices(int Mo, int No, int R, int nkb_override = 0) {
  SliceChoice c;
  c.block_n = (No % 256 == 0) ? 256 : (No % 128 == 0) ? 128 : 64;
  c.m_tiles = cdiv(Mo, kBM);
  c.n_tiles = cdiv(No, c.block_n);
  c.nkb = nkb_override ? nkb_override : cdiv(R, kBK);
  const int tiles = c.m_tiles * c.n_tiles;
  int s = (2 * sm_count() + tiles - 1) / tiles;
  if (s > 64) s = 64;
  if (s > c.nkb) s = c.nkb;
  if (s < 1) s = 1;
  c.kb_per = cdiv(c.nkb, s);
  c.slices = cdiv(c.nkb, c.kb_per);
  return c;
}

template <int BN, bool MN>
int launch_splitk(const CUtensorMap& tmA, const CUtensorMap& tmB, const SplitKParams& p, int grid, size_t smem,
                  cudaStream_t st) {
  static size_t attr = 0;
  if (smem > attr) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gemm_splitk_kernel<BN, MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  gemm_splitk_kernel<BN, MN><<<grid, kThreads, smem, st>>>(tmA, tmB, p);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// shared tail of the two entry points: slicing, launch, ordered slab reduction
template <bool MN>
int run_splitk(const CUtensorMap& tmA, const CUtensorMap& tmB, const SliceChoice& c, int Mo, int No, int nsplit, float scale,
               int accumulate, float* slabs, float* D, int ldd, cudaStream_t st, const SplitKParams* conv = nullptr) {
  const int np = nsplit == 3 ? 2 : 1;
  SplitKParams p = conv ? *conv : SplitKParams();
  if (!conv) p.conv = 0;
  p.M = Mo; p.N = No; p.num_k_blocks = c.nkb; p.kb_per_slice = c.kb_per; p.slices = c.slices; p.nsplit = nsplit;
  p.m_tiles = c.m_tiles; p.n_tiles = c.n_tiles; p.slabs = slabs;
  const size_t stage_bytes = (size_t)np * (kBM * kBK * 2 + c.block_n * kBK * 2);
  int stages = (int)((227 * 1024 - 2048) / stage_bytes);
  if (stages > kStagesMax) stages = kStagesMax;
  MAED_CHECK_ARG(stages >= 2, "gemm_wgrad_splitk: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = 1024 + stages * stage_bytes + 256;
  const int tiles = c.m_tiles * c.n_tiles * c.slices;
  const int grid = tiles < sm_count() ? tiles : sm_count();
  int rc;
  if (c.block_n == 256) rc = launch_splitk<256, MN>(tmA, tmB, p, grid, smem, st);
  else if (c.block_n == 128) rc = launch_splitk<128, MN>(tmA, tmB, p, grid, smem, st);
  else rc = launch_splitk<64, MN>(tmA, tmB, p, grid, smem, st);
  MAED_PROPAGATE(rc);
  const long long mn = (long long)Mo * No;
  splitk_reduce_kernel<<<bw::grid_for(mn, 256), 256, 0, st>>>(slabs, c.slices, mn, No, scale, accumulate, D, ldd);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace

// Upper bound over every R <= the given one: the slice count is not monotonic in R (slices = ceil(nkb / ceil(nkb / s))), and
// callers size one scratch buffer for a family of launches (e.g. BT*197 token rows and BT*196 patch rows), so the bound uses
// the cap `s` on the number of slices instead of the exact count for this R.
size_t splitk_slab_floats(int Mo, int No, int R) {
  const SliceChoice c = choose_slices(Mo, No, R);
  const int tiles = c.m_tiles * c.n_tiles;
  int s = (2 * sm_count() + tiles - 1) / tiles;
  if (s > 64) s = 64;                                   // (no clamp by the K-block count: the implicit-conv mode has more K
  if (s < c.slices) s = c.slices;                       //  blocks than R / 64 when its pixel patches are ragged)
  return (size_t)s * Mo * No;
}

int gemm_wgrad_splitk(const __half* A, long long a_plane, int lda, const __half* B, long long b_plane, int ldb, int Mo,
                      int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                      cudaStream_t st) {
  MAED_CHECK_ARG(A && B && slabs && D, "gemm_wgrad_splitk: null argument");
  MAED_CHECK_ARG(Mo >= 1 && No >= 32 && No % 32 == 0 && R >= 1, "gemm_wgrad_splitk: bad shape Mo=%d No=%d R=%d", Mo, No, R);
  MAED_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= R && ldb >= R, "gemm_wgrad_splitk: row strides must be multiples of 8 "
                 "and >= R (lda=%d ldb=%d R=%d)", lda, ldb, R);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "gemm_wgrad_splitk: nsplit must be 1 or 3");
  const int np = nsplit == 3 ? 2 : 1;
  const SliceChoice c = choose_slices(Mo, No, R);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)R, (uint64_t)Mo, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(np == 2 ? a_plane : (long long)Mo * lda) * 2};
    const uint32_t box[3] = {(uint32_t)kBK, (uint32_t)kBM, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, A, 3, dims, str, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)R, (uint64_t)No, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(np == 2 ? b_plane : (long long)No * ldb) * 2};
    const uint32_t box[3] = {(uint32_t)kBK, (uint32_t)c.block_n, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmB, B, 3, dims, str, box));
  }
  return run_splitk<false>(tmA, tmB, c, Mo, No, nsplit, scale, accumulate, slabs, D, ldd, st);
}

// Same contraction on the operands as the tape holds them: dY [R, Mo] (row stride ld_dy), X [R, No_x] (row stride ld_x), both
// fp16 hi/lo planes; D [Mo, No] with No >= No_x a multiple of 32 — columns No_x..No of D come out as zeros (TMA zero-fills the
// out-of-range part of a box; used for the stem's 152 -> 160 columns).
int gemm_wgrad_rows(const __half* dY, long long dy_plane, int ld_dy, const __half* X, long long x_plane, int ld_x, int No_x,
                    int Mo, int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                    cudaStream_t st) {
  MAED_CHECK_ARG(dY && X && slabs && D, "gemm_wgrad_rows: null argument");
  MAED_CHECK_ARG(Mo >= 1 && No >= 32 && No % 32 == 0 && R >= 1 && No_x >= 1 && No_x <= No,
                 "gemm_wgrad_rows: bad shape Mo=%d No=%d (%d) R=%d", Mo, No, No_x, R);
  MAED_CHECK_ARG(ld_dy % 8 == 0 && ld_x % 8 == 0 && ld_dy >= Mo && ld_x >= No_x, "gemm_wgrad_rows: row strides must be multiples "
                 "of 8 and cover the rows (ld_dy=%d Mo=%d ld_x=%d No=%d)", ld_dy, Mo, ld_x, No_x);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "gemm_wgrad_rows: nsplit must be 1 or 3");
  MAED_CHECK_ARG(ldd >= No, "gemm_wgrad_rows: ldd=%d < No=%d", ldd, No);
  const int np = nsplit == 3 ? 2 : 1;
  SliceChoice c = choose_slices(Mo, No, R);
  CUtensorMap tmA, tmB;
  {
    const uint64_t dims[3] = {(uint64_t)Mo, (uint64_t)R, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ld_dy * 2, (uint64_t)(np == 2 ? dy_plane : (long long)R * ld_dy) * 2};
    const uint32_t box[3] = {64, (uint32_t)kBK, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, dY, 3, dims, str, box));
  }
  {
    const uint64_t dims[3] = {(uint64_t)No_x, (uint64_t)R, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ld_x * 2, (uint64_t)(np == 2 ? x_plane : (long long)R * ld_x) * 2};
    const uint32_t box[3] = {64, (uint32_t)kBK, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmB, X, 3, dims, str, box));
  }
  return run_splitk<true>(tmA, tmB, c, Mo, No, nsplit, scale, accumulate, slabs, D, ldd, st);
}

// 64-pixel patch (tile_h x tile_w) covering an H x W map with the fewest patches
static void wgrad_patch_shape(int H, int W, int* th, int* tw) {
  int best = 1 << 30;
  *th = 8; *tw = 8;
  for (int h = 1; h <= 64; h *= 2) {
    const int w = 64 / h;
    const int n = cdiv(H, h) * cdiv(W, w);
    if (n < best) { best = n; *th = h; *tw = w; }
  }
}
int wgrad_conv_k_blocks(int n_img, int H, int W) {
  int th, tw;
  wgrad_patch_shape(H, W, &th, &tw);
  return n_img * cdiv(H, th) * cdiv(W, tw);
}

// Weight gradient of a stride-1 k x k convolution without an im2col matrix: D[Cout, k*k*Cin] (+)= scale * sum over the output
// pixels of dY[pixel, Cout]^T x[pixel shifted by the tap, Cin].  dY [n_img*H*W, Cout] and x NHWC [n_img, H, W, Cin] are fp16
// hi/lo planes; Cin and Cout multiples of 64; `pad` = zero padding at the top / left.
int gemm_wgrad_conv(const __half* dY, long long dy_plane, const __half* X, long long x_plane, int n_img, int H, int W, int Cin,
                    int Cout, int KH, int KW, int pad, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                    cudaStream_t st) {
  MAED_CHECK_ARG(dY && X && slabs && D, "gemm_wgrad_conv: null argument");
  MAED_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0 && KH >= 1 && KW >= 1 && n_img >= 1, "gemm_wgrad_conv: bad shape Cin=%d Cout=%d",
                 Cin, Cout);
  MAED_CHECK_ARG(nsplit == 3, "gemm_wgrad_conv: split precision only");
  const int No = KH * KW * Cin;
  MAED_CHECK_ARG(ldd >= No, "gemm_wgrad_conv: ldd=%d < %d", ldd, No);
  SplitKParams cp;
  cp.conv = 1; cp.Cin = Cin; cp.KW = KW; cp.pad = pad;
  wgrad_patch_shape(H, W, &cp.tile_h, &cp.tile_w);
  cp.tiles_w = cdiv(W, cp.tile_w);
  cp.patches = cdiv(H, cp.tile_h) * cp.tiles_w;
  const SliceChoice c = choose_slices(Cout, No, 0, n_img * cp.patches);
  CUtensorMap tmA, tmB;
  auto tmap5 = [&](CUtensorMap* tm, const __half* base, long long plane, int Cc) -> int {
    const uint64_t dims[5] = {(uint64_t)Cc, (uint64_t)W, (uint64_t)H, (uint64_t)n_img, 2};
    const uint64_t str[4] = {(uint64_t)Cc * 2, (uint64_t)W * Cc * 2, (uint64_t)H * W * Cc * 2, (uint64_t)plane * 2};
    const uint32_t box[5] = {64, (uint32_t)cp.tile_w, (uint32_t)cp.tile_h, 1, 1};
    return make_tmap_f16(tm, base, 5, dims, str, box);
  };
  MAED_PROPAGATE(tmap5(&tmA, dY, dy_plane, Cout));
  MAED_PROPAGATE(tmap5(&tmB, X, x_plane, Cin));
  return run_splitk<true>(tmA, tmB, c, Cout, No, nsplit, scale, accumulate, slabs, D, ldd, st, &cp);
}

}  // namespace maed
