// Host side of the tcgen05 GEMM: tensor-map construction, tile-shape choice, launch.
#include <stdlib.h>

#include "gemm_host.h"

#include <algorithm>
#include <mutex>
#include <stdarg.h>
#include <string.h>

#include "gemm_sm100.cuh"
#include "gemm2_sm100.cuh"

namespace maed {

// ------------------------------------------------------------------------------------------- errors
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// --------------------------------------------------------------------------------------- driver API
int get_encode_tiled(PFN_encodeTiled* fn) {
  static PFN_encodeTiled cached = nullptr;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (!cached) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      set_error("cuTensorMapEncodeTiled not available from the CUDA driver (%s)", cudaGetErrorString(e));
      return MAED_ERR_DRIVER;
    }
    cached = reinterpret_cast<PFN_encodeTiled>(p);
  }
  *fn = cached;
  return MAED_OK;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes) {
  PFN_encodeTiled enc;
  MAED_PROPAGATE(get_encode_tiled(&enc));
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u]", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return MAED_ERR_DRIVER;
  }
  return MAED_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------- tile choices
void conv_tile_shape(int H, int W, int* tile_h, int* tile_w) {
  double best = -1.0;
  int bh = 1, bw = 1;
  for (int tw = 1; tw <= W && tw <= 128; ++tw) {
    for (int th = 1; th * tw <= 128 && th <= H; ++th) {
      const double tiles = (double)cdiv(H, th) * cdiv(W, tw);
      const double util = (double)H * W / (tiles * 128.0);
      // prefer wide tiles (longer contiguous TMA rows) on ties
      if (util > best + 1e-9 || (util > best - 1e-9 && tw > bw)) { best = util; bh = th; bw = tw; }
    }
  }
  *tile_h = bh;
  *tile_w = bw;
}

static int choose_block_n(long long m_tiles, int N, int force) {
  if (force) return force;
  const int cands[3] = {256, 128, 64};
  int best_bn = 0;
  double best_eff = -1.0;
  const int sms = sm_count();
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (N % bn != 0) continue;
    const long long tiles = m_tiles * (N / bn);
    const double eff = (double)tiles / ((double)cdiv(tiles, sms) * sms);
    if (eff >= 0.85) return bn;            // largest tile that fills the machine well
    if (eff > best_eff) { best_eff = eff; best_bn = bn; }
  }
  return best_bn;
}

template <int BN, bool TMA_EPI>
static int launch_bn(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmO, GemmParams& p, cudaStream_t st) {
  const int nplanes = p.nsplit == 3 ? 2 : 1;
  const size_t stage_bytes = (size_t)nplanes * (kBlockM * kBlockK * 2 + BN * kBlockK * 2);
  const size_t epi_bytes = TMA_EPI ? kEpiStageBytes : 0;
  const size_t budget = 232448 - 1024 - 512 - epi_bytes;
  int stages = (int)(budget / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) { set_error("gemm: tile too large for shared memory"); return MAED_ERR_UNSUPPORTED; }
  p.stages = stages;
  const size_t smem = 1024 + epi_bytes + (size_t)stages * stage_bytes + 512;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, TMA_EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const long long tiles = (long long)p.m_tiles * p.n_tiles;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  gemm_tc_kernel<BN, TMA_EPI><<<grid, kGemmThreads, smem, st>>>(tmA, tmB, tmO, p);
  MAED_CUDA_CHECK(cudaGetLastError());
  return MAED_OK;
}

// CTA-pair kernel (gemm2_sm100.cuh): 256-row tiles, each CTA loads its 128 rows of A and half of the B tile
template <int BN>
static int launch_pair(const CUtensorMap& tmA, const GemmArgs& g, GemmParams p, cudaStream_t st) {
  const int nplanes = p.nsplit == 3 ? 2 : 1;
  CUtensorMap tmB;
  {
    const int ldb = g.ldb ? g.ldb : g.K;
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.N, (uint64_t)nplanes};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(nplanes == 2 ? g.b_plane : (long long)g.N * ldb) * 2};
    const uint32_t box[3] = {64, (uint32_t)(BN / 2), 1};
    MAED_PROPAGATE(make_tmap_f16(&tmB, g.B, 3, dims, str, box));
  }
  p.n_tiles = g.N / BN;                                     // p.m_tiles: 128-row tiles, set by the caller (plain or conv)
  const size_t stage_bytes = (size_t)nplanes * (kBlockM * kBlockK * 2 + (BN / 2) * kBlockK * 2);
  int stages = (int)((232448 - 1024 - 512) / stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 512;
  static bool attr_set = false;
  if (!attr_set) {
    MAED_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    attr_set = true;
  }
  const long long tiles = (long long)((p.m_tiles + 1) / 2) * p.n_tiles;
  const long long pairs = std::min<long long>(tiles, sm_count() / 2);
  gemm_tc2_kernel<BN><<<(int)(2 * pairs), kGemmThreads, smem, st>>>(tmA, tmB, p);
  MAED_CUDA_CHECK(cudaGetLastError());
  return MAED_OK;
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  MAED_CHECK_ARG(g.nsplit == 1 || g.nsplit == 3, "gemm: nsplit must be 1 or 3");
  MAED_CHECK_ARG(g.N % 32 == 0, "gemm: N=%d must be a multiple of 32", g.N);
  MAED_CHECK_ARG(g.M > 0 && g.K > 0, "gemm: empty problem M=%d K=%d", g.M, g.K);
  const int nplanes = g.nsplit == 3 ? 2 : 1;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M; p.N = g.N; p.K = g.K; p.nsplit = g.nsplit;
  p.bias = g.bias; p.residual = g.residual; p.act = g.act; p.out_mode = g.out_mode; p.out = g.out;
  p.res_hi = g.res_hi; p.res_plane = g.res_plane; p.act_post = g.act_post;
  p.out_plane_stride = g.out_plane; p.ldc = g.ldc ? g.ldc : g.N;
  MAED_CHECK_ARG(!g.res_hi || ((reinterpret_cast<uintptr_t>(g.res_hi) & 15) == 0 && g.res_plane % 8 == 0 && p.ldc % 8 == 0),
                 "gemm: plane residual needs a 16-byte aligned base, plane stride and row stride");

  CUtensorMap tmA, tmB;
  if (g.conv) {
    MAED_CHECK_ARG(g.Cin % 64 == 0, "gemm(conv): Cin=%d must be a multiple of 64", g.Cin);
    MAED_CHECK_ARG(g.K == g.KH * g.KW * g.Cin, "gemm(conv): K mismatch");
    MAED_CHECK_ARG(g.M == g.n_img * g.H * g.W, "gemm(conv): M mismatch");
    p.conv = 1; p.H = g.H; p.W = g.W; p.cin_blocks = g.Cin / 64; p.KW = g.KW; p.pad_h = g.pad_h; p.pad_w = g.pad_w;
    conv_tile_shape(g.H, g.W, &p.tile_h, &p.tile_w);
    p.tiles_h = cdiv(g.H, p.tile_h); p.tiles_w = cdiv(g.W, p.tile_w);
    p.m_tiles = g.n_img * p.tiles_h * p.tiles_w;
    p.num_k_blocks = g.KH * g.KW * p.cin_blocks;
    p.a_tx_bytes = (uint32_t)(p.tile_h * p.tile_w * 128);
    const uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.n_img, (uint64_t)nplanes};
    const uint64_t str[4] = {(uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.H * g.W * g.Cin * 2,
                             (uint64_t)(nplanes == 2 ? g.a_plane : (long long)g.M * g.Cin) * 2};
    const uint32_t box[5] = {64, (uint32_t)p.tile_w, (uint32_t)p.tile_h, 1, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, g.A, 5, dims, str, box));
  } else {
    const int lda = g.lda ? g.lda : g.K;
    MAED_CHECK_ARG(lda % 8 == 0, "gemm: lda=%d must be a multiple of 8 (16-byte TMA strides)", lda);
    p.m_tiles = cdiv(g.M, kBlockM);
    p.num_k_blocks = cdiv(g.K, kBlockK);
    p.a_tx_bytes = kBlockM * kBlockK * 2;
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.M, (uint64_t)nplanes};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(nplanes == 2 ? g.a_plane : (long long)g.M * lda) * 2};
    const uint32_t box[3] = {64, 128, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmA, g.A, 3, dims, str, box));
  }
  // opt-in CTA-pair path (MAED_B200_GEMM_2CTA=1): plain GEMMs with at least one full pair tile
  static const bool pair_on = getenv("MAED_B200_GEMM_2CTA") != nullptr;
  if (pair_on && p.m_tiles >= 2 && (g.N % 128) == 0 && (g.force_block_n == 0 || g.force_block_n >= 128)) {
    const int bn2 = g.force_block_n ? g.force_block_n : ((g.N % 256) == 0 ? 256 : 128);
    return bn2 == 256 ? launch_pair<256>(tmA, g, p, st) : launch_pair<128>(tmA, g, p, st);
  }
  const int bn = choose_block_n(p.m_tiles, g.N, g.force_block_n);
  MAED_CHECK_ARG(bn == 64 || bn == 128 || bn == 256, "gemm: no tile width for N=%d", g.N);
  MAED_CHECK_ARG(g.N % bn == 0, "gemm: N=%d not a multiple of BLOCK_N=%d", g.N, bn);
  p.n_tiles = g.N / bn;
  {
    const int ldb = g.ldb ? g.ldb : g.K;
    MAED_CHECK_ARG(ldb % 8 == 0, "gemm: ldb=%d must be a multiple of 8", ldb);
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.N, (uint64_t)nplanes};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(nplanes == 2 ? g.b_plane : (long long)g.N * ldb) * 2};
    const uint32_t box[3] = {64, (uint32_t)bn, 1};
    MAED_PROPAGATE(make_tmap_f16(&tmB, g.B, 3, dims, str, box));
  }
  // TMA-store epilogue (default since round 2: fc1 396 -> 430 TFLOP/s algorithmic on a B200; MAED_B200_GEMM_TMA_EPI=0 selects
  // the direct-store epilogue): plain mode, 16-byte aligned output rows
  static const bool tma_epi_on = [] { const char* v = getenv("MAED_B200_GEMM_TMA_EPI"); return !(v && v[0] == '0'); }();
  const bool split_out = g.out_mode == OUT_F16_SPLIT;
  const bool tma_epi = tma_epi_on && !g.conv && g.out != nullptr && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0 &&
                       (g.out_mode == OUT_F32 ? (p.ldc % 4 == 0) : (p.ldc % 8 == 0)) && (!split_out || g.out_plane % 8 == 0);
  CUtensorMap tmO = tmB;
  if (tma_epi) {
    if (g.out_mode == OUT_F32) {      // fp32 [M, ldc] described as 16-bit elements: 2N columns, 64-wide (= 32 floats) boxes
      const uint64_t dims[3] = {(uint64_t)g.N * 2, (uint64_t)g.M, 1};
      const uint64_t str[2] = {(uint64_t)p.ldc * 4, (uint64_t)g.M * p.ldc * 4};
      const uint32_t box[3] = {64, 128, 1};
      MAED_PROPAGATE(make_tmap_f16(&tmO, g.out, 3, dims, str, box, 128));
    } else {
      const uint64_t dims[3] = {(uint64_t)g.N, (uint64_t)g.M, (uint64_t)(split_out ? 2 : 1)};
      const uint64_t str[2] = {(uint64_t)p.ldc * 2, (uint64_t)(split_out ? g.out_plane : (long long)g.M * p.ldc) * 2};
      const uint32_t box[3] = {32, 128, 1};
      MAED_PROPAGATE(make_tmap_f16(&tmO, g.out, 3, dims, str, box, 64));
    }
    if (bn == 256) return launch_bn<256, true>(tmA, tmB, tmO, p, st);
    if (bn == 128) return launch_bn<128, true>(tmA, tmB, tmO, p, st);
    return launch_bn<64, true>(tmA, tmB, tmO, p, st);
  }
  if (bn == 256) return launch_bn<256, false>(tmA, tmB, tmO, p, st);
  if (bn == 128) return launch_bn<128, false>(tmA, tmB, tmO, p, st);
  return launch_bn<64, false>(tmA, tmB, tmO, p, st);
}

}  // namespace maed
