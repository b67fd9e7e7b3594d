// CPU restatements of the CONTRACTS of the tcgen05 / TMA kernels (gemm_sm100.cuh, gemm_gn_sm100.cu, stem_sm100.cu,
// attention.cu, gemm_splitk_sm100.cu) for the CUDA-on-CPU test build — TEST INFRASTRUCTURE ONLY (tests/emu/README.md).
//
// The tensor-core kernels themselves cannot run on a CPU; what can be checked without a GPU is everything around them:
// the host orchestration of engine.cu / train.cu (buffer plumbing, operand layouts, transposed / flipped weight packs,
// strides, the order of the backward walk) and the CUDA-core kernels, which are compiled from the real sources.  Each stub
// below (a) repeats the argument checks of the real launcher, including the constraints cuTensorMapEncodeTiled would
// enforce on the tensor maps it builds (16-byte base and strides, box <= 256), and (b) computes what the kernel is
// specified to compute, in fp32 on the hi + lo operand planes.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "bwd_kernels.h"
#include "emu_gemm_enums.h"      // OUT_* / ACT_* extracted from gemm_sm100.cuh by build_emu.py
#include "gemm_host.h"
#include "kernels.h"

namespace emu { extern double g_prof[8]; }
namespace {
struct Prof {
  double t0; int slot;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
  explicit Prof(int s) : t0(now()), slot(s) {}
  ~Prof() { emu::g_prof[slot] += now() - t0; }
};
}  // namespace

namespace maed {

// ------------------------------------------------------------------------------------------ common host helpers
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
int sm_count() { return 148; }

static int workers() {
  static int n = 0;
  if (!n) {
    const char* v = getenv("MAED_EMU_THREADS");
    n = v ? atoi(v) : (int)std::thread::hardware_concurrency();
    n = std::max(1, std::min(n, 64));
  }
  return n;
}
template <class F>
static void parallel_for(long long n, F f) {
  const int nw = (int)std::min<long long>(workers(), n);
  if (nw <= 1) { for (long long i = 0; i < n; ++i) f(i); return; }
  std::atomic<long long> next{0};
  auto body = [&]() { for (;;) { const long long i = next.fetch_add(1); if (i >= n) break; f(i); } };
  std::vector<std::thread> pool;
  for (int i = 1; i < nw; ++i) pool.emplace_back(body);
  body();
  for (auto& t : pool) t.join();
}

// what cuTensorMapEncodeTiled checks for the maps the real launchers build (dims / box innermost first, fp16 elements)
static int check_tmap(const char* what, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box) {
  MAED_CHECK_ARG(base && ((uintptr_t)base & 15) == 0, "%s: tensor-map base %p is not 16-byte aligned", what, base);
  for (int i = 0; i < rank; ++i) {
    MAED_CHECK_ARG(dims[i] >= 1 && dims[i] <= (1ull << 32), "%s: tensor-map dim %d = %llu", what, i, (unsigned long long)dims[i]);
    MAED_CHECK_ARG(box[i] >= 1 && box[i] <= 256, "%s: tensor-map box %d = %u (1..256)", what, i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i)
    MAED_CHECK_ARG(strides_bytes[i] % 16 == 0 && strides_bytes[i] < (1ull << 40), "%s: tensor-map stride %d = %llu bytes "
                   "(multiple of 16 required)", what, i, (unsigned long long)strides_bytes[i]);
  MAED_CHECK_ARG(box[0] * 2 <= 128, "%s: inner box %u elements exceeds the 128-byte swizzle span", what, box[0]);
  return MAED_OK;
}

static inline float ld(const __half* p, long long plane) {
  return __half2float(p[0]) + (plane ? __half2float(p[plane]) : 0.f);
}
static inline void st_split(__half* p, long long plane, float v, bool lo) {
  const __half h = __float2half_rn(v);
  p[0] = h;
  if (lo) p[plane] = __float2half_rn(v - __half2float(h));
}

// C[M,N] = A[M,K] * B[N,K]^T, fp32, row-major dense.  Register-tiled 6 x 32 micro-kernel on GCC vector types over a
// K-major packed panel of B (no reassociation inside a dot product beyond the fixed k order).
typedef float vf16 __attribute__((vector_size(64), aligned(4)));
static void sgemm_nt(long long M, int N, int K, const float* A, const float* B, float* C) {
  constexpr int NR = 32, MR = 6;
  const int npanels = (N + NR - 1) / NR;
  // Bp[panel][k][NR]
  std::vector<float> Bp((size_t)npanels * K * NR, 0.f);
  parallel_for(npanels, [&](long long pb) {
    float* dst = &Bp[(size_t)pb * K * NR];
    const int n0 = (int)pb * NR, nn = std::min(NR, N - n0);
    for (int n = 0; n < nn; ++n) {
      const float* src = B + (size_t)(n0 + n) * K;
      for (int k = 0; k < K; ++k) dst[(size_t)k * NR + n] = src[k];
    }
  });
  const long long mblocks = (M + MR - 1) / MR;
  const long long chunk = 8;                                   // row blocks per work item
  parallel_for((mblocks + chunk - 1) / chunk, [&](long long wi) {
    for (long long mb = wi * chunk; mb < std::min(mblocks, (wi + 1) * chunk); ++mb) {
      const long long m0 = mb * MR;
      const int mm = (int)std::min<long long>(MR, M - m0);
      const float* a[MR];
      for (int i = 0; i < MR; ++i) a[i] = A + (size_t)(m0 + (i < mm ? i : 0)) * K;
      for (int pb = 0; pb < npanels; ++pb) {
        const float* bp = &Bp[(size_t)pb * K * NR];
        vf16 acc[MR][2];
        for (int i = 0; i < MR; ++i) { acc[i][0] = vf16{} ; acc[i][1] = vf16{}; }
        for (int k = 0; k < K; ++k) {
          const vf16 b0 = *reinterpret_cast<const vf16*>(bp + (size_t)k * NR);
          const vf16 b1 = *reinterpret_cast<const vf16*>(bp + (size_t)k * NR + 16);
#pragma GCC unroll 6
          for (int i = 0; i < MR; ++i) {
            const float av = a[i][k];
            acc[i][0] += av * b0;
            acc[i][1] += av * b1;
          }
        }
        const int n0 = pb * NR, nn = std::min(NR, N - n0);
        for (int i = 0; i < mm; ++i) {
          float tmp[NR];
          memcpy(tmp, &acc[i][0], 64);
          memcpy(tmp + 16, &acc[i][1], 64);
          memcpy(C + (size_t)(m0 + i) * N + n0, tmp, (size_t)nn * 4);
        }
      }
    }
  });
}

// planes [rows, ld] (cols used: K) -> dense fp32 [rows, K]
static void planes_to_dense(const __half* p, long long plane, long long rows, int ldp, int K, float* out) {
  parallel_for((rows + 255) / 256, [&](long long rb) {
    const long long r1 = std::min(rows, (rb + 1) * 256);
    for (long long r = rb * 256; r < r1; ++r)
      for (int k = 0; k < K; ++k) out[(size_t)r * K + k] = ld(p + r * ldp + k, plane);
  });
}

// ------------------------------------------------------------------------------------------------- launch_gemm
int launch_gemm(const GemmArgs& g, cudaStream_t) {
  Prof prof(1);
  MAED_CHECK_ARG(g.nsplit == 1 || g.nsplit == 3, "gemm: nsplit must be 1 or 3");
  MAED_CHECK_ARG(g.N % 32 == 0, "gemm: N=%d must be a multiple of 32", g.N);
  MAED_CHECK_ARG(g.M > 0 && g.K > 0, "gemm: empty problem M=%d K=%d", g.M, g.K);
  MAED_CHECK_ARG(g.A && g.B && g.out, "gemm: null operand");
  const int np = g.nsplit == 3 ? 2 : 1;
  const long long a_plane = np == 2 ? g.a_plane : 0, b_plane = np == 2 ? g.b_plane : 0;
  const int ldc = g.ldc ? g.ldc : g.N;
  std::vector<float> A((size_t)g.M * g.K), B((size_t)g.N * g.K), C((size_t)g.M * g.N);
  if (g.conv) {
    MAED_CHECK_ARG(g.Cin % 64 == 0, "gemm(conv): Cin=%d must be a multiple of 64", g.Cin);
    MAED_CHECK_ARG(g.K == g.KH * g.KW * g.Cin, "gemm(conv): K mismatch");
    MAED_CHECK_ARG(g.M == g.n_img * g.H * g.W, "gemm(conv): M mismatch");
    int th, tw;
    conv_tile_shape(g.H, g.W, &th, &tw);
    const uint64_t dims[5] = {(uint64_t)g.Cin, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.n_img, (uint64_t)np};
    const uint64_t str[4] = {(uint64_t)g.Cin * 2, (uint64_t)g.W * g.Cin * 2, (uint64_t)g.H * g.W * g.Cin * 2,
                             (uint64_t)(np == 2 ? g.a_plane : (long long)g.M * g.Cin) * 2};
    const uint32_t box[5] = {64, (uint32_t)tw, (uint32_t)th, 1, 1};
    MAED_PROPAGATE(check_tmap("gemm(conv) A", g.A, 5, dims, str, box));
    MAED_CHECK_ARG(np == 1 || g.a_plane >= (long long)g.M * g.Cin, "gemm(conv): A planes overlap");
    // im2col with zero padding (TMA out-of-bounds fill): k = (kh*KW + kw)*Cin + c
    parallel_for((long long)g.n_img * g.H, [&](long long nh) {
      const int n = (int)(nh / g.H), h = (int)(nh % g.H);
      for (int w = 0; w < g.W; ++w) {
        float* row = &A[((size_t)nh * g.W + w) * g.K];
        for (int kh = 0; kh < g.KH; ++kh)
          for (int kw = 0; kw < g.KW; ++kw) {
            const int ih = h + kh - g.pad_h, iw = w + kw - g.pad_w;
            float* dst = row + (size_t)(kh * g.KW + kw) * g.Cin;
            if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) { for (int c = 0; c < g.Cin; ++c) dst[c] = 0.f; continue; }
            const __half* src = g.A + (((long long)n * g.H + ih) * g.W + iw) * g.Cin;
            for (int c = 0; c < g.Cin; ++c) dst[c] = ld(src + c, a_plane);
          }
      }
    });
  } else {
    const int lda = g.lda ? g.lda : g.K;
    MAED_CHECK_ARG(lda % 8 == 0, "gemm: lda=%d must be a multiple of 8 (16-byte TMA strides)", lda);
    MAED_CHECK_ARG(lda >= g.K, "gemm: lda=%d < K=%d", lda, g.K);
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.M, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(np == 2 ? g.a_plane : (long long)g.M * lda) * 2};
    const uint32_t box[3] = {64, 128, 1};
    MAED_PROPAGATE(check_tmap("gemm A", g.A, 3, dims, str, box));
    planes_to_dense(g.A, a_plane, g.M, lda, g.K, A.data());
  }
  {
    MAED_CHECK_ARG(g.N % 64 == 0, "gemm: no tile width for N=%d", g.N);
    const int ldb = g.ldb ? g.ldb : g.K;
    MAED_CHECK_ARG(ldb % 8 == 0, "gemm: ldb=%d must be a multiple of 8", ldb);
    MAED_CHECK_ARG(ldb >= g.K, "gemm: ldb=%d < K=%d", ldb, g.K);
    const uint64_t dims[3] = {(uint64_t)g.K, (uint64_t)g.N, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(np == 2 ? g.b_plane : (long long)g.N * ldb) * 2};
    const uint32_t box[3] = {64, 64, 1};
    MAED_PROPAGATE(check_tmap("gemm B", g.B, 3, dims, str, box));
    planes_to_dense(g.B, b_plane, g.N, ldb, g.K, B.data());
  }
  // epilogue stores are 16-byte vectors
  MAED_CHECK_ARG(((uintptr_t)g.out & 15) == 0 && (g.out_mode == OUT_F32 ? ldc % 4 == 0 : ldc % 8 == 0),
                 "gemm: output base / row stride not 16-byte aligned");
  MAED_CHECK_ARG(!g.residual || (((uintptr_t)g.residual & 15) == 0), "gemm: residual not 16-byte aligned");
  MAED_CHECK_ARG(!g.bias || (((uintptr_t)g.bias & 15) == 0), "gemm: bias not 16-byte aligned");
  MAED_CHECK_ARG(!g.res_hi || ((((uintptr_t)g.res_hi) & 15) == 0 && g.res_plane % 8 == 0 && ldc % 8 == 0),
                 "gemm: plane residual needs a 16-byte aligned base, plane stride and row stride");
  MAED_CHECK_ARG(g.out_mode != OUT_F16_SPLIT || g.out_plane % 8 == 0, "gemm: out_plane not 16-byte aligned");
  sgemm_nt(g.M, g.N, g.K, A.data(), B.data(), C.data());
  parallel_for((g.M + 63) / 64, [&](long long rb) {
    const long long r1 = std::min<long long>(g.M, (rb + 1) * 64);
    for (long long r = rb * 64; r < r1; ++r)
      for (int n = 0; n < g.N; ++n) {
        float v = C[(size_t)r * g.N + n];
        if (g.bias) v += g.bias[n];
        if (g.act == ACT_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
        else if (g.act == ACT_RELU) v = fmaxf(v, 0.f);
        else if (g.act == ACT_TANH) v = tanhf(v);
        if (g.residual) v += g.residual[r * ldc + n];
        if (g.res_hi) v += ld(g.res_hi + r * ldc + n, g.res_plane);
        if (g.act_post == ACT_RELU) v = fmaxf(v, 0.f);
        if (g.out_mode == OUT_F32) static_cast<float*>(g.out)[r * ldc + n] = v;
        else st_split(static_cast<__half*>(g.out) + r * ldc + n, g.out_plane, v, g.out_mode == OUT_F16_SPLIT);
      }
  });
  count_launch();
  return MAED_OK;
}

void conv_tile_shape(int H, int W, int* tile_h, int* tile_w) {
  double best = -1.0;
  int bh = 1, bw = 1;
  for (int tw = 1; tw <= W && tw <= 128; ++tw)
    for (int th = 1; th * tw <= 128 && th <= H; ++th) {
      const double tiles = (double)cdiv(H, th) * cdiv(W, tw);
      const double util = (double)H * W / (tiles * 128.0);
      if (util > best + 1e-9 || (util > best - 1e-9 && tw > bw)) { best = util; bh = th; bw = tw; }
    }
  *tile_h = bh;
  *tile_w = bw;
}

// fused conv + GroupNorm: the emulated engine always takes the unfused path (launch_gemm + gn_stats + gn_apply), which
// engine.cu selects when this returns MAED_ERR_UNSUPPORTED
int conv_gn_fused(const ConvGnArgs&, cudaStream_t) { return MAED_ERR_UNSUPPORTED; }

// --------------------------------------------------------------------------------------------------- stem conv
int stem_conv(const float* x, int n_img, const __half* w_hi, long long w_plane, int k_pad, int nsplit, float* out,
              double* stats, cudaStream_t) {
  Prof prof(2);
  MAED_CHECK_ARG(k_pad % 8 == 0 && k_pad >= 147, "stem_conv: bad k_pad %d", k_pad);
  const int np = nsplit == 3 ? 2 : 1;
  const uint64_t dims[3] = {(uint64_t)k_pad, 64, (uint64_t)np};
  const uint64_t str[2] = {(uint64_t)k_pad * 2, (uint64_t)(np == 2 ? w_plane : 64LL * k_pad) * 2};
  const uint32_t box[3] = {64, 64, 1};
  MAED_PROPAGATE(check_tmap("stem_conv W", w_hi, 3, dims, str, box));
  std::vector<float> W(64 * 147);
  for (int co = 0; co < 64; ++co)
    for (int k = 0; k < 147; ++k) W[co * 147 + k] = ld(w_hi + (long long)co * k_pad + k, np == 2 ? w_plane : 0);
  // 7x7 stride 2, TF-SAME: pad_total = 5 -> 2 top / left (reference resnetv2.py:51-59)
  parallel_for((long long)n_img * 112, [&](long long noh) {
    const int n = (int)(noh / 112), oh = (int)(noh % 112);
    for (int ow = 0; ow < 112; ++ow) {
      float patch[147];
      for (int r = 0; r < 7; ++r)
        for (int s = 0; s < 7; ++s) {
          const int ih = oh * 2 + r - 2, iw = ow * 2 + s - 2;
          for (int c = 0; c < 3; ++c)
            patch[(r * 7 + s) * 3 + c] = (ih < 0 || ih >= 224 || iw < 0 || iw >= 224)
                                             ? 0.f : x[(((long long)n * 3 + c) * 224 + ih) * 224 + iw];
        }
      float* o = out + (((long long)n * 112 + oh) * 112 + ow) * 64;
      for (int co = 0; co < 64; ++co) {
        float acc = 0.f;
        for (int k = 0; k < 147; ++k) acc += patch[k] * W[co * 147 + k];
        o[co] = acc;
      }
    }
  });
  for (int n = 0; n < n_img; ++n)
    for (int g = 0; g < 32; ++g) {
      double s = 0, q = 0;
      for (long long p = 0; p < 112 * 112; ++p)
        for (int c = 2 * g; c < 2 * g + 2; ++c) {
          const double v = out[((long long)n * 12544 + p) * 64 + c];
          s += v; q += v * v;
        }
      stats[((long long)n * 32 + g) * 2] += s;
      stats[((long long)n * 32 + g) * 2 + 1] += q;
    }
  count_launch();
  return MAED_OK;
}

// --------------------------------------------------------------------------------------------------- attention
// softmax(q k^T * scale) v over the rows  row(i) of each (group, head); rows given by a callback
template <class RowFn>
static void attention_ref(const __half* qkv, long long plane, long long groups, int seq, int heads, float scale, RowFn row_of,
                          float* out_f32, __half* out_hi, long long out_plane, float* lse = nullptr) {
  Prof prof(3);
  const int ldq = 3 * heads * 64, C = heads * 64;
  parallel_for(groups * heads, [&](long long gh) {
    const long long g = gh / heads;
    const int h = (int)(gh % heads);
    std::vector<float> q((size_t)seq * 64), k((size_t)seq * 64), v((size_t)seq * 64), s(seq);
    for (int i = 0; i < seq; ++i) {
      const __half* r = qkv + row_of(g, i) * ldq + h * 64;
      for (int d = 0; d < 64; ++d) {
        q[(size_t)i * 64 + d] = ld(r + d, plane);
        k[(size_t)i * 64 + d] = ld(r + C + d, plane);
        v[(size_t)i * 64 + d] = ld(r + 2 * C + d, plane);
      }
    }
    for (int i = 0; i < seq; ++i) {
      float mx = -INFINITY;
      for (int j = 0; j < seq; ++j) {
        float a = 0.f;
        for (int d = 0; d < 64; ++d) a += q[(size_t)i * 64 + d] * k[(size_t)j * 64 + d];
        s[j] = a * scale;
        mx = fmaxf(mx, s[j]);
      }
      float sum = 0.f;
      for (int j = 0; j < seq; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
      if (lse) lse[row_of(g, i) * heads + h] = (mx + logf(sum)) * 1.4426950408889634f;     // log2 domain, like the kernels
      float o[64];
      for (int d = 0; d < 64; ++d) o[d] = 0.f;
      for (int j = 0; j < seq; ++j) {
        const float p = s[j] / sum;
        for (int d = 0; d < 64; ++d) o[d] += p * v[(size_t)j * 64 + d];
      }
      const long long off = row_of(g, i) * C + h * 64;
      for (int d = 0; d < 64; ++d) {
        if (out_f32) out_f32[off + d] = o[d];
        if (out_hi) st_split(out_hi + off + d, out_plane, o[d], true);
      }
    }
  });
  count_launch();
}

int attn_spatial(const __half* qkv_hi, long long qkv_plane, int BT, int ntok, int heads, float scale, int nsplit,
                 float* out_f32, __half* out_hi, long long out_plane, cudaStream_t, float* lse) {
  MAED_CHECK_ARG(ntok >= 1 && ntok <= 208, "attn_spatial: ntok=%d unsupported (1..208)", ntok);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "attn_spatial: nsplit must be 1 or 3");
  const int np = nsplit == 3 ? 2 : 1;
  const int ldq = 3 * heads * 64;
  const long long rows = (long long)BT * ntok;
  const uint64_t dims[3] = {(uint64_t)ldq, (uint64_t)rows, (uint64_t)np};
  const uint64_t str[2] = {(uint64_t)ldq * 2, (uint64_t)(np == 2 ? qkv_plane : rows * ldq) * 2};
  const uint32_t boxkv[3] = {64, 208, 1};
  MAED_PROPAGATE(check_tmap("attn_spatial qkv", qkv_hi, 3, dims, str, boxkv));
  if (out_hi)
    MAED_CHECK_ARG(out_plane >= rows * heads * 64 && (out_plane % 8) == 0 && ((uintptr_t)out_hi & 15) == 0,
                   "attn_spatial: output planes must be 16-byte aligned and at least rows*heads*64 apart");
  MAED_CHECK_ARG(out_f32 || out_hi, "attn_spatial: no output");
  attention_ref(qkv_hi, np == 2 ? qkv_plane : 0, BT, ntok, heads, scale,
                [=](long long g, int i) { return g * ntok + i; }, out_f32, out_hi, out_plane, lse);
  return MAED_OK;
}

int attn_temporal(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale, float* out_f32,
                  __half* out_hi, long long out_plane, cudaStream_t, float* lse) {
  MAED_CHECK_ARG(T >= 1 && T <= 32, "attn_temporal: T=%d unsupported (1..32)", T);
  MAED_CHECK_ARG(qkv_plane % 8 == 0 && out_plane % 4 == 0, "attn_temporal: plane strides must be 16-byte aligned");
  // group = (clip b, token n); row(i) = (b*T + i)*ntok + n
  attention_ref(qkv_hi, qkv_plane, (long long)B * ntok, T, heads, scale,
                [=](long long g, int i) { return ((g / ntok) * T + i) * ntok + (g % ntok); }, out_f32, out_hi, out_plane, lse);
  return MAED_OK;
}
int attn_temporal_tc(const __half* qkv_hi, long long qkv_plane, int B, int T, int ntok, int heads, float scale, float* out_f32,
                     __half* out_hi, long long out_plane, cudaStream_t st, float* lse) {
  MAED_CHECK_ARG((T == 4 || T == 8 || T == 16 || T == 32) && qkv_plane != 0, "attn_temporal_tc: T=%d unsupported", T);
  return attn_temporal(qkv_hi, qkv_plane, B, T, ntok, heads, scale, out_f32, out_hi, out_plane, st, lse);
}

int attn_generic(const __half* qkv_hi, long long qkv_plane, int batch, int seq, int heads, float scale, int, int,
                 float* out_f32, __half* out_hi, long long out_plane, cudaStream_t) {
  attention_ref(qkv_hi, qkv_plane, batch, seq, heads, scale, [=](long long g, int i) { return g * seq + i; }, out_f32, out_hi,
                out_plane);
  return MAED_OK;
}

// ------------------------------------------------------------------------------------------- split-K weight gradient
size_t splitk_slab_floats(int Mo, int No, int) { return (size_t)Mo * No; }

int gemm_wgrad_splitk(const __half* A, long long a_plane, int lda, const __half* B, long long b_plane, int ldb, int Mo,
                      int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, cudaStream_t) {
  Prof prof(4);
  MAED_CHECK_ARG(A && B && slabs && D, "gemm_wgrad_splitk: null argument");
  MAED_CHECK_ARG(Mo >= 1 && No >= 32 && No % 32 == 0 && R >= 1, "gemm_wgrad_splitk: bad shape Mo=%d No=%d R=%d", Mo, No, R);
  MAED_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && lda >= R && ldb >= R, "gemm_wgrad_splitk: row strides must be multiples of 8 "
                 "and >= R (lda=%d ldb=%d R=%d)", lda, ldb, R);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "gemm_wgrad_splitk: nsplit must be 1 or 3");
  MAED_CHECK_ARG(ldd >= No, "gemm_wgrad_splitk: ldd=%d < No=%d", ldd, No);
  const int np = nsplit == 3 ? 2 : 1;
  {
    const uint64_t dims[3] = {(uint64_t)R, (uint64_t)Mo, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)lda * 2, (uint64_t)(np == 2 ? a_plane : (long long)Mo * lda) * 2};
    const uint32_t box[3] = {64, 128, 1};
    MAED_PROPAGATE(check_tmap("wgrad A", A, 3, dims, str, box));
    MAED_CHECK_ARG(np == 1 || a_plane >= (long long)(Mo - 1) * lda + R, "wgrad: A planes overlap");
  }
  {
    const uint64_t dims[3] = {(uint64_t)R, (uint64_t)No, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ldb * 2, (uint64_t)(np == 2 ? b_plane : (long long)No * ldb) * 2};
    const uint32_t box[3] = {64, 64, 1};
    MAED_PROPAGATE(check_tmap("wgrad B", B, 3, dims, str, box));
    MAED_CHECK_ARG(np == 1 || b_plane >= (long long)(No - 1) * ldb + R, "wgrad: B planes overlap");
  }
  std::vector<float> Af((size_t)Mo * R), Bf((size_t)No * R), C((size_t)Mo * No);
  planes_to_dense(A, np == 2 ? a_plane : 0, Mo, lda, R, Af.data());
  planes_to_dense(B, np == 2 ? b_plane : 0, No, ldb, R, Bf.data());
  sgemm_nt(Mo, No, R, Af.data(), Bf.data(), C.data());
  for (long long m = 0; m < Mo; ++m)
    for (int n = 0; n < No; ++n) {
      const float v = scale * C[(size_t)m * No + n];
      D[m * ldd + n] = accumulate ? D[m * ldd + n] + v : v;
    }
  slabs[0] = 0.f;                              // the real kernel scribbles over the slab scratch
  count_launch(2);
  return MAED_OK;
}

// ------------------------------------------------------------------------- spatial attention backward (tcgen05 kernel)
// contract: the CUDA-core kernel of attention_bwd.cu (compiled from the real source) on d_out = hi + lo
// (lse / Dv: the real kernel's shortcut around two reduction passes; the contract — the gradients — does not depend on them, but
//  they must be what the forward / attn_rowdot produce: checked here against a recomputation on a few rows)
int attn_spatial_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int BT, int ntok,
                        int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse, const float* Dv) {
  MAED_CHECK_ARG((lse == nullptr) == (Dv == nullptr), "attn_spatial_bwd_tc: lse and D come together");
  if (lse) {
    std::vector<float> chk((size_t)BT * ntok * heads), o((size_t)BT * ntok * heads * 64);
    attention_ref(qkv_hi, qkv_plane, BT, ntok, heads, scale, [=](long long g, int i) { return g * ntok + i; }, o.data(), nullptr, 0,
                  chk.data());
    for (long long i = 0; i < (long long)BT * ntok * heads; i += 37)
      MAED_CHECK_ARG(fabsf(chk[i] - lse[i]) <= 1e-3f * (1.f + fabsf(chk[i])), "attn_spatial_bwd_tc: lse[%lld] = %f, expected %f", i,
                     (double)lse[i], (double)chk[i]);
  }
  MAED_CHECK_ARG(qkv_hi && dout_hi && d_qkv, "attn_spatial_bwd_tc: null argument");
  MAED_CHECK_ARG(ntok >= 1 && ntok <= 208, "attn_spatial_bwd_tc: ntok=%d unsupported (1..208)", ntok);
  const long long rows = (long long)BT * ntok;
  const int ld3 = 3 * heads * 64, ldo = heads * 64;
  MAED_CHECK_ARG(qkv_plane >= rows * ld3 && dout_plane >= rows * ldo && qkv_plane % 8 == 0 && dout_plane % 8 == 0,
                 "attn_spatial_bwd_tc: operand planes overlap or are misaligned");
  {
    const uint64_t dims[3] = {(uint64_t)ldo, (uint64_t)rows, 2};
    const uint64_t str[2] = {(uint64_t)ldo * 2, (uint64_t)dout_plane * 2};
    const uint32_t box[3] = {64, 208, 1};
    MAED_PROPAGATE(check_tmap("attn bwd dO", dout_hi, 3, dims, str, box));
  }
  std::vector<float> d((size_t)rows * ldo);
  planes_to_dense(dout_hi, dout_plane, rows, ldo, ldo, d.data());
  return attn_spatial_bwd(qkv_hi, qkv_plane, d.data(), BT, ntok, heads, scale, accumulate, d_qkv, st);
}

int attn_temporal_bwd_tc(const __half* qkv_hi, long long qkv_plane, const __half* dout_hi, long long dout_plane, int B, int T,
                         int ntok, int heads, float scale, int accumulate, float* d_qkv, cudaStream_t st, const float* lse,
                         const float* Dv) {
  MAED_CHECK_ARG((lse == nullptr) == (Dv == nullptr), "attn_temporal_bwd_tc: lse and D come together");
  if (lse) {
    std::vector<float> chk((size_t)B * T * ntok * heads), o((size_t)B * T * ntok * heads * 64);
    attention_ref(qkv_hi, qkv_plane, (long long)B * ntok, T, heads, scale,
                  [=](long long g, int i) { return ((g / ntok) * T + i) * ntok + (g % ntok); }, o.data(), nullptr, 0, chk.data());
    for (long long i = 0; i < (long long)B * T * ntok * heads; i += 37)
      MAED_CHECK_ARG(fabsf(chk[i] - lse[i]) <= 1e-3f * (1.f + fabsf(chk[i])), "attn_temporal_bwd_tc: lse[%lld] = %f, expected %f", i,
                     (double)lse[i], (double)chk[i]);
  }
  MAED_CHECK_ARG(qkv_hi && dout_hi && d_qkv, "attn_temporal_bwd_tc: null argument");
  MAED_CHECK_ARG(T == 4 || T == 8 || T == 16 || T == 32, "attn_temporal_bwd_tc: T=%d unsupported (4, 8, 16, 32)", T);
  const long long rows = (long long)B * T * ntok;
  const int ld3 = 3 * heads * 64, ldo = heads * 64;
  MAED_CHECK_ARG(qkv_plane >= rows * ld3 && dout_plane >= rows * ldo && qkv_plane % 8 == 0 && dout_plane % 8 == 0,
                 "attn_temporal_bwd_tc: operand planes overlap or are misaligned");
  std::vector<float> d((size_t)rows * ldo);
  planes_to_dense(dout_hi, dout_plane, rows, ldo, ldo, d.data());
  return attn_temporal_bwd(qkv_hi, qkv_plane, d.data(), B, T, ntok, heads, scale, accumulate, d_qkv, st);
}

int gemm_wgrad_rows(const __half* dY, long long dy_plane, int ld_dy, const __half* X, long long x_plane, int ld_x, int No_x,
                    int Mo, int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, cudaStream_t);
// contract: explicit im2col (the real CUDA-core kernel) followed by the row-major weight gradient
int gemm_wgrad_conv(const __half* dY, long long dy_plane, const __half* X, long long x_plane, int n_img, int H, int W, int Cin,
                    int Cout, int KH, int KW, int pad, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd,
                    cudaStream_t st) {
  MAED_CHECK_ARG(dY && X && slabs && D, "gemm_wgrad_conv: null argument");
  MAED_CHECK_ARG(Cin % 64 == 0 && Cout % 64 == 0 && KH >= 1 && KW >= 1 && n_img >= 1, "gemm_wgrad_conv: bad shape Cin=%d Cout=%d",
                 Cin, Cout);
  MAED_CHECK_ARG(nsplit == 3, "gemm_wgrad_conv: split precision only");
  const int No = KH * KW * Cin;
  MAED_CHECK_ARG(ldd >= No, "gemm_wgrad_conv: ldd=%d < %d", ldd, No);
  const long long M = (long long)n_img * H * W;
  std::vector<__half> col((size_t)2 * M * No);
  MAED_PROPAGATE(im2col_nhwc(X, x_plane, n_img, H, W, Cin, KH, KW, 1, pad, pad, H, W, col.data(), M * No, st));
  return gemm_wgrad_rows(dY, dy_plane, Cout, col.data(), M * No, No, No, Cout, No, (int)M, nsplit, scale, accumulate, slabs, D, ldd, st);
}

int gemm_wgrad_rows(const __half* dY, long long dy_plane, int ld_dy, const __half* X, long long x_plane, int ld_x, int No_x,
                    int Mo, int No, int R, int nsplit, float scale, int accumulate, float* slabs, float* D, int ldd, cudaStream_t) {
  Prof prof(4);
  MAED_CHECK_ARG(dY && X && slabs && D, "gemm_wgrad_rows: null argument");
  MAED_CHECK_ARG(Mo >= 1 && No >= 32 && No % 32 == 0 && R >= 1 && No_x >= 1 && No_x <= No,
                 "gemm_wgrad_rows: bad shape Mo=%d No=%d (%d) R=%d", Mo, No, No_x, R);
  MAED_CHECK_ARG(ld_dy % 8 == 0 && ld_x % 8 == 0 && ld_dy >= Mo && ld_x >= No_x, "gemm_wgrad_rows: row strides must be multiples "
                 "of 8 and cover the rows (ld_dy=%d Mo=%d ld_x=%d No=%d)", ld_dy, Mo, ld_x, No_x);
  MAED_CHECK_ARG(nsplit == 1 || nsplit == 3, "gemm_wgrad_rows: nsplit must be 1 or 3");
  MAED_CHECK_ARG(ldd >= No, "gemm_wgrad_rows: ldd=%d < No=%d", ldd, No);
  const int np = nsplit == 3 ? 2 : 1;
  {
    const uint64_t dims[3] = {(uint64_t)Mo, (uint64_t)R, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ld_dy * 2, (uint64_t)(np == 2 ? dy_plane : (long long)R * ld_dy) * 2};
    const uint32_t box[3] = {64, 64, 1};
    MAED_PROPAGATE(check_tmap("wgrad dY", dY, 3, dims, str, box));
    MAED_CHECK_ARG(np == 1 || dy_plane >= (long long)(R - 1) * ld_dy + Mo, "wgrad: dY planes overlap");
  }
  {
    const uint64_t dims[3] = {(uint64_t)No_x, (uint64_t)R, (uint64_t)np};
    const uint64_t str[2] = {(uint64_t)ld_x * 2, (uint64_t)(np == 2 ? x_plane : (long long)R * ld_x) * 2};
    const uint32_t box[3] = {64, 64, 1};
    MAED_PROPAGATE(check_tmap("wgrad X", X, 3, dims, str, box));
    MAED_CHECK_ARG(np == 1 || x_plane >= (long long)(R - 1) * ld_x + No_x, "wgrad: X planes overlap");
  }
  // dense fp32 copies of hi + lo, transposed to [Mo, R] / [No, R] for the shared NT kernel
  std::vector<float> Yf((size_t)R * Mo), Xf((size_t)R * No_x), At((size_t)Mo * R), Bt((size_t)No * R, 0.f), C((size_t)Mo * No);
  planes_to_dense(dY, np == 2 ? dy_plane : 0, R, ld_dy, Mo, Yf.data());
  planes_to_dense(X, np == 2 ? x_plane : 0, R, ld_x, No_x, Xf.data());
  for (long long r = 0; r < R; ++r) {
    for (int m = 0; m < Mo; ++m) At[(size_t)m * R + r] = Yf[(size_t)r * Mo + m];
    for (int n = 0; n < No_x; ++n) Bt[(size_t)n * R + r] = Xf[(size_t)r * No_x + n];
  }
  sgemm_nt(Mo, No, R, At.data(), Bt.data(), C.data());
  for (long long m = 0; m < Mo; ++m)
    for (int n = 0; n < No; ++n) {
      const float v = scale * C[(size_t)m * No + n];
      D[m * ldd + n] = accumulate ? D[m * ldd + n] + v : v;
    }
  slabs[0] = 0.f;                              // the real kernel scribbles over the slab scratch
  count_launch(2);
  return MAED_OK;
}

}  // namespace maed
