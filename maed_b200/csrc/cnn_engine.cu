// 'cnn' encoder of MAED: torchvision ResNet-50 with fc = Identity (reference lib/models/maed.py:35-37, the stage-1
// configuration configs/config_stage1.yaml:68) feeding the same KTD / iterative decoders (engine.cu run_decoder).
// Inference only: BatchNorm uses its running statistics and is folded into the conv weights when they are packed
// (conv + BN + ReLU = one tcgen05 GEMM with a bias / ReLU epilogue).
//
//   stem     conv1 7x7/2 pad 3 (explicit im2col from the fp32 NCHW frame, K padded 147 -> 152) + BN + ReLU -> fp32
//            MaxPool(3, 2, 1) -> planes
//   layer1-4 Bottleneck x (3, 4, 6, 3), stride on the 3x3 conv (v1.5):
//              conv1 1x1 + BN + ReLU                      -> planes    (plain GEMM)
//              conv2 3x3 (stride 1 or 2) + BN + ReLU      -> planes    (implicit GEMM by 5-D TMA; stride 2 and the 7x7 maps
//                                                                       of layer4 through an explicit im2col)
//              conv3 1x1 + BN + identity, then ReLU       -> planes    (GEMM epilogue: bias + plane residual + post-ReLU)
//            identity = block input, or downsample conv 1x1 (stride s) + BN -> planes in the first block of a layer
//   avgpool  the last conv3 writes fp32; mean over the 7x7 positions -> feature [BT, 2048]
//
// Every conv + BN (+ identity) + ReLU is ONE gemm_tc_kernel launch; activations exist only as fp16 hi/lo planes.
#include <algorithm>
#include <string>

#include "engine_internal.h"
#include "gemm_host.h"
#include "gemm_sm100.cuh"
#include "kernels.h"

namespace maed {

static const int kCnnDepth[4] = {3, 4, 6, 3};
static const int kCnnMid[4] = {64, 128, 256, 512};
static constexpr float kBnEps = 1e-5f;                       // nn.BatchNorm2d default
static constexpr long long kMapPerImg = 12544LL * 64;        // largest activation map per frame (stem conv = layer1 output)

static Engine::CnnConv make_conv(Engine& e, int (*add_param)(Engine&, const std::string&, long long), const std::string& conv,
                                 const std::string& bn, int cin, int cout, int k, int stride) {
  Engine::CnnConv c;
  c.cin = cin; c.cout = cout; c.k = k; c.stride = stride;
  c.k_pad = (k * k * cin + 7) / 8 * 8;
  c.off_w = c.off_b = c.off_w_raw = 0;
  c.w = add_param(e, conv + ".weight", (long long)cout * cin * k * k);
  c.bn = add_param(e, bn + ".weight", cout);
  add_param(e, bn + ".bias", cout);
  add_param(e, bn + ".running_mean", cout);
  add_param(e, bn + ".running_var", cout);
  return c;
}

// torchvision state_dict keys (fc is nn.Identity: no keys; num_batches_tracked is an int64 counter the forward never reads)
void cnn_add_params(Engine& e, int (*add_param)(Engine&, const std::string&, long long)) {
  const std::string enc = "encoder.";
  e.cnn.push_back(make_conv(e, add_param, enc + "conv1", enc + "bn1", 3, 64, 7, 2));
  int prev = 64;
  for (int l = 0; l < 4; ++l) {
    const int mid = kCnnMid[l], out = mid * 4;
    for (int b = 0; b < kCnnDepth[l]; ++b) {
      const std::string p = enc + "layer" + std::to_string(l + 1) + "." + std::to_string(b) + ".";
      const int stride = (l > 0 && b == 0) ? 2 : 1;
      if (b == 0) e.cnn.push_back(make_conv(e, add_param, p + "downsample.0", p + "downsample.1", prev, out, 1, stride));
      e.cnn.push_back(make_conv(e, add_param, p + "conv1", p + "bn1", prev, mid, 1, 1));
      e.cnn.push_back(make_conv(e, add_param, p + "conv2", p + "bn2", mid, mid, 3, stride));
      e.cnn.push_back(make_conv(e, add_param, p + "conv3", p + "bn3", mid, out, 1, 1));
      prev = out;
    }
  }
}

void cnn_add_packed(Engine& e, size_t& off) {
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 1023) / 1024 * 1024; return o; };
  long long max_w = 0;
  for (Engine::CnnConv& c : e.cnn) {
    c.off_w = take((size_t)c.cout * c.k_pad * 2 * 2);
    c.off_b = take((size_t)c.cout * 4);
    c.off_w_raw = take((size_t)c.cout * c.k_pad * 2 * 2);
    max_w = std::max(max_w, (long long)c.cout * c.cin * c.k * c.k);
  }
  e.off_cnn_scratch = take((size_t)max_w * 4);
}

int cnn_pack(const Engine& e, const void* const* params, void* packed, cudaStream_t st) {
  uint8_t* pk = (uint8_t*)packed;
  auto P = [&](int i) { return (const float*)params[i]; };
  float* scratch = (float*)(pk + e.off_cnn_scratch);
  for (const Engine::CnnConv& c : e.cnn) {
    const long long E = (long long)c.cin * c.k * c.k;
    MAED_PROPAGATE(fold_bn(P(c.w), c.cout, E, P(c.bn), P(c.bn + 1), P(c.bn + 2), P(c.bn + 3), kBnEps, scratch,
                           (float*)(pk + c.off_b), st));
    MAED_PROPAGATE(prep_conv_weight(scratch, c.cout, c.cin, c.k, c.k, c.k_pad, 0, (__half*)(pk + c.off_w),
                                    (long long)c.cout * c.k_pad, st));
    MAED_PROPAGATE(prep_conv_weight(P(c.w), c.cout, c.cin, c.k, c.k, c.k_pad, 0, (__half*)(pk + c.off_w_raw),
                                    (long long)c.cout * c.k_pad, st));
  }
  return MAED_OK;
}

// ---- workspace
namespace {
struct CnnWs {
  __half* col; long long col_plane;          // explicit im2col matrix (stem; stride-2 convs; 3x3 convs on 7x7 maps)
  __half* p[2]; __half* t1; __half* t2;      // block input / output planes (ping-pong), conv1 and conv2 outputs
  __half* pds;                               // downsample output (the identity of a layer's first block)
  float* f32;                                // stem conv output (before the max-pool); the last block's output
  long long plane;
  TailWs tail;
  size_t total;
};
void cnn_carve(const Engine& e, int BT, uint8_t* base, CnnWs& w) {
  Carver cv(base);
  w.col_plane = 12544LL * 152 * BT;
  w.plane = kMapPerImg * BT;
  w.col = (__half*)cv.take((size_t)w.col_plane * 2 * 2);
  for (int i = 0; i < 2; ++i) w.p[i] = (__half*)cv.take((size_t)w.plane * 2 * 2);
  w.t1 = (__half*)cv.take((size_t)w.plane * 2 * 2);
  w.t2 = (__half*)cv.take((size_t)w.plane * 2 * 2);
  w.pds = (__half*)cv.take((size_t)w.plane * 2 * 2);
  w.f32 = (float*)cv.take((size_t)w.plane * 4);
  carve_tail(e, BT, cv, w.tail);
  w.total = cv.off;
}
}  // namespace

size_t cnn_workspace_bytes(const Engine& e, int BT) {
  CnnWs w;
  cnn_carve(e, BT, nullptr, w);
  return w.total + 1024;
}

// conv + folded BN (+ ReLU | + identity planes, then ReLU) of an NHWC map [BT, Hin, Hin, cin] held as planes
static int conv_bn(const Engine& e, const uint8_t* pk, CnnWs& w, const Engine::CnnConv& c, int BT, const __half* in, int Hin,
                   int relu, const __half* identity, int out_mode, void* out, cudaStream_t st) {
  const int pad = c.k / 2;
  const int Hout = (Hin + 2 * pad - c.k) / c.stride + 1;
  GemmArgs g;
  g.nsplit = e.cfg.nsplit;
  g.B = (const __half*)(pk + c.off_w); g.b_plane = (long long)c.cout * c.k_pad; g.ldb = c.k_pad;
  g.M = BT * Hout * Hout; g.N = c.cout; g.K = c.k * c.k * c.cin;
  g.bias = (const float*)(pk + c.off_b);
  g.act = (relu && !identity) ? ACT_RELU : ACT_NONE;
  if (identity) {                                            // relu(conv3 + bn3 + identity)
    g.res_hi = identity; g.res_plane = e.cfg.nsplit == 3 ? w.plane : 0; g.act_post = relu ? ACT_RELU : ACT_NONE;
  }
  g.out_mode = out_mode; g.out = out; g.out_plane = w.plane; g.ldc = c.cout;
  if (c.k == 1 && c.stride == 1) {
    g.A = in; g.a_plane = w.plane;
  } else if (c.stride == 1 && Hin >= 14) {
    g.A = in; g.a_plane = w.plane;
    g.conv = 1; g.n_img = BT; g.H = Hin; g.W = Hin; g.Cin = c.cin; g.KH = c.k; g.KW = c.k; g.pad_h = pad; g.pad_w = pad;
  } else {
    MAED_CHECK_ARG((long long)g.M * g.K <= w.col_plane, "cnn: im2col matrix %d x %d exceeds the workspace", g.M, g.K);
    MAED_PROPAGATE(im2col_nhwc(in, w.plane, BT, Hin, Hin, c.cin, c.k, c.k, c.stride, pad, pad, Hout, Hout, w.col, w.col_plane, st));
    g.A = w.col; g.a_plane = w.col_plane;
  }
  return launch_gemm(g, st);
}

int cnn_forward(const Engine& e, const void* const* params, const void* packed, const float* x_in, int N, int T, void* workspace,
                size_t workspace_bytes, const EngineOutputs* outs, float* const* taps, cudaStream_t st) {
  const int BT = N * T;
  CnnWs w;
  cnn_carve(e, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "engine_forward(cnn): workspace too small (%zu < %zu)", workspace_bytes,
                 w.total + 1024);
  const uint8_t* pk = (const uint8_t*)packed;
  const int om = e.cfg.nsplit == 3 ? OUT_F16_SPLIT : OUT_F16;
  const long long tap_plane = e.cfg.nsplit == 3 ? w.plane : 0;
  auto tap = [&](int which, const __half* src, long long n) -> int {
    if (taps && taps[which]) MAED_PROPAGATE(planes_to_f32(src, tap_plane, n, taps[which], st));
    return MAED_OK;
  };

  // ---- stem: conv1 7x7/2 pad 3 + BN + ReLU (112x112x64, fp32) -> MaxPool(3, 2, 1) (56x56x64, planes)
  size_t ci = 0;
  {
    const Engine::CnnConv& c = e.cnn[ci++];
    MAED_PROPAGATE(im2col_stem(x_in, BT, 3, 224, 224, 7, 7, 2, 3, 3, 112, 112, c.k_pad, w.col, w.col_plane, st));
    GemmArgs g;
    g.nsplit = e.cfg.nsplit;
    g.A = w.col; g.a_plane = w.col_plane; g.lda = c.k_pad;
    g.B = (const __half*)(pk + c.off_w); g.b_plane = (long long)c.cout * c.k_pad; g.ldb = c.k_pad;
    g.M = BT * 12544; g.N = 64; g.K = c.k_pad;
    g.bias = (const float*)(pk + c.off_b); g.act = ACT_RELU; g.out_mode = OUT_F32; g.out = w.f32; g.ldc = 64;
    MAED_PROPAGATE(launch_gemm(g, st));
    MAED_PROPAGATE(maxpool3x3s2(w.f32, BT, 112, 112, 64, nullptr, w.p[0], w.plane, st));
  }
  MAED_PROPAGATE(tap(TAP_STEM, w.p[0], (long long)BT * 3136 * 64));

  // ---- layer1..layer4
  int cur = 0, Hc = 56, prev = 64;
  for (int l = 0; l < 4; ++l) {
    const int mid = kCnnMid[l], out = mid * 4;
    for (int b = 0; b < kCnnDepth[l]; ++b) {
      const int stride = (l > 0 && b == 0) ? 2 : 1;
      const int Ho = Hc / stride;
      const bool last = (l == 3 && b == kCnnDepth[l] - 1);
      const __half* identity = w.p[cur];
      if (b == 0) {
        MAED_PROPAGATE(conv_bn(e, pk, w, e.cnn[ci++], BT, w.p[cur], Hc, 0, nullptr, om, w.pds, st));
        identity = w.pds;
      }
      MAED_PROPAGATE(conv_bn(e, pk, w, e.cnn[ci++], BT, w.p[cur], Hc, 1, nullptr, om, w.t1, st));
      MAED_PROPAGATE(conv_bn(e, pk, w, e.cnn[ci++], BT, w.t1, Hc, 1, nullptr, om, w.t2, st));
      if (last)                                               // fp32 for the average pool
        MAED_PROPAGATE(conv_bn(e, pk, w, e.cnn[ci++], BT, w.t2, Ho, 1, identity, OUT_F32, w.f32, st));
      else
        MAED_PROPAGATE(conv_bn(e, pk, w, e.cnn[ci++], BT, w.t2, Ho, 1, identity, om, w.p[cur ^ 1], st));
      cur ^= 1;
      Hc = Ho;
      prev = out;
    }
    // TAP_STAGE0..2 = layer1..3; layer4 lands in the TAP_EMBED slot (fp32 NHWC)
    if (l < 3) {
      MAED_PROPAGATE(tap(TAP_STAGE0 + l, w.p[cur], (long long)BT * Hc * Hc * prev));
    } else if (taps && taps[TAP_EMBED]) {
      MAED_CUDA_CHECK(cudaMemcpyAsync(taps[TAP_EMBED], w.f32, (size_t)BT * Hc * Hc * prev * 4, cudaMemcpyDeviceToDevice, st));
    }
  }

  // ---- AdaptiveAvgPool2d(1) + flatten -> [BT, 2048]; fc = Identity
  MAED_PROPAGATE(token_mean(w.f32, BT, Hc * Hc, prev, outs->feat, prev, 0, st));
  return run_decoder(e, params, pk, BT, w.tail, outs, st);
}

}  // namespace maed
