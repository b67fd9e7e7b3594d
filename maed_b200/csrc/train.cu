// Training path of the MAED engine: forward with a saved-activation tape + backward to every parameter
// (reference: autograd of lib/models/maed.py:52-66 as driven by lib/core/trainer.py:238-255).
//
// Boundary: the decoder outputs pose6d / shape / cam (the geometry tail behind them — rot6d -> rotmat -> angle-axis,
// projection — is differentiated by autograd on the Python side; it is O(BT * 24) work).
//
//   train_forward : same arithmetic as engine_forward, but every layer writes into its own tape slot (GroupNorm runs
//                   unfused so the fp32 conv outputs and their statistics are kept; fc1 keeps its pre-GELU output).
//   train_backward: walks the tape in reverse.  All data-gradient and weight-gradient contractions run on the tensor
//                   cores in split precision: dX = dY * W through the forward GEMM kernel with transposed packed
//                   weights (3x3: the implicit-GEMM conv kernel with flipped taps), dW = dY^T * X through the split-K
//                   kernel on transposed activation planes.  Activation gradients carry `loss_scale`; parameter
//                   gradients are written un-scaled, in the reference's parameter layout, to `grads[i]`.
//
// Supported for training: st_mode parallel / series / vanilla with the KTD decoder (the published configurations).
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "bwd_kernels.h"
#include "engine_internal.h"
#include "fork_join.h"
#include "gemm_host.h"
#include "gemm_sm100.cuh"
#include "kernels.h"
#include "train.h"

namespace maed {

namespace {

size_t align_up(size_t v, size_t a = 1024) { return (v + a - 1) / a * a; }

// the tcgen05 temporal kernels (forward with row statistics, backward) apply when T is 4 / 8 / 16 / 32 and a token group is full
bool temporal_tc(int T, int ntok) {
  static const bool tc_on = [] { const char* v = getenv("MAED_B200_TEMPORAL_TC"); return !(v && v[0] == '0'); }();
  return tc_on && (T == 4 || T == 8 || T == 16 || T == 32) && ntok >= 128 / T;
}

// ------------------------------------------------------------------------------------ backbone layer table
struct ConvL {
  int Hin, Cin, Cout, k, stride, Hout, relu;
  int w_idx, g_idx;            // parameter indices (conv weight; norm weight, bias = +1)
  size_t pk_off;               // forward packed weight (engine pack)
  size_t tp_off;               // dgrad packed weight (train pack); stem: unused
  int in_layer;                // producing layer (-1: network input for the stem)
  int res_layer;               // layer whose output is added before the ReLU (-2: none)
  int pad = -1;                // top / left zero padding; -1: TF "SAME" (ResNetV2), torchvision's convs: k / 2
  size_t fw_off = 0;           // 'cnn' nets: un-folded forward weight planes in the ENGINE pack (BatchNorm runs on batch statistics)
  int pad_top() const { return pad >= 0 ? pad : std::max((Hout - 1) * stride + k - Hin, 0) / 2; }
  long long Min(int BT) const { return (long long)BT * Hin * Hin; }
  long long Mout(int BT) const { return (long long)BT * Hout * Hout; }
  int Kcols() const { return k * k * Cin; }
};
struct BlockL { int ds, c1, c2, c3, in_layer; };   // layer ids (ds = -1 when absent)

struct Net {
  std::vector<ConvL> L;        // 0 = stem (its "output" is the pooled 56x56x64 map), then the 52 bottleneck convs
  std::vector<BlockL> B;
  bool bn = false;             // 'cnn' encoder: BatchNorm2d (batch statistics) instead of weight standardisation + GroupNorm
  size_t tp_proj;              // dgrad weights of patch_embed.proj
  struct SteT { size_t qkv, proj, fc1, fc2, ts; };
  std::vector<SteT> ste;
  size_t tpack_bytes;
};

Net build_net(const Engine& e) {
  Net n;
  size_t off = 0;
  auto planes = [&](long long elems) { size_t o = off; off = align_up(off + (size_t)elems * 2 * 2); return o; };
  ConvL stem{224, 3, 64, 7, 2, 112, 1, e.i_stem_w, e.i_stem_g, e.off_stem, 0, -1, -2};
  n.L.push_back(stem);
  int prev = 64, Hc = 56, cur = 0, bi = 0;
  for (int s = 0; s < 3; ++s) {
    const int out = kStageOut[s], mid = out / 4;
    for (int b = 0; b < kStageDepth[s]; ++b, ++bi) {
      const Engine::BlockIdx& ix = e.bb[bi];
      const Engine::BlockOff& of = e.bb_off[bi];
      const int stride = (s > 0 && b == 0) ? 2 : 1;
      const int Ho = Hc / stride;
      BlockL bl;
      bl.in_layer = cur;
      bl.ds = -1;
      int shortcut = cur;
      if (b == 0) {
        n.L.push_back(ConvL{Hc, prev, out, 1, stride, Ho, 0, ix.ds_w, ix.ds_g, of.ds, planes((long long)out * prev), cur, -2});
        bl.ds = (int)n.L.size() - 1;
        shortcut = bl.ds;
      }
      n.L.push_back(ConvL{Hc, prev, mid, 1, 1, Hc, 1, ix.c1_w, ix.c1_g, of.c1, planes((long long)mid * prev), cur, -2});
      bl.c1 = (int)n.L.size() - 1;
      n.L.push_back(ConvL{Hc, mid, mid, 3, stride, Ho, 1, ix.c2_w, ix.c2_g, of.c2, planes((long long)mid * mid * 9), bl.c1, -2});
      bl.c2 = (int)n.L.size() - 1;
      n.L.push_back(ConvL{Ho, mid, out, 1, 1, Ho, 1, ix.c3_w, ix.c3_g, of.c3, planes((long long)out * mid), bl.c2, shortcut});
      bl.c3 = (int)n.L.size() - 1;
      n.B.push_back(bl);
      cur = bl.c3;
      prev = out;
      Hc = Ho;
    }
  }
  n.tp_proj = planes(768LL * 1024);
  const long long CC = 768LL * 768;
  for (int i = 0; i < e.cfg.num_blocks; ++i) {
    Net::SteT t;
    t.qkv = planes(3 * CC);
    t.proj = planes(CC);
    t.fc1 = planes(4 * CC);
    t.fc2 = planes(4 * CC);
    t.ts = e.cfg.mode == MODE_PARALLEL ? planes(4 * CC) : 0;        // attentive-addition linear [2C, 2C] (parallel mode only)
    n.ste.push_back(t);
  }
  n.tpack_bytes = off;
  return n;
}

// --------------------------------------------------------------------------------------------- workspace
struct SteTape {
  float* x_in; __half* ln1; __half* qkv; float* xs; float* xt; __half* ao_s; __half* qkv2; float* pooled; float* logits;
  __half* ao; float* x_mid; __half* ln2; float* h_pre; __half* hid;
  float* lse_s; float* lse_t;          // [rows, heads] log2-domain row log-sum-exp of the spatial / temporal attention (tcgen05 kernels)
};
struct TrainWs {
  // ---- tape
  std::vector<__half*> out; std::vector<float*> convout; std::vector<double*> stats;
  std::vector<long long> out_plane;
  unsigned char* pool_idx;
  std::vector<float*> bn_mean, bn_rstd;          // 'cnn': batch statistics of every BatchNorm (tape)
  double* bn_partial;                            // 'cnn': scratch of the BatchNorm reductions
  float* tok;                                    // proj output [BT*196, 768]
  std::vector<SteTape> ste;
  float* x_final;
  float* cls_ln; float* h1; float* h2; float* base; unsigned char* mask1; unsigned char* mask2;
  float* feat_copy; float* pose_copy;
  // iterative regressor (spin.py:51-74): input, hidden activations (post dropout) and dropout masks of the 3 iterations
  float* it_xc[3]; float* it_h1[3]; float* it_h2[3]; unsigned char* it_m1[3]; unsigned char* it_m2[3]; float* it_dxc;
  // ---- scratch shared by forward and backward
  __half* col; long long col_plane;              // im2col matrix / small planes
  __half* pl_a; long long pl_a_plane;            // generic planes (max rows*3072 or M*C)
  __half* dil; long long dil_plane;
  float* fa; float* fb; float* fc; float* fd;    // fp32 activation-gradient buffers (backbone size)
  float* big;                                    // fp32 [rows, 3072]
  float* dxs; float* dxt;
  float* attn_D;                                 // [rows, heads] D = rowsum(dO o O) of the attention backward
  float* wg;                                     // fp32 dW_hat temp
  float* slabs;
  float* red; float* dgb; float* colsum_scratch; float* ln_partial;
  float* small[8];                               // [BT, 2048] fp32 each
  __half* small_p; long long small_plane;
  float* anc_grad;
  long long ln_rows, ln_plane, qkv_plane, hid_plane;
  size_t total;
};

long long max_ll(long long a, long long b) { return a > b ? a : b; }

void carve(const Engine& e, const Net& net, int BT, uint8_t* base, TrainWs& w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off = align_up(off + bytes); return p; };
  const int nl = (int)net.L.size();
  w.out.resize(nl); w.convout.resize(nl); w.stats.resize(nl); w.out_plane.resize(nl);
  long long max_mc = 0, max_col = 0, max_slab = 0, max_wg = 0;
  for (int l = 0; l < nl; ++l) {
    const ConvL& L = net.L[l];
    const long long Mo = L.Mout(BT);
    const long long out_elems = (l == 0) ? (long long)BT * 3136 * 64 : Mo * L.Cout;
    w.out_plane[l] = out_elems;
    w.out[l] = (__half*)take((size_t)out_elems * 4);
    w.convout[l] = (float*)take((size_t)Mo * L.Cout * 4);
    w.stats[l] = (double*)take((size_t)BT * 64 * 8);
    max_mc = max_ll(max_mc, max_ll(Mo * L.Cout, L.Min(BT) * L.Cin));
    const int kc = (l == 0) ? kStemKPad : L.Kcols();
    const int kc_pad = (kc + 31) / 32 * 32;
    if (L.k > 1 || L.stride > 1) max_col = max_ll(max_col, Mo * kc);
    max_slab = max_ll(max_slab, (long long)splitk_slab_floats(L.Cout, kc_pad, (int)Mo));
    max_wg = max_ll(max_wg, (long long)L.Cout * kc_pad);
  }
  w.pool_idx = (unsigned char*)take((size_t)BT * 3136 * 64);
  const long long rows = (long long)BT * 197;
  const int C = 768;
  w.tok = (float*)take((size_t)BT * 196 * C * 4);
  w.ln_rows = rows; w.ln_plane = rows * C; w.qkv_plane = rows * 3 * C; w.hid_plane = rows * 4 * C;
  w.ste.resize(e.cfg.num_blocks);
  for (int i = 0; i < e.cfg.num_blocks; ++i) {
    SteTape& t = w.ste[i];
    t.x_in = (float*)take((size_t)rows * C * 4);
    t.ln1 = (__half*)take((size_t)w.ln_plane * 4);
    t.qkv = (__half*)take((size_t)w.qkv_plane * 4);
    t.xs = (float*)take((size_t)rows * C * 4);
    t.xt = (float*)take((size_t)rows * C * 4);
    t.ao_s = (e.cfg.mode == MODE_SERIES) ? (__half*)take((size_t)w.ln_plane * 4) : nullptr;
    t.qkv2 = (e.cfg.mode == MODE_SERIES) ? (__half*)take((size_t)w.qkv_plane * 4) : nullptr;
    t.pooled = (float*)take((size_t)BT * 2 * C * 4);
    t.logits = (float*)take((size_t)BT * 2 * C * 4);
    t.ao = (__half*)take((size_t)w.ln_plane * 4);
    t.x_mid = (float*)take((size_t)rows * C * 4);
    t.ln2 = (__half*)take((size_t)w.ln_plane * 4);
    t.h_pre = (float*)take((size_t)rows * 4 * C * 4);
    t.hid = (__half*)take((size_t)w.hid_plane * 4);
    t.lse_s = (float*)take((size_t)rows * e.cfg.num_heads * 4);
    t.lse_t = (float*)take((size_t)rows * e.cfg.num_heads * 4);
  }
  w.x_final = (float*)take((size_t)rows * C * 4);
  const int HD = e.cfg.hidden_dim;
  w.cls_ln = (float*)take((size_t)BT * C * 4);
  w.h1 = (float*)take((size_t)BT * HD * 4);
  w.h2 = (float*)take((size_t)BT * HD * 4);
  w.base = (float*)take((size_t)BT * 192 * 4);
  w.mask1 = (unsigned char*)take((size_t)BT * HD);
  w.mask2 = (unsigned char*)take((size_t)BT * HD);
  w.feat_copy = (float*)take((size_t)BT * C * 4);
  w.pose_copy = (float*)take((size_t)BT * 144 * 4);
  if (e.cfg.decoder == DEC_ITERATIVE) {
    const int KI = e.feat_dim() + 157;
    for (int i = 0; i < 3; ++i) {
      w.it_xc[i] = (float*)take((size_t)BT * KI * 4);
      w.it_h1[i] = (float*)take((size_t)BT * HD * 4);
      w.it_h2[i] = (float*)take((size_t)BT * HD * 4);
      w.it_m1[i] = (unsigned char*)take((size_t)BT * HD);
      w.it_m2[i] = (unsigned char*)take((size_t)BT * HD);
    }
    w.it_dxc = (float*)take((size_t)BT * KI * 4);
  }
  // ---- scratch
  const long long ste_big = rows * 4 * C;                                   // rows x 3072
  const long long pl_elems = max_ll(max_mc, ste_big);
  w.col_plane = max_ll(max_col, 8); w.col = (__half*)take((size_t)w.col_plane * 4);
  w.pl_a_plane = pl_elems; w.pl_a = (__half*)take((size_t)pl_elems * 4);
  w.dil_plane = max_mc; w.dil = (__half*)take((size_t)max_mc * 4);
  w.fa = (float*)take((size_t)max_mc * 4);
  w.fb = (float*)take((size_t)max_mc * 4);
  w.fc = (float*)take((size_t)max_mc * 4);
  w.fd = (float*)take((size_t)max_mc * 4);
  w.big = (float*)take((size_t)ste_big * 4);
  w.dxs = (float*)take((size_t)rows * C * 4);
  w.dxt = (float*)take((size_t)rows * C * 4);
  w.attn_D = (float*)take((size_t)rows * e.cfg.num_heads * 4);
  // split-K slabs / weight-gradient temp: also cover the STE and proj linears
  const int lin_shapes[6][2] = {{3 * C, C}, {C, C}, {4 * C, C}, {C, 4 * C}, {C, 1024}, {2 * C, 2 * C}};
  for (int i = 0; i < 6; ++i) {
    max_slab = max_ll(max_slab, (long long)splitk_slab_floats(lin_shapes[i][0], lin_shapes[i][1], (int)rows));
    max_wg = max_ll(max_wg, (long long)lin_shapes[i][0] * lin_shapes[i][1]);
  }
  w.wg = (float*)take((size_t)max_wg * 4);
  w.slabs = (float*)take((size_t)max_slab * 4);
  w.red = (float*)take((size_t)BT * (64 + 32 * 1024) * 4);
  w.dgb = (float*)take((size_t)BT * 2 * 1024 * 4);
  w.colsum_scratch = (float*)take((size_t)kColsumChunks * 197 * C * 4);
  w.ln_partial = (float*)take((size_t)ln_bwd_partial_rows() * 2 * C * 4);
  for (int i = 0; i < 8; ++i) w.small[i] = (float*)take((size_t)BT * 2048 * 4);
  w.small_plane = (long long)BT * 2048; w.small_p = (__half*)take((size_t)w.small_plane * 4);
  w.anc_grad = (float*)take(36 * 95 * 4);
  w.total = off;
}

// ------------------------------------------------------------------------------------------ 'cnn' encoder
// torchvision ResNet-50 (cnn_engine.cu) as a layer table: 0 = stem (its "output" is the max-pooled 56x56x64 map), then per
// bottleneck [downsample], conv1, conv2, conv3 in the order of Engine::cnn.  Symmetric padding k / 2.
Net build_cnn_net(const Engine& e) {
  Net n;
  n.bn = true;
  size_t off = 0;
  auto planes = [&](long long elems) { size_t o = off; off = align_up(off + (size_t)elems * 2 * 2); return o; };
  size_t ci = 0;
  auto add = [&](int Hin, int Cin, int Cout, int k, int stride, int relu, int in_layer, int res_layer) {
    const Engine::CnnConv& cc = e.cnn[ci++];
    ConvL L{Hin, Cin, Cout, k, stride, (Hin + 2 * (k / 2) - k) / stride + 1, relu, cc.w, cc.bn, 0, 0, in_layer, res_layer};
    L.pad = k / 2;
    L.tp_off = planes((long long)Cout * Cin * k * k);
    L.fw_off = cc.off_w_raw;
    n.L.push_back(L);
    return (int)n.L.size() - 1;
  };
  add(224, 3, 64, 7, 2, 1, -1, -2);
  static const int depth[4] = {3, 4, 6, 3}, mids[4] = {64, 128, 256, 512};
  int prev = 64, Hc = 56, cur = 0;
  for (int l = 0; l < 4; ++l) {
    const int mid = mids[l], out = 4 * mid;
    for (int b = 0; b < depth[l]; ++b) {
      const int stride = (l > 0 && b == 0) ? 2 : 1;
      const int Ho = Hc / stride;
      BlockL bl;
      bl.in_layer = cur;
      bl.ds = -1;
      int shortcut = cur;
      if (b == 0) {
        bl.ds = add(Hc, prev, out, 1, stride, 0, cur, -2);
        shortcut = bl.ds;
      }
      bl.c1 = add(Hc, prev, mid, 1, 1, 1, cur, -2);
      bl.c2 = add(Hc, mid, mid, 3, stride, 1, bl.c1, -2);
      bl.c3 = add(Ho, mid, out, 1, 1, 1, bl.c2, shortcut);
      n.B.push_back(bl);
      cur = bl.c3;
      prev = out;
      Hc = Ho;
    }
  }
  n.tp_proj = 0;
  n.tpack_bytes = off;
  return n;
}

void carve_cnn(const Engine& e, const Net& net, int BT, uint8_t* base, TrainWs& w) {
  size_t off = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base ? base + off : nullptr; off = align_up(off + bytes); return p; };
  const int nl = (int)net.L.size();
  w.out.resize(nl); w.convout.resize(nl); w.stats.assign(nl, nullptr); w.out_plane.resize(nl);
  w.bn_mean.resize(nl); w.bn_rstd.resize(nl);
  long long max_mc = 0, max_col = 0, max_slab = 0, max_wg = 0;
  size_t max_partial = 0;
  for (int l = 0; l < nl; ++l) {
    const ConvL& L = net.L[l];
    const long long Mo = L.Mout(BT);
    const long long out_elems = (l == 0) ? (long long)BT * 3136 * 64 : Mo * L.Cout;
    w.out_plane[l] = out_elems;
    w.out[l] = (__half*)take((size_t)out_elems * 4);
    w.convout[l] = (float*)take((size_t)Mo * L.Cout * 4);
    w.bn_mean[l] = (float*)take((size_t)L.Cout * 4);
    w.bn_rstd[l] = (float*)take((size_t)L.Cout * 4);
    max_mc = max_ll(max_mc, max_ll(Mo * L.Cout, L.Min(BT) * L.Cin));
    const int kc = (l == 0) ? kStemKPad : L.Kcols();
    const int kc_pad = (kc + 31) / 32 * 32;
    if (L.k > 1 || L.stride > 1) max_col = max_ll(max_col, Mo * kc);
    max_slab = max_ll(max_slab, (long long)splitk_slab_floats(L.Cout, kc_pad, (int)Mo));
    max_wg = max_ll(max_wg, (long long)L.Cout * kc_pad);
    max_partial = std::max(max_partial, bn_scratch_doubles(Mo, L.Cout));
  }
  w.pool_idx = (unsigned char*)take((size_t)BT * 3136 * 64);
  const int HD = e.cfg.hidden_dim, F = e.feat_dim();
  w.h1 = (float*)take((size_t)BT * HD * 4);
  w.h2 = (float*)take((size_t)BT * HD * 4);
  w.base = (float*)take((size_t)BT * 192 * 4);
  w.mask1 = (unsigned char*)take((size_t)BT * HD);
  w.mask2 = (unsigned char*)take((size_t)BT * HD);
  w.feat_copy = (float*)take((size_t)BT * F * 4);
  w.pose_copy = (float*)take((size_t)BT * 144 * 4);
  if (e.cfg.decoder == DEC_ITERATIVE) {
    const int KI = e.feat_dim() + 157;
    for (int i = 0; i < 3; ++i) {
      w.it_xc[i] = (float*)take((size_t)BT * KI * 4);
      w.it_h1[i] = (float*)take((size_t)BT * HD * 4);
      w.it_h2[i] = (float*)take((size_t)BT * HD * 4);
      w.it_m1[i] = (unsigned char*)take((size_t)BT * HD);
      w.it_m2[i] = (unsigned char*)take((size_t)BT * HD);
    }
    w.it_dxc = (float*)take((size_t)BT * KI * 4);
  }
  w.col_plane = max_ll(max_col, 8); w.col = (__half*)take((size_t)w.col_plane * 4);
  w.pl_a_plane = max_mc; w.pl_a = (__half*)take((size_t)max_mc * 4);
  w.dil_plane = max_mc; w.dil = (__half*)take((size_t)max_mc * 4);
  w.fa = (float*)take((size_t)max_mc * 4);
  w.fb = (float*)take((size_t)max_mc * 4);
  w.fc = (float*)take((size_t)max_mc * 4);
  w.fd = (float*)take((size_t)max_mc * 4);
  w.wg = (float*)take((size_t)max_wg * 4);
  w.slabs = (float*)take((size_t)max_slab * 4);
  w.colsum_scratch = (float*)take((size_t)kColsumChunks * 2048 * 4);
  w.bn_partial = (double*)take(max_partial * 8);
  for (int i = 0; i < 8; ++i) w.small[i] = (float*)take((size_t)BT * 2048 * 4);
  w.small_plane = (long long)BT * 2048; w.small_p = (__half*)take((size_t)w.small_plane * 4);
  w.anc_grad = (float*)take(36 * 95 * 4);
  w.total = off;
}

struct Ctx {
  const Engine& e; const Net& net; TrainWs& w; const void* const* params; const uint8_t* pk; const uint8_t* tp;
  int BT, N, T; cudaStream_t st; float inv_ls; float* const* grads; const float* x_img;
  const float* P(int i) const { return (const float*)params[i]; }
  float* G(int i) const { return grads[i]; }
  const __half* H(size_t off) const { return (const __half*)(pk + off); }
  const __half* TH(size_t off) const { return (const __half*)(tp + off); }
};

int gemm_plain(const Ctx& c, const __half* A, long long a_plane, int M, int K, const __half* B, long long b_plane, int Nout,
               const float* bias, int act, const float* residual, int out_mode, void* out, long long out_plane) {
  GemmArgs g;
  g.nsplit = 3;
  g.A = A; g.a_plane = a_plane; g.B = B; g.b_plane = b_plane;
  g.M = M; g.N = Nout; g.K = K; g.bias = bias; g.act = act; g.residual = residual; g.out_mode = out_mode; g.out = out;
  g.out_plane = out_plane; g.ldc = Nout;
  return launch_gemm(g, c.st);
}

}  // namespace

// ================================================================================================ API
size_t train_pack_bytes(const Engine* e) {
  return (e->cfg.encoder == ENC_CNN ? build_cnn_net(*e) : build_net(*e)).tpack_bytes + 1024;
}

size_t train_workspace_bytes(const Engine* e, int BT) {
  TrainWs w;
  if (e->cfg.encoder == ENC_CNN) {
    carve_cnn(*e, build_cnn_net(*e), BT, nullptr, w);
  } else {
    const Net net = build_net(*e);
    carve(*e, net, BT, nullptr, w);
  }
  return w.total + 1024;
}

static int check_train_cfg(const Engine& e) {
  const int m = e.cfg.mode;
  MAED_CHECK_ARG(e.cfg.encoder == ENC_CNN || (m >= MODE_VANILLA && m <= MODE_TEMPORAL), "training: unknown st_mode %d", m);
  MAED_CHECK_ARG(e.cfg.nsplit == 3, "training runs in split precision (precision='split')");
  return MAED_OK;
}

static int cnn_train_pack(const Engine& e, const void* const* params, void* tpack, cudaStream_t st) {
  const Net net = build_cnn_net(e);
  uint8_t* tp = (uint8_t*)(((uintptr_t)tpack + 1023) & ~(uintptr_t)1023);
  auto P = [&](int i) { return (const float*)params[i]; };
  for (size_t l = 1; l < net.L.size(); ++l) {             // data-gradient weights (the stem needs none)
    const ConvL& L = net.L[l];
    MAED_PROPAGATE(prep_conv_weight_dgrad(P(L.w_idx), L.Cout, L.Cin, L.k, L.k, 0, (__half*)(tp + L.tp_off),
                                          (long long)L.Cout * L.Cin * L.k * L.k, st));
  }
  return MAED_OK;
}

int train_set_exchange(Engine* e, int (*fn)(void*, int), void* user, double* buf, int capacity) {
  MAED_CHECK_ARG(e, "train_set_exchange: null engine");
  MAED_CHECK_ARG(!fn || (buf && capacity >= 2 * 2048 + 1), "train_set_exchange: the buffer must hold 2 * 2048 + 1 doubles");
  e->bn_exchange = BnExchange{fn, user, buf, capacity};
  return MAED_OK;
}

int train_set_progress(Engine* e, int (*fn)(void*, int, int), void* user) {
  MAED_CHECK_ARG(e, "train_set_progress: null engine");
  e->progress = Engine::Progress{fn, user};
  return MAED_OK;
}

int train_pack(const Engine* ep, const void* const* params, void* tpack, cudaStream_t st) {
  MAED_CHECK_ARG(ep, "train_pack: null engine");
  MAED_PROPAGATE(check_train_cfg(*ep));
  MAED_CHECK_ARG(ep && params && tpack, "train_pack: null argument");
  const Engine& e = *ep;
  if (e.cfg.encoder == ENC_CNN) return cnn_train_pack(e, params, tpack, st);
  const Net net = build_net(e);
  uint8_t* tp = (uint8_t*)(((uintptr_t)tpack + 1023) & ~(uintptr_t)1023);
  auto P = [&](int i) { return (const float*)params[i]; };
  ForkJoin fj(st);                                         // ~85 independent small kernels: side streams, joined at the end
  for (size_t l = 1; l < net.L.size(); ++l) {
    const ConvL& L = net.L[l];
    MAED_PROPAGATE(prep_conv_weight_dgrad(P(L.w_idx), L.Cout, L.Cin, L.k, L.k, 1, (__half*)(tp + L.tp_off),
                                          (long long)L.Cout * L.Cin * L.k * L.k, fj.next()));
  }
  MAED_PROPAGATE(split_f32_transposed(P(e.i_proj_w), 768, 1024, (__half*)(tp + net.tp_proj), 768LL * 1024, fj.next()));
  const int C = 768;
  const long long CC = (long long)C * C;
  for (int i = 0; i < e.cfg.num_blocks; ++i) {
    const Engine::SteIdx& ix = e.blk[i];
    MAED_PROPAGATE(split_f32_transposed(P(ix.qkv_w), 3 * C, C, (__half*)(tp + net.ste[i].qkv), 3 * CC, fj.next()));
    MAED_PROPAGATE(split_f32_transposed(P(ix.proj_w), C, C, (__half*)(tp + net.ste[i].proj), CC, fj.next()));
    MAED_PROPAGATE(split_f32_transposed(P(ix.fc1_w), 4 * C, C, (__half*)(tp + net.ste[i].fc1), 4 * CC, fj.next()));
    MAED_PROPAGATE(split_f32_transposed(P(ix.fc2_w), C, 4 * C, (__half*)(tp + net.ste[i].fc2), 4 * CC, fj.next()));
    if (e.cfg.mode == MODE_PARALLEL)
      MAED_PROPAGATE(split_f32_transposed(P(ix.ts_w), 2 * C, 2 * C, (__half*)(tp + net.ste[i].ts), 4 * CC, fj.next()));
  }
  return fj.join();
}

// -------------------------------------------------------------------------------------------- forward
static int conv_fwd(const Ctx& c, int l) {
  const ConvL& L = c.net.L[l];
  TrainWs& w = c.w;
  const int BT = c.BT;
  const __half* A = w.out[L.in_layer];
  const long long a_plane = w.out_plane[L.in_layer];
  const long long M = L.Mout(BT);
  GemmArgs g;
  g.nsplit = 3;
  g.B = c.net.bn ? c.H(L.fw_off) : c.H(L.pk_off); g.b_plane = (long long)L.Cout * L.Kcols();
  g.M = (int)M; g.N = L.Cout; g.out_mode = OUT_F32; g.out = w.convout[l]; g.ldc = L.Cout;
  const int pad = L.pad_top();
  if (L.k == 1 && L.stride == 1) {
    g.A = A; g.a_plane = a_plane; g.K = L.Cin;
  } else if (L.stride == 1) {
    g.A = A; g.a_plane = a_plane; g.K = L.Kcols();
    g.conv = 1; g.n_img = BT; g.H = L.Hin; g.W = L.Hin; g.Cin = L.Cin; g.KH = L.k; g.KW = L.k;
    g.pad_h = pad; g.pad_w = pad;
  } else {
    MAED_PROPAGATE(im2col_nhwc(A, a_plane, BT, L.Hin, L.Hin, L.Cin, L.k, L.k, L.stride, pad, pad, L.Hout, L.Hout, w.col,
                               w.col_plane, c.st));
    g.A = w.col; g.a_plane = w.col_plane; g.K = L.Kcols();
  }
  MAED_PROPAGATE(launch_gemm(g, c.st));
  const __half* res = L.res_layer >= 0 ? w.out[L.res_layer] : nullptr;
  const long long res_plane = L.res_layer >= 0 ? w.out_plane[L.res_layer] : 0;
  if (c.net.bn) {                                          // BatchNorm2d.train(): batch statistics + running-buffer update
    MAED_PROPAGATE(bn_train_stats(w.convout[l], M, L.Cout, 1e-5f, 0.1f, w.bn_partial, w.bn_mean[l], w.bn_rstd[l],
                                  const_cast<float*>(c.P(L.g_idx + 2)), const_cast<float*>(c.P(L.g_idx + 3)), &c.e.bn_exchange, c.st));
    return bn_apply(w.convout[l], w.bn_mean[l], w.bn_rstd[l], c.P(L.g_idx), c.P(L.g_idx + 1), M, L.Cout, L.relu, res, res_plane,
                    w.out[l], w.out_plane[l], c.st);
  }
#ifndef MAED_EMU
  {                                                        // one cluster per image: statistics + apply with a single HBM read of convout
    const int rc = groupnorm_fwd_cluster(w.convout[l], c.P(L.g_idx), c.P(L.g_idx + 1), BT, L.Hout * L.Hout, L.Cout, 1e-5f, L.relu, res,
                                         res_plane, w.out[l], w.out_plane[l], w.stats[l], c.st);
    if (rc != MAED_ERR_UNSUPPORTED) return rc;
  }
#endif
  MAED_CUDA_CHECK(cudaMemsetAsync(w.stats[l], 0, (size_t)BT * 64 * 8, c.st));
  // the conv GEMM wrote convout upwards: statistics walk the images downwards (the tail is still in L2), the apply pass upwards
  MAED_PROPAGATE(gn_stats(w.convout[l], BT, L.Hout * L.Hout, L.Cout, w.stats[l], c.st, 1));
  MAED_PROPAGATE(gn_apply(w.convout[l], w.stats[l], c.P(L.g_idx), c.P(L.g_idx + 1), BT, L.Hout * L.Hout, L.Cout, 1e-5f, L.relu,
                          res, res_plane, w.out[l], w.out_plane[l], c.st));
  return MAED_OK;
}

// one BT-row linear of the tail: fp32 activation -> planes -> split-precision tensor-core GEMM -> fp32
static int tail_gemm(const Ctx& c, const float* a_f32, int K, size_t w_off, int Nout, const float* bias, int act, float* out) {
  TrainWs& w = c.w;
  MAED_PROPAGATE(split_f32(a_f32, w.small_p, w.small_plane, (long long)c.BT * K, c.st));
  return gemm_plain(c, w.small_p, w.small_plane, c.BT, K, c.H(w_off), (long long)Nout * K, Nout, bias, act, nullptr, OUT_F32, out, 0);
}

// KTD decoder forward from the encoder feature w.feat_copy [BT, F] (reference ktd.py:69-88), dropout masks kept
// iterative regressor forward (reference spin.py:51-74), fp32 CUDA-core GEMMs like the inference engine; tape: it_*
static int decoder_fwd_iterative(const Ctx& c, int C, float dropout_p, unsigned long long seed, const TrainOutputs* outs) {
  const Engine& e = c.e;
  TrainWs& w = c.w;
  const int BT = c.BT, HD = e.cfg.hidden_dim, KI = C + 157;
  cudaStream_t st = c.st;
  MAED_PROPAGATE(broadcast_row(c.P(e.i_init_pose), 144, BT, w.pose_copy, st));
  MAED_PROPAGATE(broadcast_row(c.P(e.i_init_shape), 10, BT, outs->shape, st));
  MAED_PROPAGATE(broadcast_row(c.P(e.i_init_cam), 3, BT, outs->cam, st));
  for (int it = 0; it < 3; ++it) {
    MAED_PROPAGATE(concat_cols(w.feat_copy, C, w.pose_copy, 144, outs->shape, 10, outs->cam, 3, BT, w.it_xc[it], st));
    MAED_PROPAGATE(linear_f32(w.it_xc[it], KI, c.P(e.i_fc1_w), KI, c.P(e.i_fc1_b), BT, HD, KI, 0, nullptr, 0, w.it_h1[it], HD, st));
    if (dropout_p > 0.f)
      MAED_PROPAGATE(dropout_fwd(w.it_h1[it], (long long)BT * HD, dropout_p, seed + 2 * it, w.it_m1[it], st));
    MAED_PROPAGATE(linear_f32(w.it_h1[it], HD, c.P(e.i_fc2_w), HD, c.P(e.i_fc2_b), BT, HD, HD, 0, nullptr, 0, w.it_h2[it], HD, st));
    if (dropout_p > 0.f)
      MAED_PROPAGATE(dropout_fwd(w.it_h2[it], (long long)BT * HD, dropout_p, (seed + 2 * it + 1) ^ 0x5851F42D4C957F2Dull, w.it_m2[it], st));
    MAED_PROPAGATE(linear_f32(w.it_h2[it], HD, c.P(e.i_decpose_w), HD, c.P(e.i_decpose_b), BT, 144, HD, 0, w.pose_copy, 144,
                              w.pose_copy, 144, st));
    MAED_PROPAGATE(linear_f32(w.it_h2[it], HD, c.P(e.i_shape_w), HD, c.P(e.i_shape_b), BT, 10, HD, 0, outs->shape, 10, outs->shape, 10, st));
    MAED_PROPAGATE(linear_f32(w.it_h2[it], HD, c.P(e.i_cam_w), HD, c.P(e.i_cam_b), BT, 3, HD, 0, outs->cam, 3, outs->cam, 3, st));
  }
  MAED_CUDA_CHECK(cudaMemcpyAsync(outs->pose6d, w.pose_copy, (size_t)BT * 144 * 4, cudaMemcpyDeviceToDevice, st));
  if (outs->feat) MAED_CUDA_CHECK(cudaMemcpyAsync(outs->feat, w.feat_copy, (size_t)BT * C * 4, cudaMemcpyDeviceToDevice, st));
  return MAED_OK;
}

static int decoder_fwd(const Ctx& c, int C, float dropout_p, unsigned long long seed, const TrainOutputs* outs) {
  const Engine& e = c.e;
  if (e.cfg.decoder == DEC_ITERATIVE) return decoder_fwd_iterative(c, C, dropout_p, seed, outs);
  TrainWs& w = c.w;
  const int BT = c.BT, HD = e.cfg.hidden_dim;
  cudaStream_t st = c.st;
  auto tail = [&](const float* a_f32, int K, size_t w_off, int Nout, const float* bias, int act, float* out) -> int {
    return tail_gemm(c, a_f32, K, w_off, Nout, bias, act, out);
  };
  MAED_PROPAGATE(tail(w.feat_copy, C, e.off_kfc1, HD, c.P(e.i_fc1_b), ACT_NONE, w.h1));
  if (dropout_p > 0.f) MAED_PROPAGATE(dropout_fwd(w.h1, (long long)BT * HD, dropout_p, seed, w.mask1, st));
  MAED_PROPAGATE(tail(w.h1, HD, e.off_kfc2, HD, c.P(e.i_fc2_b), ACT_NONE, w.h2));
  if (dropout_p > 0.f) MAED_PROPAGATE(dropout_fwd(w.h2, (long long)BT * HD, dropout_p, seed ^ 0x5851F42D4C957F2Dull, w.mask2, st));
  MAED_PROPAGATE(tail(w.h2, HD, e.off_kheads, 192, (const float*)(c.pk + e.off_kheads_b), ACT_NONE, w.base));
  MAED_PROPAGATE(ktd_tree(w.base, 192, (const float*)(c.pk + e.off_ktd_anc), BT, w.pose_copy, outs->shape, outs->cam, st));
  MAED_CUDA_CHECK(cudaMemcpyAsync(outs->pose6d, w.pose_copy, (size_t)BT * 144 * 4, cudaMemcpyDeviceToDevice, st));
  if (outs->feat) MAED_CUDA_CHECK(cudaMemcpyAsync(outs->feat, w.feat_copy, (size_t)BT * C * 4, cudaMemcpyDeviceToDevice, st));
  return MAED_OK;
}

// ================================================================================== 'cnn' encoder: training path
// conv (un-folded weights) -> BatchNorm2d on the statistics of the batch (running buffers updated in place, momentum 0.1)
// -> (+ identity) -> ReLU; MaxPool(3, 2, 1) keeps its arg-max; 7x7 average pool; KTD decoder.  SyncBatchNorm (reference
// train.py:95): when an exchange callback is installed (train_set_exchange) the per-channel sums of every BatchNorm, forward
// and backward, are added up over the data-parallel ranks before they are used.
static int cnn_train_forward(const Engine& e, const void* const* params, const void* packed, const float* x_in, int N, int T, void* workspace, size_t workspace_bytes, float dropout_p, unsigned long long seed,
                             const TrainOutputs* outs, cudaStream_t st) {
  const int BT = N * T;
  const Net net = build_cnn_net(e);
  TrainWs w;
  carve_cnn(e, net, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "train_forward(cnn): workspace too small (%zu < %zu)", workspace_bytes,
                 w.total + 1024);
  Ctx c{e, net, w, params, (const uint8_t*)packed, nullptr, BT, N, T, st, 1.f, nullptr, x_in};
  // ---- stem: conv1 7x7/2 pad 3 -> BatchNorm (batch statistics) -> ReLU -> MaxPool(3, 2, 1) with arg-max
  {
    const ConvL& L = net.L[0];
    MAED_PROPAGATE(im2col_stem(x_in, BT, 3, 224, 224, 7, 7, 2, 3, 3, 112, 112, kStemKPad, w.col, w.col_plane, st));
    MAED_PROPAGATE(gemm_plain(c, w.col, w.col_plane, BT * 12544, kStemKPad, c.H(L.fw_off), 64LL * kStemKPad, 64, nullptr, ACT_NONE,
                              nullptr, OUT_F32, w.convout[0], 0));
    MAED_PROPAGATE(bn_train_stats(w.convout[0], (long long)BT * 12544, 64, 1e-5f, 0.1f, w.bn_partial, w.bn_mean[0], w.bn_rstd[0],
                                  const_cast<float*>(c.P(L.g_idx + 2)), const_cast<float*>(c.P(L.g_idx + 3)), &e.bn_exchange, st));
    MAED_PROPAGATE(maxpool3x3s2_idx(w.convout[0], w.bn_mean[0], w.bn_rstd[0], c.P(L.g_idx), c.P(L.g_idx + 1), BT, 112, 112, 64,
                                    w.out[0], w.out_plane[0], w.pool_idx, st));
  }
  for (size_t l = 1; l < net.L.size(); ++l) MAED_PROPAGATE(conv_fwd(c, (int)l));
  // ---- AdaptiveAvgPool2d(1): mean over the 7x7 positions of layer4's output
  const int last = (int)net.L.size() - 1, F = e.feat_dim();
  MAED_PROPAGATE(planes_to_f32(w.out[last], w.out_plane[last], (long long)BT * 49 * F, w.fa, st));
  MAED_PROPAGATE(token_mean(w.fa, BT, 49, F, w.feat_copy, F, 0, st));
  return decoder_fwd(c, F, dropout_p, seed, outs);
}

int train_forward(const Engine* ep, const void* const* params, const void* packed, const float* x_in, int N, int T,
                  void* workspace, size_t workspace_bytes, float dropout_p, unsigned long long seed, const TrainOutputs* outs,
                  cudaStream_t st) {
  MAED_CHECK_ARG(ep, "train_forward: null engine");
  MAED_PROPAGATE(check_train_cfg(*ep));
  MAED_CHECK_ARG(ep && params && packed && x_in && workspace && outs, "train_forward: null argument");
  const Engine& e = *ep;
  const EngineConfig& cf = e.cfg;
  const int BT = N * T;
  MAED_CHECK_ARG(N >= 1 && T >= 1 && T <= 32, "train_forward: bad batch N=%d T=%d", N, T);
  MAED_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "train_forward: dropout_p=%f", (double)dropout_p);
  if (cf.encoder == ENC_CNN)
    return cnn_train_forward(e, params, packed, x_in, N, T, workspace, workspace_bytes, dropout_p, seed, outs, st);
  const bool has_temp = e.i_temp >= 0;
  MAED_CHECK_ARG(!has_temp || T <= cf.temp_frames, "train_forward: seqlen T=%d exceeds temp_embed frames %d", T, cf.temp_frames);
  MAED_CHECK_ARG(dropout_p >= 0.f && dropout_p < 1.f, "train_forward: dropout_p=%f", (double)dropout_p);
  const Net net = build_net(e);
  TrainWs w;
  carve(e, net, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "train_forward: workspace too small (%zu < %zu)", workspace_bytes,
                 w.total + 1024);
  Ctx c{e, net, w, params, (const uint8_t*)packed, nullptr, BT, N, T, st, 1.f, nullptr, x_in};

  // ---- backbone (GroupNorm unfused: conv outputs and statistics stay on the tape)
  MAED_CUDA_CHECK(cudaMemsetAsync(w.stats[0], 0, (size_t)BT * 64 * 8, st));
  MAED_PROPAGATE(stem_conv(x_in, BT, c.H(e.off_stem), 64LL * kStemKPad, kStemKPad, 3, w.convout[0], w.stats[0], st));
  MAED_PROPAGATE(gn_apply_maxpool_idx(w.convout[0], w.stats[0], c.P(e.i_stem_g), c.P(e.i_stem_g + 1), BT, 112, 112, 64, 1e-5f,
                                      w.out[0], w.out_plane[0], w.pool_idx, st));
  for (size_t l = 1; l < net.L.size(); ++l) MAED_PROPAGATE(conv_fwd(c, (int)l));
  const int last = (int)net.L.size() - 1;

  // ---- patch embedding
  const int ntok = 197, C = 768, heads = cf.num_heads;
  const int rows = BT * ntok;
  MAED_PROPAGATE(gemm_plain(c, w.out[last], w.out_plane[last], BT * 196, 1024, c.H(e.off_proj), 768LL * 1024, 768,
                            c.P(e.i_proj_b), ACT_NONE, nullptr, OUT_F32, w.tok, 0));
  MAED_PROPAGATE(embed_assemble(w.tok, c.P(e.i_cls), c.P(e.i_pos), has_temp ? c.P(e.i_temp) : nullptr, BT, T, ntok, C,
                                w.ste[0].x_in, st));

  // ---- STE blocks
  const float scale = 0.125f;
  const long long CC = (long long)C * C;
  for (int i = 0; i < cf.num_blocks; ++i) {
    const Engine::SteIdx& ix = e.blk[i];
    const Engine::SteOff& of = e.blk_off[i];
    SteTape& t = w.ste[i];
    float* x_out = (i + 1 < cf.num_blocks) ? w.ste[i + 1].x_in : w.x_final;
    if (cf.mode == MODE_TEMPORAL) {
      // vision_transformer.py:167-173: token mean of LN1(x) -> qkv -> attention across the T frames -> proj, broadcast over
      // the tokens.  Tape: pooled[:, :C] = the token mean, qkv / ao planes of BT rows, logits[:, :C] = proj output.
      MAED_PROPAGATE(layernorm_f32(t.x_in, C, c.P(ix.n1), c.P(ix.n1 + 1), rows, C, 1e-6f, w.dxs, st));
      MAED_PROPAGATE(token_mean(w.dxs, BT, ntok, C, t.pooled, C, 0, st));
      MAED_PROPAGATE(split_f32(t.pooled, w.small_p, w.small_plane, (long long)BT * C, st));
      MAED_PROPAGATE(gemm_plain(c, w.small_p, w.small_plane, BT, C, c.H(of.qkv), 3 * CC, 3 * C, c.P(ix.qkv_b), ACT_NONE, nullptr,
                                OUT_F16_SPLIT, t.qkv, w.qkv_plane));
      MAED_PROPAGATE(attn_temporal(t.qkv, w.qkv_plane, N, T, 1, heads, scale, nullptr, t.ao, w.ln_plane, st));
      MAED_PROPAGATE(gemm_plain(c, t.ao, w.ln_plane, BT, C, c.H(of.proj), CC, C, c.P(ix.proj_b), ACT_NONE, nullptr, OUT_F32,
                                t.logits, 0));
      MAED_CUDA_CHECK(cudaMemcpyAsync(t.x_mid, t.x_in, (size_t)rows * C * 4, cudaMemcpyDeviceToDevice, st));
      MAED_PROPAGATE(broadcast_add(t.x_mid, t.logits, BT, ntok, C, st));
    } else {
    MAED_PROPAGATE(layernorm_planes(t.x_in, C, c.P(ix.n1), c.P(ix.n1 + 1), rows, C, 1e-6f, t.ln1, w.ln_plane, st));
    MAED_PROPAGATE(gemm_plain(c, t.ln1, w.ln_plane, rows, C, c.H(of.qkv), 3 * CC, 3 * C, c.P(ix.qkv_b), ACT_NONE, nullptr,
                              OUT_F16_SPLIT, t.qkv, w.qkv_plane));
    float* lse_t = temporal_tc(T, ntok) ? t.lse_t : nullptr;     // only the tensor-core temporal kernel produces row statistics
    if (cf.mode == MODE_PARALLEL) {
      MAED_PROPAGATE(attn_temporal(t.qkv, w.qkv_plane, N, T, ntok, heads, scale, t.xt, nullptr, 0, st, lse_t));
      MAED_PROPAGATE(attn_spatial(t.qkv, w.qkv_plane, BT, ntok, heads, scale, 3, t.xs, nullptr, 0, st, t.lse_s));
      MAED_PROPAGATE(token_mean(t.xs, BT, ntok, C, t.pooled, 2 * C, 0, st));
      MAED_PROPAGATE(token_mean(t.xt, BT, ntok, C, t.pooled, 2 * C, C, st));
      MAED_PROPAGATE(split_f32(t.pooled, w.small_p, w.small_plane, (long long)BT * 2 * C, st));
      MAED_PROPAGATE(gemm_plain(c, w.small_p, w.small_plane, BT, 2 * C, c.H(of.ts), 4 * CC, 2 * C, c.P(ix.ts_b), ACT_NONE, nullptr,
                                OUT_F32, t.logits, 0));
      MAED_PROPAGATE(ts_blend(t.xs, t.xt, t.logits, BT, ntok, C, t.ao, w.ln_plane, st));
    } else if (cf.mode == MODE_SERIES) {
      MAED_PROPAGATE(attn_spatial(t.qkv, w.qkv_plane, BT, ntok, heads, scale, 3, nullptr, t.ao_s, w.ln_plane, st, t.lse_s));
      MAED_PROPAGATE(gemm_plain(c, t.ao_s, w.ln_plane, rows, C, c.H(of.qkv), 3 * CC, 3 * C, c.P(ix.qkv_b), ACT_NONE, nullptr,
                                OUT_F16_SPLIT, t.qkv2, w.qkv_plane));
      MAED_PROPAGATE(attn_temporal(t.qkv2, w.qkv_plane, N, T, ntok, heads, scale, nullptr, t.ao, w.ln_plane, st, lse_t));
    } else if (cf.mode == MODE_COUPLING) {                 // joint attention over the T * 197 tokens of a clip
      MAED_PROPAGATE(attn_generic(t.qkv, w.qkv_plane, N, T * ntok, heads, scale, ntok, T, nullptr, t.ao, w.ln_plane, st));
    } else {
      MAED_PROPAGATE(attn_spatial(t.qkv, w.qkv_plane, BT, ntok, heads, scale, 3, nullptr, t.ao, w.ln_plane, st, t.lse_s));
    }
    MAED_PROPAGATE(gemm_plain(c, t.ao, w.ln_plane, rows, C, c.H(of.proj), CC, C, c.P(ix.proj_b), ACT_NONE, t.x_in, OUT_F32,
                              t.x_mid, 0));
    }
    MAED_PROPAGATE(layernorm_planes(t.x_mid, C, c.P(ix.n2), c.P(ix.n2 + 1), rows, C, 1e-6f, t.ln2, w.ln_plane, st));
    MAED_PROPAGATE(gemm_plain(c, t.ln2, w.ln_plane, rows, C, c.H(of.fc1), 4 * CC, 4 * C, c.P(ix.fc1_b), ACT_NONE, nullptr,
                              OUT_F32, t.h_pre, 0));
    MAED_PROPAGATE(gelu_fwd_planes(t.h_pre, (long long)rows * 4 * C, t.hid, w.hid_plane, st));
    MAED_PROPAGATE(gemm_plain(c, t.hid, w.hid_plane, rows, 4 * C, c.H(of.fc2), 4 * CC, C, c.P(ix.fc2_b), ACT_NONE, t.x_mid,
                              OUT_F32, x_out, 0));
  }

  // ---- tail (always split precision; fp32 copies of every activation stay on the tape)
  MAED_PROPAGATE(layernorm_f32(w.x_final, (long long)ntok * C, c.P(e.i_norm), c.P(e.i_norm + 1), BT, C, 1e-6f, w.cls_ln, st));
  MAED_PROPAGATE(tail_gemm(c, w.cls_ln, C, e.off_pl, C, c.P(e.i_pl_b), ACT_TANH, w.feat_copy));
  return decoder_fwd(c, C, dropout_p, seed, outs);
}

// ------------------------------------------------------------------------------------------- backward
// spatial attention backward on the tensor cores (attention_bwd_sm100.cu): d_out -> fp16 hi/lo planes in pl_a (free at every
// call site: the planes of the previous linear's gradient have been consumed), d_qkv overwritten
// (lse: the forward's row statistics; O: the forward's output as fp32 or as planes — D = rowsum(dO o O) replaces two passes)
static int spatial_bwd(const Ctx& c, const __half* qkv, const float* d_out, float* dqkv, const float* lse, const float* o_f32,
                       const __half* o_hi, long long o_plane) {
  TrainWs& w = c.w;
  const int heads = c.e.cfg.num_heads;
  const long long rows = (long long)c.BT * 197, n = rows * heads * 64;
  MAED_PROPAGATE(attn_rowdot(d_out, o_f32, o_hi, o_plane, rows, heads, w.attn_D, c.st));
  MAED_PROPAGATE(split_f32(d_out, w.pl_a, w.pl_a_plane, n, c.st));
  return attn_spatial_bwd_tc(qkv, w.qkv_plane, w.pl_a, w.pl_a_plane, c.BT, 197, heads, 0.125f, 0, dqkv, c.st, lse, w.attn_D);
}

// temporal attention backward: tensor-core kernel when T is 4 / 8 / 16 / 32 and a token group is full (ntok >= 128 / T)
static int temporal_bwd(const Ctx& c, const __half* qkv, const float* d_out, int ntok, int accumulate, float* dqkv,
                        const float* lse, const float* o_f32, const __half* o_hi, long long o_plane) {
  TrainWs& w = c.w;
  const int heads = c.e.cfg.num_heads, T = c.T;
  if (temporal_tc(T, ntok)) {
    const long long rows = (long long)c.BT * ntok, n = rows * heads * 64;
    MAED_PROPAGATE(attn_rowdot(d_out, o_f32, o_hi, o_plane, rows, heads, w.attn_D, c.st));
    MAED_PROPAGATE(split_f32(d_out, w.pl_a, w.pl_a_plane, n, c.st));
    return attn_temporal_bwd_tc(qkv, w.qkv_plane, w.pl_a, w.pl_a_plane, c.N, T, ntok, heads, 0.125f, accumulate, dqkv, c.st, lse,
                                w.attn_D);
  }
  return attn_temporal_bwd(qkv, w.qkv_plane, d_out, c.N, T, ntok, heads, 0.125f, accumulate, dqkv, c.st);
}

// dW [Nw, Kw] = scale * dY^T X  for a linear layer; dY, X as planes [R, *] (dense rows)
static int linear_wgrad(const Ctx& c, const __half* dy, long long dy_plane, int Nw, const __half* x, long long x_plane, int Kw,
                        int R, int accumulate, float* dW) {
  // MN-major tcgen05 operands straight from the tape: no transposed copies of dY / X (gemm_splitk_sm100.cu)
  return gemm_wgrad_rows(dy, dy_plane, Nw, x, x_plane, Kw, Kw, Nw, Kw, R, 3, c.inv_ls, accumulate, c.w.slabs, dW, Kw, c.st);
}

// Backward of one conv + GroupNorm layer.  d_y: gradient w.r.t. the GN output (ReLU mask already applied), fp32
// [Mout, Cout].  Writes the parameter gradients; when d_in != nullptr also d_in = dgrad (+ d_in_add), fp32 [Min, Cin]
// (d_in may alias d_in_add).  tmp_f32: scratch of Mout*Cin floats (1x1 stride-2 data gradient only).
// relu_fused: d_y is the gradient behind the ReLU that follows this layer's GroupNorm and the mask has NOT been applied yet (the
// GroupNorm backward recomputes it; BatchNorm layers ignore the flag, their callers mask d_y themselves).  order: see groupnorm_bwd.
static int conv_layer_bwd(const Ctx& c, int l, const float* d_y, const float* d_in_add, float* d_in, float* tmp_f32,
                          bool relu_fused = false, int order = 0) {
  const ConvL& L = c.net.L[l];
  TrainWs& w = c.w;
  const int BT = c.BT;
  const long long Mo = L.Mout(BT);
  const int HWo = L.Hout * L.Hout;
  // ---- norm backward -> dconv planes [Mo, Cout] in pl_a; dgamma / dbeta
  if (c.net.bn) {
    MAED_PROPAGATE(bn_bwd(d_y, w.convout[l], w.bn_mean[l], w.bn_rstd[l], c.P(L.g_idx), Mo, L.Cout, c.inv_ls, w.bn_partial,
                          c.G(L.g_idx), c.G(L.g_idx + 1), w.pl_a, w.pl_a_plane, &c.e.bn_exchange, c.st));
  } else {
    MAED_PROPAGATE(groupnorm_bwd(d_y, w.convout[l], w.stats[l], c.P(L.g_idx), BT, HWo, L.Cout, 1e-5f, w.red, w.dgb, w.pl_a,
                                 w.pl_a_plane, c.st, relu_fused ? c.P(L.g_idx + 1) : nullptr, order));
    if (c.G(L.g_idx + 1) == c.G(L.g_idx) + L.Cout) {       // gamma and beta gradients are neighbours in the flat buffer: one pass
      MAED_PROPAGATE(colsum_f32(w.dgb, 2 * L.Cout, BT, 2 * L.Cout, c.inv_ls, 0, w.colsum_scratch, c.G(L.g_idx), c.st));
    } else {
      MAED_PROPAGATE(colsum_f32(w.dgb, 2 * L.Cout, BT, L.Cout, c.inv_ls, 0, w.colsum_scratch, c.G(L.g_idx), c.st));
      MAED_PROPAGATE(colsum_f32(w.dgb + L.Cout, 2 * L.Cout, BT, L.Cout, c.inv_ls, 0, w.colsum_scratch, c.G(L.g_idx + 1), c.st));
    }
  }
  // ---- weight gradient: dW_hat [Cout, kc] = dconv^T * im2col(x), then through the weight standardisation
  const int kc = (l == 0) ? kStemKPad : L.Kcols();
  const int kc_pad = (kc + 31) / 32 * 32;                  // split-K output width (multiple of 32); extra columns are zero
  const int pad = L.pad_top();
  const __half* xm;                                        // [Mo, kc] activation matrix of the wgrad
  long long xm_plane;
  if (l == 0) {
    MAED_PROPAGATE(im2col_stem(c.x_img, BT, 3, 224, 224, 7, 7, 2, pad, pad, 112, 112, kStemKPad, w.col, w.col_plane, c.st));
    xm = w.col; xm_plane = w.col_plane;
  } else if (L.k == 1 && L.stride == 1) {
    xm = w.out[L.in_layer]; xm_plane = w.out_plane[L.in_layer];
  } else if (L.stride == 1 && L.Cin % 64 == 0 && L.Cout % 64 == 0) {
    // stride-1 k x k: no im2col matrix, the split-K kernel reads the tap-shifted patches of x itself
    xm = nullptr; xm_plane = 0;
    MAED_PROPAGATE(gemm_wgrad_conv(w.pl_a, w.pl_a_plane, w.out[L.in_layer], w.out_plane[L.in_layer], BT, L.Hin, L.Hin, L.Cin, L.Cout,
                                   L.k, L.k, pad, 3, 1.0f, 0, w.slabs, w.wg, kc_pad, c.st));
  } else {
    MAED_PROPAGATE(im2col_nhwc(w.out[L.in_layer], w.out_plane[L.in_layer], BT, L.Hin, L.Hin, L.Cin, L.k, L.k, L.stride, pad, pad,
                               L.Hout, L.Hout, w.col, w.col_plane, c.st));
    xm = w.col; xm_plane = w.col_plane;
  }
  // dconv [Mo, Cout] and the activation matrix [Mo, kc] feed the tensor cores as MN-major operands (no transposed copies);
  // columns kc..kc_pad of dW_hat come out as zeros
  if (xm)
    MAED_PROPAGATE(gemm_wgrad_rows(w.pl_a, w.pl_a_plane, L.Cout, xm, xm_plane, kc, kc, L.Cout, kc_pad, (int)Mo, 3, 1.0f, 0,
                                   w.slabs, w.wg, kc_pad, c.st));
  if (c.net.bn)                                            // plain conv: only the [Cout][kh][kw][Cin] -> OIHW permute
    MAED_PROPAGATE(wgrad_permute(w.wg, kc_pad, L.Cout, L.Cin, L.k, L.k, c.inv_ls, c.G(L.w_idx), c.st));
  else
    MAED_PROPAGATE(wstd_bwd(w.wg, kc_pad, c.P(L.w_idx), L.Cout, L.Cin, L.k, L.k, 1e-5f, c.inv_ls, c.G(L.w_idx), c.st));
  if (!d_in) return MAED_OK;
  // ---- data gradient
  const long long tplane = (long long)L.Cout * L.Cin * L.k * L.k;
  if (L.k == 1 && L.stride == 1) {
    return gemm_plain(c, w.pl_a, w.pl_a_plane, (int)Mo, L.Cout, c.TH(L.tp_off), tplane, L.Cin, nullptr, ACT_NONE, d_in_add,
                      OUT_F32, d_in, 0);
  }
  if (L.k == 1) {                                          // 1x1 stride 2: dense GEMM, then scatter to the even positions
    MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, (int)Mo, L.Cout, c.TH(L.tp_off), tplane, L.Cin, nullptr, ACT_NONE, nullptr,
                              OUT_F32, tmp_f32, 0));
    return scatter_stride2_f32(tmp_f32, BT, L.Hout, L.Hout, L.Cin, L.Hin, L.Hin, d_in_add, d_in, c.st);
  }
  // k x k: stride-1 convolution of (dilated) dconv with the flipped, channel-transposed kernel
  const __half* src = w.pl_a;
  long long src_plane = w.pl_a_plane;
  if (L.stride == 2) {
    MAED_PROPAGATE(dilate2_planes(w.pl_a, w.pl_a_plane, BT, L.Hout, L.Hout, L.Cout, L.Hin, L.Hin, w.dil, w.dil_plane, c.st));
    src = w.dil; src_plane = w.dil_plane;
  }
  GemmArgs g;
  g.nsplit = 3;
  g.A = src; g.a_plane = src_plane; g.B = c.TH(L.tp_off); g.b_plane = tplane;
  g.M = (int)L.Min(BT); g.N = L.Cin; g.K = L.k * L.k * L.Cout;
  g.conv = 1; g.n_img = BT; g.H = L.Hin; g.W = L.Hin; g.Cin = L.Cout; g.KH = L.k; g.KW = L.k;
  g.pad_h = (L.k - 1) - pad; g.pad_w = (L.k - 1) - pad;
  g.residual = d_in_add; g.out_mode = OUT_F32; g.out = d_in; g.ldc = L.Cin;
  return launch_gemm(g, c.st);
}

// KTD decoder backward (reference ktd.py:69-88): parameter gradients of the heads, fc2, fc1 and d_feat [BT, C] (C = feature
// width: 768 'ste', 2048 'cnn').  The loss scale enters here: every activation gradient below carries it.
// iterative regressor backward: walks the 3 iterations in reverse; the state gradients (pose / shape / cam) flow both into the
// heads of the previous iteration and, through the concatenated input, into fc1; weight gradients accumulate over iterations
static int decoder_bwd_iterative(const Ctx& c, int C, const float* d_pose6d, const float* d_shape, const float* d_cam,
                                 float loss_scale, float dropout_p, float* d_feat) {
  const Engine& e = c.e;
  TrainWs& w = c.w;
  const int BT = c.BT, HD = e.cfg.hidden_dim, KI = C + 157;
  cudaStream_t st = c.st;
  float** sm = w.small;
  float* dP = sm[0]; float* dS = sm[1]; float* dC = sm[2]; float* d_h2 = sm[5]; float* d_h1 = sm[6];
  MAED_PROPAGATE(scale_f32(d_pose6d, loss_scale, (long long)BT * 144, dP, st));
  MAED_PROPAGATE(scale_f32(d_shape, loss_scale, (long long)BT * 10, dS, st));
  MAED_PROPAGATE(scale_f32(d_cam, loss_scale, (long long)BT * 3, dC, st));
  for (int it = 2; it >= 0; --it) {
    const float beta = it == 2 ? 0.f : 1.f;                // weight gradients: written by the first visited iteration, then added
    const int acc = it == 2 ? 0 : 1;
    // heads: state_{it+1} = head(h2_it) + state_it
    MAED_PROPAGATE(sgemm_f32(1, 0, 144, HD, BT, c.inv_ls, dP, 144, w.it_h2[it], HD, beta, c.G(e.i_decpose_w), HD, st));
    MAED_PROPAGATE(colsum_f32(dP, 144, BT, 144, c.inv_ls, acc, w.colsum_scratch, c.G(e.i_decpose_b), st));
    MAED_PROPAGATE(sgemm_f32(1, 0, 10, HD, BT, c.inv_ls, dS, 10, w.it_h2[it], HD, beta, c.G(e.i_shape_w), HD, st));
    MAED_PROPAGATE(colsum_f32(dS, 10, BT, 10, c.inv_ls, acc, w.colsum_scratch, c.G(e.i_shape_b), st));
    MAED_PROPAGATE(sgemm_f32(1, 0, 3, HD, BT, c.inv_ls, dC, 3, w.it_h2[it], HD, beta, c.G(e.i_cam_w), HD, st));
    MAED_PROPAGATE(colsum_f32(dC, 3, BT, 3, c.inv_ls, acc, w.colsum_scratch, c.G(e.i_cam_b), st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 144, 1.f, dP, 144, c.P(e.i_decpose_w), HD, 0.f, d_h2, HD, st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 10, 1.f, dS, 10, c.P(e.i_shape_w), HD, 1.f, d_h2, HD, st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 3, 1.f, dC, 3, c.P(e.i_cam_w), HD, 1.f, d_h2, HD, st));
    if (dropout_p > 0.f) MAED_PROPAGATE(dropout_bwd(d_h2, (long long)BT * HD, dropout_p, w.it_m2[it], st));
    // fc2
    MAED_PROPAGATE(sgemm_f32(1, 0, HD, HD, BT, c.inv_ls, d_h2, HD, w.it_h1[it], HD, beta, c.G(e.i_fc2_w), HD, st));
    MAED_PROPAGATE(colsum_f32(d_h2, HD, BT, HD, c.inv_ls, acc, w.colsum_scratch, c.G(e.i_fc2_b), st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, HD, 1.f, d_h2, HD, c.P(e.i_fc2_w), HD, 0.f, d_h1, HD, st));
    if (dropout_p > 0.f) MAED_PROPAGATE(dropout_bwd(d_h1, (long long)BT * HD, dropout_p, w.it_m1[it], st));
    // fc1 on xc = [feat | pose | shape | cam]
    MAED_PROPAGATE(sgemm_f32(1, 0, HD, KI, BT, c.inv_ls, d_h1, HD, w.it_xc[it], KI, beta, c.G(e.i_fc1_w), KI, st));
    MAED_PROPAGATE(colsum_f32(d_h1, HD, BT, HD, c.inv_ls, acc, w.colsum_scratch, c.G(e.i_fc1_b), st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, KI, HD, 1.f, d_h1, HD, c.P(e.i_fc1_w), KI, 0.f, w.it_dxc, KI, st));
    MAED_PROPAGATE(add_cols_f32(d_feat, C, w.it_dxc, KI, BT, C, acc, st));
    MAED_PROPAGATE(add_cols_f32(dP, 144, w.it_dxc + C, KI, BT, 144, 1, st));
    MAED_PROPAGATE(add_cols_f32(dS, 10, w.it_dxc + C + 144, KI, BT, 10, 1, st));
    MAED_PROPAGATE(add_cols_f32(dC, 3, w.it_dxc + C + 154, KI, BT, 3, 1, st));
  }
  return MAED_OK;
}

static int decoder_bwd(const Ctx& c, int C, const float* d_pose6d, const float* d_shape, const float* d_cam, float loss_scale,
                       float dropout_p, float* d_feat) {
  const Engine& e = c.e;
  if (e.cfg.decoder == DEC_ITERATIVE) return decoder_bwd_iterative(c, C, d_pose6d, d_shape, d_cam, loss_scale, dropout_p, d_feat);
  TrainWs& w = c.w;
  const int BT = c.BT, HD = e.cfg.hidden_dim;
  cudaStream_t st = c.st;
  float** sm = w.small;
  // loss scale enters here: every activation gradient below carries it
  float* dpose = sm[0]; float* dshape = sm[1]; float* dcam = sm[2];
  MAED_PROPAGATE(scale_f32(d_pose6d, loss_scale, (long long)BT * 144, dpose, st));
  MAED_PROPAGATE(scale_f32(d_shape, loss_scale, (long long)BT * 10, dshape, st));
  MAED_PROPAGATE(scale_f32(d_cam, loss_scale, (long long)BT * 3, dcam, st));
  float* g_total = sm[3];                       // [BT, 144]
  float* d_base = sm[4];                        // [BT, 192]
  const float* w_anc = (const float*)(c.pk + e.off_ktd_anc);
  MAED_PROPAGATE(ktd_tree_bwd(dpose, dshape, dcam, w_anc, BT, g_total, d_base, 192, st));
  MAED_PROPAGATE(ktd_anc_wgrad(g_total, w.pose_copy, BT, c.inv_ls, w.anc_grad, st));
  // heads: base[:, 0:144] = h2 Wx^T + b (joint_regs.*.weight[:, :HD]), [144:154] decshape, [154:157] deccam
  float* d_h2 = sm[5];
  {
    // per-joint weight / bias gradients: joint_regs.j.weight = [6, HD + 6k]: first HD columns from dbase^T h2, last 6k from anc_grad
    int aoff = 0;
    for (int j = 0; j < 24; ++j) {
      const int k = kAncCnt[j];
      float* gw = c.G(e.i_joint0 + 2 * j);
      const int ldw = HD + 6 * k;
      MAED_PROPAGATE(sgemm_f32(1, 0, 6, HD, BT, c.inv_ls, d_base + 6 * j, 192, w.h2, HD, 0.f, gw, ldw, st));
      if (k > 0)
        MAED_CUDA_CHECK(cudaMemcpy2DAsync(gw + HD, (size_t)ldw * 4, w.anc_grad + aoff, (size_t)6 * k * 4, (size_t)6 * k * 4, 6,
                                          cudaMemcpyDeviceToDevice, st));
      MAED_PROPAGATE(colsum_f32(d_base + 6 * j, 192, BT, 6, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_joint0 + 2 * j + 1), st));
      aoff += 36 * k;
    }
    MAED_PROPAGATE(sgemm_f32(1, 0, 10, HD, BT, c.inv_ls, d_base + 144, 192, w.h2, HD, 0.f, c.G(e.i_shape_w), HD, st));
    MAED_PROPAGATE(colsum_f32(d_base + 144, 192, BT, 10, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_shape_b), st));
    MAED_PROPAGATE(sgemm_f32(1, 0, 3, HD, BT, c.inv_ls, d_base + 154, 192, w.h2, HD, 0.f, c.G(e.i_cam_w), HD, st));
    MAED_PROPAGATE(colsum_f32(d_base + 154, 192, BT, 3, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_cam_b), st));
    // d_h2 = dbase[:, :144] Wx + dbase[:, 144:154] Wshape + dbase[:, 154:157] Wcam
    const float* wx = (const float*)(c.pk + e.off_ktd_wx);            // [144, HD] fp32 (engine pack)
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 144, 1.f, d_base, 192, wx, HD, 0.f, d_h2, HD, st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 10, 1.f, d_base + 144, 192, c.P(e.i_shape_w), HD, 1.f, d_h2, HD, st));
    MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, 3, 1.f, d_base + 154, 192, c.P(e.i_cam_w), HD, 1.f, d_h2, HD, st));
  }
  if (dropout_p > 0.f) MAED_PROPAGATE(dropout_bwd(d_h2, (long long)BT * HD, dropout_p, w.mask2, st));
  // fc2: h2 = h1 W2^T + b2
  float* d_h1 = sm[6];
  MAED_PROPAGATE(sgemm_f32(1, 0, HD, HD, BT, c.inv_ls, d_h2, HD, w.h1, HD, 0.f, c.G(e.i_fc2_w), HD, st));
  MAED_PROPAGATE(colsum_f32(d_h2, HD, BT, HD, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_fc2_b), st));
  MAED_PROPAGATE(sgemm_f32(0, 0, BT, HD, HD, 1.f, d_h2, HD, c.P(e.i_fc2_w), HD, 0.f, d_h1, HD, st));
  if (dropout_p > 0.f) MAED_PROPAGATE(dropout_bwd(d_h1, (long long)BT * HD, dropout_p, w.mask1, st));
  // fc1: h1 = feat W1^T + b1
  MAED_PROPAGATE(sgemm_f32(1, 0, HD, C, BT, c.inv_ls, d_h1, HD, w.feat_copy, C, 0.f, c.G(e.i_fc1_w), C, st));
  MAED_PROPAGATE(colsum_f32(d_h1, HD, BT, HD, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_fc1_b), st));
  MAED_PROPAGATE(sgemm_f32(0, 0, BT, C, HD, 1.f, d_h1, HD, c.P(e.i_fc1_w), C, 0.f, d_feat, C, st));
  return MAED_OK;
}

static int cnn_train_backward(const Engine& e, const void* const* params, const void* packed, const void* tpack, const float* x_in,
                              int N, int T, void* workspace, size_t workspace_bytes, const float* d_pose6d, const float* d_shape,
                              const float* d_cam, float loss_scale, float dropout_p, float* const* grads, cudaStream_t st) {
  const int BT = N * T;
  const Net net = build_cnn_net(e);
  TrainWs w;
  carve_cnn(e, net, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "train_backward(cnn): workspace too small");
  const uint8_t* tp = (const uint8_t*)(((uintptr_t)tpack + 1023) & ~(uintptr_t)1023);
  Ctx c{e, net, w, params, (const uint8_t*)packed, tp, BT, N, T, st, 1.0f / loss_scale, grads, x_in};
  const int F = e.feat_dim();
  float* d_feat = w.small[7];
  MAED_PROPAGATE(decoder_bwd(c, F, d_pose6d, d_shape, d_cam, loss_scale, dropout_p, d_feat));
  // average pool: every position of the 7x7 map receives d_feat / 49
  MAED_PROPAGATE(avgpool_bwd(d_feat, BT, 49, F, w.fc, st));
  float* bufs[4] = {w.fc, w.fa, w.fb, w.fd};
  for (int b = (int)net.B.size() - 1; b >= 0; --b) {
    const BlockL& bl = net.B[b];
    float* g = bufs[0]; float* d_short = bufs[1]; float* d_t2 = bufs[2]; float* d_t1 = bufs[3];
    const ConvL& L3 = net.L[bl.c3];
    const ConvL& L2 = net.L[bl.c2];
    const ConvL& L1 = net.L[bl.c1];
    MAED_PROPAGATE(relu_mask_f32(g, w.out[bl.c3], L3.Mout(BT) * L3.Cout, st, 1));
    const float* shortcut_grad = g;
    float* d_xin = g;
    if (bl.ds >= 0) {
      MAED_PROPAGATE(conv_layer_bwd(c, bl.ds, g, nullptr, d_short, d_t2));
      shortcut_grad = d_short;
      d_xin = d_short;
    }
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c3, g, nullptr, d_t2, nullptr));
    MAED_PROPAGATE(relu_mask_f32(d_t2, w.out[bl.c2], L2.Mout(BT) * L2.Cout, st, 1));
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c2, d_t2, nullptr, d_t1, nullptr));
    MAED_PROPAGATE(relu_mask_f32(d_t1, w.out[bl.c1], L1.Mout(BT) * L1.Cout, st, 1));
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c1, d_t1, shortcut_grad, d_xin, nullptr));
    if (bl.ds >= 0) std::swap(bufs[0], bufs[1]);
  }
  // stem: max-pool (gather by arg-max) -> ReLU mask recomputed from the conv output -> BatchNorm -> conv1 (weights only)
  float* d_pool = bufs[0];
  float* d_y = bufs[1];
  const ConvL& L0 = net.L[0];
  MAED_PROPAGATE(maxpool3x3s2_bwd(d_pool, w.pool_idx, BT, 112, 112, 64, d_y, st));
  MAED_PROPAGATE(bn_relu_mask(d_y, w.convout[0], w.bn_mean[0], w.bn_rstd[0], c.P(L0.g_idx), c.P(L0.g_idx + 1),
                              (long long)BT * 12544, 64, st));
  return conv_layer_bwd(c, 0, d_y, nullptr, nullptr, nullptr);
}

int train_backward(const Engine* ep, const void* const* params, const void* packed, const void* tpack, const float* x_in, int N,
                   int T, void* workspace, size_t workspace_bytes, const float* d_pose6d, const float* d_shape,
                   const float* d_cam, float loss_scale, float dropout_p, float* const* grads, cudaStream_t st) {
  MAED_CHECK_ARG(ep, "train_backward: null engine");
  MAED_PROPAGATE(check_train_cfg(*ep));
  MAED_CHECK_ARG(ep && params && packed && tpack && x_in && workspace && d_pose6d && d_shape && d_cam && grads,
                 "train_backward: null argument");
  MAED_CHECK_ARG(loss_scale > 0.f, "train_backward: loss_scale must be positive");
  const Engine& e = *ep;
  if (e.cfg.encoder == ENC_CNN)
    return cnn_train_backward(e, params, packed, tpack, x_in, N, T, workspace, workspace_bytes, d_pose6d, d_shape, d_cam, loss_scale,
                              dropout_p, grads, st);
  const EngineConfig& cf = e.cfg;
  const int BT = N * T;
  const Net net = build_net(e);
  TrainWs w;
  carve(e, net, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "train_backward: workspace too small");
  const uint8_t* tp = (const uint8_t*)(((uintptr_t)tpack + 1023) & ~(uintptr_t)1023);
  Ctx c{e, net, w, params, (const uint8_t*)packed, tp, BT, N, T, st, 1.0f / loss_scale, grads, x_in};
  const int ntok = 197, C = 768, heads = cf.num_heads, HD = cf.hidden_dim;
  const int rows = BT * ntok;
  const float scale = 0.125f;
  const long long CC = (long long)C * C;
  const bool has_temp = e.i_temp >= 0;
  float** sm = w.small;

  // ================================================================================ tail (fp32 CUDA cores)
  float* d_feat = sm[7];
  MAED_PROPAGATE(decoder_bwd(c, C, d_pose6d, d_shape, d_cam, loss_scale, dropout_p, d_feat));
  // pre_logits: feat = tanh(cls_ln Wp^T + bp)
  float* d_pre = sm[0];
  MAED_PROPAGATE(tanh_bwd(d_feat, w.feat_copy, (long long)BT * C, d_pre, st));
  MAED_PROPAGATE(sgemm_f32(1, 0, C, C, BT, c.inv_ls, d_pre, C, w.cls_ln, C, 0.f, c.G(e.i_pl_w), C, st));
  MAED_PROPAGATE(colsum_f32(d_pre, C, BT, C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_pl_b), st));
  float* d_cls = sm[1];
  MAED_PROPAGATE(sgemm_f32(0, 0, BT, C, C, 1.f, d_pre, C, c.P(e.i_pl_w), C, 0.f, d_cls, C, st));
  // final LayerNorm acts on the cls row of every frame; all other rows of dL/dx_final are zero
  float* dx = w.fa;                              // residual-stream gradient [rows, C]
  float* dx2 = w.fb;
  MAED_CUDA_CHECK(cudaMemsetAsync(dx, 0, (size_t)rows * C * 4, st));
  MAED_PROPAGATE(layernorm_bwd(d_cls, C, w.x_final, (long long)ntok * C, c.P(e.i_norm), BT, C, 1e-6f, nullptr, dx,
                               (long long)ntok * C, w.ln_partial, st));
  const int lnr = ln_bwd_partial_rows();
  MAED_PROPAGATE(colsum_f32(w.ln_partial, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_norm), st));
  MAED_PROPAGATE(colsum_f32(w.ln_partial + C, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_norm + 1), st));

  // ====================================================================================== STE blocks
  for (int i = cf.num_blocks - 1; i >= 0; --i) {
    const Engine::SteIdx& ix = e.blk[i];
    const SteTape& t = w.ste[i];
    const Net::SteT& tt = net.ste[i];
    // ---- MLP: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))));  dx = dL/dx_out
    MAED_PROPAGATE(split_f32(dx, w.pl_a, w.pl_a_plane, (long long)rows * C, st));                       // d_y2 planes
    MAED_PROPAGATE(colsum_f32(dx, C, rows, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.fc2_b), st));
    MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, C, t.hid, w.hid_plane, 4 * C, rows, 0, c.G(ix.fc2_w)));
    MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, rows, C, c.TH(tt.fc2), 4 * CC, 4 * C, nullptr, ACT_NONE, nullptr, OUT_F32,
                              w.big, 0));                                                                 // d_hid
    MAED_PROPAGATE(gelu_bwd(w.big, t.h_pre, (long long)rows * 4 * C, w.pl_a, w.pl_a_plane, st));         // d_pre planes
    MAED_PROPAGATE(colsum_planes(w.pl_a, w.pl_a_plane, 4 * C, rows, 4 * C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.fc1_b), st));
    MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, 4 * C, t.ln2, w.ln_plane, C, rows, 0, c.G(ix.fc1_w)));
    MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, rows, 4 * C, c.TH(tt.fc1), 4 * CC, C, nullptr, ACT_NONE, nullptr, OUT_F32,
                              dx2, 0));                                                                   // d_ln2
    MAED_PROPAGATE(layernorm_bwd(dx2, C, t.x_mid, C, c.P(ix.n2), rows, C, 1e-6f, dx, dx, C, w.ln_partial, st));   // dx = d_xmid
    MAED_PROPAGATE(colsum_f32(w.ln_partial, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n2), st));
    MAED_PROPAGATE(colsum_f32(w.ln_partial + C, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n2 + 1), st));
    if (cf.mode == MODE_TEMPORAL) {
      // x_mid[bt, i] = x_in[bt, i] + v[bt],  v = proj(attn_T(qkv(mean_i LN1(x_in))))   — all BT-row problems, fp32 CUDA cores
      float* d_v = sm[0]; float* ao32 = sm[1]; float* d_ao = sm[2]; float* d_mean = sm[3]; float* dqkv_t = w.big;   // [BT, 3C]
      MAED_PROPAGATE(token_sum(dx, BT, ntok, C, d_v, st));
      MAED_PROPAGATE(planes_to_f32(t.ao, w.ln_plane, (long long)BT * C, ao32, st));
      MAED_PROPAGATE(sgemm_f32(1, 0, C, C, BT, c.inv_ls, d_v, C, ao32, C, 0.f, c.G(ix.proj_w), C, st));
      MAED_PROPAGATE(colsum_f32(d_v, C, BT, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.proj_b), st));
      MAED_PROPAGATE(sgemm_f32(0, 0, BT, C, C, 1.f, d_v, C, c.P(ix.proj_w), C, 0.f, d_ao, C, st));
      MAED_PROPAGATE(attn_temporal_bwd(t.qkv, w.qkv_plane, d_ao, N, T, 1, heads, scale, 0, dqkv_t, st));
      MAED_PROPAGATE(sgemm_f32(1, 0, 3 * C, C, BT, c.inv_ls, dqkv_t, 3 * C, t.pooled, C, 0.f, c.G(ix.qkv_w), C, st));
      MAED_PROPAGATE(colsum_f32(dqkv_t, 3 * C, BT, 3 * C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.qkv_b), st));
      MAED_PROPAGATE(sgemm_f32(0, 0, BT, C, 3 * C, 1.f / (float)ntok, dqkv_t, 3 * C, c.P(ix.qkv_w), C, 0.f, d_mean, C, st));
      MAED_CUDA_CHECK(cudaMemsetAsync(dx2, 0, (size_t)rows * C * 4, st));
      MAED_PROPAGATE(broadcast_add(dx2, d_mean, BT, ntok, C, st));                                           // d LN1 output
      MAED_PROPAGATE(layernorm_bwd(dx2, C, t.x_in, C, c.P(ix.n1), rows, C, 1e-6f, dx, dx, C, w.ln_partial, st));
      MAED_PROPAGATE(colsum_f32(w.ln_partial, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n1), st));
      MAED_PROPAGATE(colsum_f32(w.ln_partial + C, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n1 + 1), st));
      continue;
    }
    // ---- attention: x_mid = x_in + proj(ao)
    MAED_PROPAGATE(split_f32(dx, w.pl_a, w.pl_a_plane, (long long)rows * C, st));
    MAED_PROPAGATE(colsum_f32(dx, C, rows, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.proj_b), st));
    MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, C, t.ao, w.ln_plane, C, rows, 0, c.G(ix.proj_w)));
    float* d_ao = dx2;
    MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, rows, C, c.TH(tt.proj), CC, C, nullptr, ACT_NONE, nullptr, OUT_F32, d_ao, 0));
    float* dqkv = w.big;                           // [rows, 3C] fp32
    if (cf.mode == MODE_PARALLEL) {
      float* d_logits = sm[2]; float* d_pool = sm[3];
      MAED_PROPAGATE(blend_bwd(d_ao, t.xs, t.xt, t.logits, BT, ntok, C, d_logits, w.dxs, w.dxt, st));
      MAED_PROPAGATE(sgemm_f32(1, 0, 2 * C, 2 * C, BT, c.inv_ls, d_logits, 2 * C, t.pooled, 2 * C, 0.f, c.G(ix.ts_w), 2 * C, st));
      MAED_PROPAGATE(colsum_f32(d_logits, 2 * C, BT, 2 * C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.ts_b), st));
      // d_pool = d_logits W_ts on the tensor cores (the fp32 CUDA-core GEMM took 0.4 ms per block for these 128 rows)
      MAED_PROPAGATE(split_f32(d_logits, w.small_p, w.small_plane, (long long)BT * 2 * C, st));
      MAED_PROPAGATE(gemm_plain(c, w.small_p, w.small_plane, BT, 2 * C, c.TH(tt.ts), 4 * CC, 2 * C, nullptr, ACT_NONE, nullptr, OUT_F32,
                                d_pool, 0));
      MAED_PROPAGATE(blend_bwd_pool(d_pool, BT, ntok, C, w.dxs, w.dxt, st));
      MAED_PROPAGATE(spatial_bwd(c, t.qkv, w.dxs, dqkv, t.lse_s, t.xs, nullptr, 0));
      MAED_PROPAGATE(temporal_bwd(c, t.qkv, w.dxt, ntok, 1, dqkv, t.lse_t, t.xt, nullptr, 0));
    } else if (cf.mode == MODE_SERIES) {
      // ao = temporal(qkv2), qkv2 = qkv(ao_s), ao_s = spatial(qkv), qkv = qkv(ln1): the qkv weights are used twice
      MAED_PROPAGATE(temporal_bwd(c, t.qkv2, d_ao, ntok, 0, dqkv, t.lse_t, nullptr, t.ao, w.ln_plane));
      MAED_PROPAGATE(split_f32(dqkv, w.pl_a, w.pl_a_plane, (long long)rows * 3 * C, st));
      MAED_PROPAGATE(colsum_f32(dqkv, 3 * C, rows, 3 * C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.qkv_b), st));
      MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, 3 * C, t.ao_s, w.ln_plane, C, rows, 0, c.G(ix.qkv_w)));
      MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, rows, 3 * C, c.TH(tt.qkv), 3 * CC, C, nullptr, ACT_NONE, nullptr, OUT_F32,
                                w.dxs, 0));                                                               // d_ao_s
      MAED_PROPAGATE(spatial_bwd(c, t.qkv, w.dxs, dqkv, t.lse_s, nullptr, t.ao_s, w.ln_plane));
    } else if (cf.mode == MODE_COUPLING) {
      MAED_PROPAGATE(attn_generic_bwd(t.qkv, w.qkv_plane, d_ao, N, T * ntok, heads, scale, 0, dqkv, w.dxt, st));   // dxt: statistics
    } else {
      MAED_PROPAGATE(spatial_bwd(c, t.qkv, d_ao, dqkv, t.lse_s, nullptr, t.ao, w.ln_plane));
    }
    const int acc = cf.mode == MODE_SERIES ? 1 : 0;
    MAED_PROPAGATE(split_f32(dqkv, w.pl_a, w.pl_a_plane, (long long)rows * 3 * C, st));
    MAED_PROPAGATE(colsum_f32(dqkv, 3 * C, rows, 3 * C, c.inv_ls, acc, w.colsum_scratch, c.G(ix.qkv_b), st));
    MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, 3 * C, t.ln1, w.ln_plane, C, rows, acc, c.G(ix.qkv_w)));
    MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, rows, 3 * C, c.TH(tt.qkv), 3 * CC, C, nullptr, ACT_NONE, nullptr, OUT_F32,
                              dx2, 0));                                                                   // d_ln1
    MAED_PROPAGATE(layernorm_bwd(dx2, C, t.x_in, C, c.P(ix.n1), rows, C, 1e-6f, dx, dx, C, w.ln_partial, st));    // dx = d_xin
    MAED_PROPAGATE(colsum_f32(w.ln_partial, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n1), st));
    MAED_PROPAGATE(colsum_f32(w.ln_partial + C, 2 * C, lnr, C, c.inv_ls, 0, w.colsum_scratch, c.G(ix.n1 + 1), st));
    // gradients of block i and of everything behind it in the table (later blocks, final norm, pre_logits, decoder) are final
    if (e.progress.fn) {
      const int hi = (i + 1 < cf.num_blocks) ? e.blk[i + 1].n1 : (int)e.names.size();
      MAED_CHECK_ARG(e.progress.fn(e.progress.user, ix.n1, hi) == 0, "train_backward: the progress callback failed");
    }
  }

  // ================================================================================== patch embedding
  // x0[bt, 0] = cls + pos[0] (+ temp[t]);  x0[bt, 1+i] = tok[bt, i] + pos[1+i] (+ temp[t])
  MAED_PROPAGATE(colsum_f32(dx, (long long)ntok * C, BT, ntok * C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_pos), st));
  MAED_CUDA_CHECK(cudaMemcpyAsync(c.G(e.i_cls), c.G(e.i_pos), (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
  if (has_temp) {
    float* tsum = sm[0];                                   // [BT, C] sum over the tokens of each frame
    MAED_PROPAGATE(token_sum(dx, BT, ntok, C, tsum, st));
    MAED_PROPAGATE(colsum_f32(tsum, (long long)T * C, N, T * C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_temp), st));
    if (T < cf.temp_frames)
      MAED_CUDA_CHECK(cudaMemsetAsync(c.G(e.i_temp) + (size_t)T * C, 0, (size_t)(cf.temp_frames - T) * C * 4, st));
  }
  // d_tok [BT*196, C]: rows 1..196 of every frame
  const int trow = BT * 196;
  // compact the token rows (drops the cls row of every frame)
  MAED_CUDA_CHECK(cudaMemcpy2DAsync(dx2, (size_t)196 * C * 4, dx + C, (size_t)ntok * C * 4, (size_t)196 * C * 4, BT,
                                    cudaMemcpyDeviceToDevice, st));
  MAED_PROPAGATE(split_f32(dx2, w.pl_a, w.pl_a_plane, (long long)trow * C, st));
  MAED_PROPAGATE(colsum_f32(dx2, C, trow, C, c.inv_ls, 0, w.colsum_scratch, c.G(e.i_proj_b), st));
  const int last = (int)net.L.size() - 1;
  MAED_PROPAGATE(linear_wgrad(c, w.pl_a, w.pl_a_plane, C, w.out[last], w.out_plane[last], 1024, trow, 0, c.G(e.i_proj_w)));
  float* d_out = w.fc;                                     // gradient w.r.t. the stage-2 output [BT*196, 1024]
  MAED_PROPAGATE(gemm_plain(c, w.pl_a, w.pl_a_plane, trow, C, c.TH(net.tp_proj), 768LL * 1024, 1024, nullptr, ACT_NONE, nullptr,
                            OUT_F32, d_out, 0));

  // ========================================================================================= backbone
  // free fp32 buffers: fa, fb, fd (+ fc holding d_out).  Roles per bottleneck: g (= d_out), d_short, d_t2, d_t1.
  float* bufs[4] = {w.fc, w.fa, w.fb, w.fd};
  for (int b = (int)net.B.size() - 1; b >= 0; --b) {
    const BlockL& bl = net.B[b];
    float* g = bufs[0]; float* d_short = bufs[1]; float* d_t2 = bufs[2]; float* d_t1 = bufs[3];
    const ConvL& L3 = net.L[bl.c3];
    const ConvL& L2 = net.L[bl.c2];
    const ConvL& L1 = net.L[bl.c1];
    // ReLU behind the block sum: g is read three times (shortcut, c3, downsample), so this mask is materialised — walking
    // downwards, g was written upwards by the GEMM before.  The ReLUs behind c1 / c2 are folded into their GroupNorm backward.
    MAED_PROPAGATE(relu_mask_f32(g, w.out[bl.c3], L3.Mout(BT) * L3.Cout, st, 1));
    const float* shortcut_grad = g;
    float* d_xin = g;
    if (bl.ds >= 0) {
      MAED_PROPAGATE(conv_layer_bwd(c, bl.ds, g, nullptr, d_short, d_t2, false, 2));
      shortcut_grad = d_short;
      d_xin = d_short;
    }
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c3, g, nullptr, d_t2, nullptr, false, bl.ds >= 0 ? 0 : 2));
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c2, d_t2, nullptr, d_t1, nullptr, true, 1));
    MAED_PROPAGATE(conv_layer_bwd(c, bl.c1, d_t1, shortcut_grad, d_xin, nullptr, true, 1));
    if (bl.ds >= 0) std::swap(bufs[0], bufs[1]);            // the block-input gradient now lives in d_short's buffer
  }
  // stem: max-pool + ReLU + GroupNorm + conv (no data gradient: the input frames need none)
  float* d_pool = bufs[0];
  float* d_y = bufs[1];
  MAED_PROPAGATE(maxpool_gn_relu_bwd(d_pool, w.pool_idx, w.convout[0], w.stats[0], c.P(e.i_stem_g), c.P(e.i_stem_g + 1), BT, 112,
                                     112, 64, 1e-5f, d_y, st));
  MAED_PROPAGATE(conv_layer_bwd(c, 0, d_y, nullptr, nullptr, nullptr));
  if (e.progress.fn && cf.num_blocks > 0)
    MAED_CHECK_ARG(e.progress.fn(e.progress.user, 0, e.blk[0].n1) == 0, "train_backward: the progress callback failed");
  return MAED_OK;
}

}  // namespace maed
