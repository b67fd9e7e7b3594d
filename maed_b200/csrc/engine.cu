// The MAED forward engine: one C++ object per model configuration that owns the launch plan of
// `MAED(encoder='ste', ...).forward` (reference lib/models/maed.py:52-66 -> vision_transformer.py:388-412 ->
// resnetv2.py:337-348 -> ktd.py:69-124 / spin.py:76-110).  Python hands over raw device pointers; every kernel
// is launched from here on the caller's stream, so a whole forward costs one ctypes call and can be captured in
// a CUDA graph.
#include "engine_internal.h"
#include "fork_join.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <string>
#include <vector>

#include "gemm_host.h"
#include "gemm_sm100.cuh"
#include "kernels.h"

namespace maed {

// ------------------------------------------------------------------------------------------ helpers
__global__ void planes_to_f32_kernel(const __half* __restrict__ hi, long long plane, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __half2float(hi[i]) + (plane ? __half2float(hi[i + plane]) : 0.f);
}
int planes_to_f32(const __half* hi, long long plane, long long n, float* out, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  planes_to_f32_kernel<<<blocks, 256, 0, st>>>(hi, plane, n, out);
  count_launch();
  MAED_CUDA_CHECK(cudaGetLastError());
  return MAED_OK;
}
__global__ void broadcast_row_kernel(const float* __restrict__ src, int C, long long total, float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i % C];
}
int broadcast_row(const float* src, int C, int R, float* dst, cudaStream_t st) {
  const long long total = (long long)R * C;
  broadcast_row_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(src, C, total, dst);
  count_launch();
  MAED_CUDA_CHECK(cudaGetLastError());
  return MAED_OK;
}

static int add_param(Engine& e, const std::string& name, long long numel) {
  e.names.push_back(name);
  e.numels.push_back(numel);
  return (int)e.names.size() - 1;
}
static size_t align_up(size_t v, size_t a = 1024) { return (v + a - 1) / a * a; }

static void build_tables(Engine& e) {
  const EngineConfig& c = e.cfg;
  const int C = 768, H = c.num_heads;
  (void)H;
  const std::string enc = "encoder.";
  const bool cnn = c.encoder == ENC_CNN;
  e.i_cls = e.i_pos = e.i_temp = e.i_stem_w = e.i_stem_g = e.i_proj_w = e.i_proj_b = e.i_norm = e.i_pl_w = e.i_pl_b = -1;
  if (cnn) {
    cnn_add_params(e, add_param);
  } else {
    e.i_cls = add_param(e, enc + "cls_token", C);
    e.i_pos = add_param(e, enc + "pos_embed", 197 * C);
    const bool has_temp = (c.mode == MODE_PARALLEL || c.mode == MODE_SERIES || c.mode == MODE_COUPLING);
    e.i_temp = has_temp ? add_param(e, enc + "temp_embed", (long long)c.temp_frames * C) : -1;
    const std::string bbp = enc + "patch_embed.backbone.";
    e.i_stem_w = add_param(e, bbp + "stem.conv.weight", 64 * 3 * 49);
    e.i_stem_g = add_param(e, bbp + "stem.norm.weight", 64);
    add_param(e, bbp + "stem.norm.bias", 64);
    int prev = 64;
    for (int s = 0; s < 3; ++s) {
      const int out = kStageOut[s], mid = out / 4;
      for (int b = 0; b < kStageDepth[s]; ++b) {
        const std::string p = bbp + "stages." + std::to_string(s) + ".blocks." + std::to_string(b) + ".";
        Engine::BlockIdx bi;
        memset(&bi, 0xff, sizeof(bi));
        if (b == 0) {
          bi.ds_w = add_param(e, p + "downsample.conv.weight", (long long)out * prev);
          bi.ds_g = add_param(e, p + "downsample.norm.weight", out);
          add_param(e, p + "downsample.norm.bias", out);
        }
        bi.c1_w = add_param(e, p + "conv1.weight", (long long)mid * prev);
        bi.c1_g = add_param(e, p + "norm1.weight", mid);
        add_param(e, p + "norm1.bias", mid);
        bi.c2_w = add_param(e, p + "conv2.weight", (long long)mid * mid * 9);
        bi.c2_g = add_param(e, p + "norm2.weight", mid);
        add_param(e, p + "norm2.bias", mid);
        bi.c3_w = add_param(e, p + "conv3.weight", (long long)out * mid);
        bi.c3_g = add_param(e, p + "norm3.weight", out);
        add_param(e, p + "norm3.bias", out);
        e.bb.push_back(bi);
        prev = out;
      }
    }
    e.i_proj_w = add_param(e, enc + "patch_embed.proj.weight", 768LL * 1024);
    e.i_proj_b = add_param(e, enc + "patch_embed.proj.bias", 768);
    for (int i = 0; i < c.num_blocks; ++i) {
      const std::string p = enc + "blocks." + std::to_string(i) + ".";
      Engine::SteIdx si;
      memset(&si, 0xff, sizeof(si));
      si.n1 = add_param(e, p + "norm1.weight", C);
      add_param(e, p + "norm1.bias", C);
      si.qkv_w = add_param(e, p + "attn.qkv.weight", 3LL * C * C);
      si.qkv_b = add_param(e, p + "attn.qkv.bias", 3 * C);
      if (c.mode == MODE_PARALLEL) {
        si.ts_w = add_param(e, p + "attn.ts_attn.weight", 4LL * C * C);
        si.ts_b = add_param(e, p + "attn.ts_attn.bias", 2 * C);
      }
      si.proj_w = add_param(e, p + "attn.proj.weight", (long long)C * C);
      si.proj_b = add_param(e, p + "attn.proj.bias", C);
      si.n2 = add_param(e, p + "norm2.weight", C);
      add_param(e, p + "norm2.bias", C);
      si.fc1_w = add_param(e, p + "mlp.fc1.weight", 4LL * C * C);
      si.fc1_b = add_param(e, p + "mlp.fc1.bias", 4 * C);
      si.fc2_w = add_param(e, p + "mlp.fc2.weight", 4LL * C * C);
      si.fc2_b = add_param(e, p + "mlp.fc2.bias", C);
      e.blk.push_back(si);
    }
    e.i_norm = add_param(e, enc + "norm.weight", C);
    add_param(e, enc + "norm.bias", C);
    e.i_pl_w = add_param(e, enc + "pre_logits.fc.weight", (long long)C * C);
    e.i_pl_b = add_param(e, enc + "pre_logits.fc.bias", C);
  }  // !cnn
  const int HD = c.hidden_dim;
  const int F = e.feat_dim();
  const std::string dec = "decoder.";
  const int fc1_in = c.decoder == DEC_KTD ? F : F + 144 + 10 + 3;
  e.i_fc1_w = add_param(e, dec + "fc1.weight", (long long)HD * fc1_in);
  e.i_fc1_b = add_param(e, dec + "fc1.bias", HD);
  e.i_fc2_w = add_param(e, dec + "fc2.weight", (long long)HD * HD);
  e.i_fc2_b = add_param(e, dec + "fc2.bias", HD);
  e.i_joint0 = e.i_decpose_w = e.i_decpose_b = e.i_init_pose = e.i_init_shape = e.i_init_cam = -1;
  if (c.decoder == DEC_KTD) {
    for (int j = 0; j < 24; ++j) {
      const int idx = add_param(e, dec + "joint_regs." + std::to_string(j) + ".weight", 6LL * (HD + 6 * kAncCnt[j]));
      if (j == 0) e.i_joint0 = idx;
      add_param(e, dec + "joint_regs." + std::to_string(j) + ".bias", 6);
    }
  } else {
    e.i_decpose_w = add_param(e, dec + "decpose.weight", 144LL * HD);
    e.i_decpose_b = add_param(e, dec + "decpose.bias", 144);
  }
  e.i_shape_w = add_param(e, dec + "decshape.weight", 10LL * HD);
  e.i_shape_b = add_param(e, dec + "decshape.bias", 10);
  e.i_cam_w = add_param(e, dec + "deccam.weight", 3LL * HD);
  e.i_cam_b = add_param(e, dec + "deccam.bias", 3);
  if (c.decoder == DEC_ITERATIVE) {
    e.i_init_pose = add_param(e, dec + "init_pose", 144);
    e.i_init_shape = add_param(e, dec + "init_shape", 10);
    e.i_init_cam = add_param(e, dec + "init_cam", 3);
  }

  // ---- packed (derived) weights: fp16 planes, always laid out for two planes (nsplit=1 uses the first)
  size_t off = 0;
  auto planes = [&](long long elems) { size_t o = off; off = align_up(off + (size_t)elems * 2 * 2); return o; };
  e.off_stem = e.off_proj = e.off_pl = 0;
  if (cnn) {
    cnn_add_packed(e, off);
  } else {
    e.off_stem = planes(64LL * kStemKPad);
    int prev = 64;
    for (int s = 0; s < 3; ++s) {
      const int out = kStageOut[s], mid = out / 4;
      for (int b = 0; b < kStageDepth[s]; ++b) {
        Engine::BlockOff bo;
        bo.ds = (b == 0) ? planes((long long)out * prev) : 0;
        bo.c1 = planes((long long)mid * prev);
        bo.c2 = planes((long long)mid * mid * 9);
        bo.c3 = planes((long long)out * mid);
        e.bb_off.push_back(bo);
        prev = out;
      }
    }
    e.off_proj = planes(768LL * 1024);
    for (int i = 0; i < c.num_blocks; ++i) {
      Engine::SteOff so;
      so.qkv = planes(3LL * C * C);
      so.proj = planes((long long)C * C);
      so.fc1 = planes(4LL * C * C);
      so.fc2 = planes(4LL * C * C);
      so.ts = (c.mode == MODE_PARALLEL) ? planes(4LL * C * C) : 0;
      e.blk_off.push_back(so);
    }
    e.off_pl = planes((long long)C * C);
  }  // !cnn
  e.off_kfc1 = planes((long long)HD * F);
  e.off_kfc2 = planes((long long)HD * HD);
  e.off_kheads = planes(192LL * HD);
  e.off_kheads_b = off; off = align_up(off + 192 * 4);
  e.off_ktd_wx = off; off = align_up(off + 144 * (size_t)HD * 4);
  e.off_ktd_b = off; off = align_up(off + 144 * 4);
  e.off_ktd_anc = off; off = align_up(off + 36 * 95 * 4);
  e.packed_bytes = off;
}

// ------------------------------------------------------------------------------------------ C-level API
int engine_create(const EngineConfig* cfg, Engine** out) {
  MAED_CHECK_ARG(cfg && out, "engine_create: null argument");
  MAED_CHECK_ARG(cfg->encoder == ENC_STE || cfg->encoder == ENC_CNN, "engine: unknown encoder %d", cfg->encoder);
  if (cfg->encoder == ENC_STE) {
    MAED_CHECK_ARG(cfg->num_heads == 12, "engine: num_heads=%d unsupported (head_dim must be 64: num_heads=12)",
                   cfg->num_heads);
    MAED_CHECK_ARG(cfg->num_blocks >= 1 && cfg->num_blocks <= 64, "engine: num_blocks=%d", cfg->num_blocks);
    MAED_CHECK_ARG(cfg->mode >= 0 && cfg->mode <= MODE_TEMPORAL, "engine: unknown st_mode %d", cfg->mode);
  }
  MAED_CHECK_ARG(cfg->decoder == DEC_KTD || cfg->decoder == DEC_ITERATIVE, "engine: unknown decoder %d", cfg->decoder);
  MAED_CHECK_ARG(cfg->nsplit == 1 || cfg->nsplit == 3, "engine: nsplit must be 1 or 3");
  MAED_CHECK_ARG(cfg->hidden_dim >= 64 && cfg->hidden_dim <= 4096, "engine: hidden_dim=%d", cfg->hidden_dim);
  MAED_CHECK_ARG(cfg->temp_frames >= 1 && cfg->temp_frames <= 32, "engine: temp_frames=%d (1..32)", cfg->temp_frames);
  Engine* e = new Engine();
  e->cfg = *cfg;
  { const char* v = getenv("MAED_B200_FUSE_GN"); e->fuse_gn = !(v && v[0] == '0'); }
  build_tables(*e);
  *out = e;
  return MAED_OK;
}
void engine_destroy(Engine* e) { delete e; }
int engine_num_params(const Engine* e) { return (int)e->names.size(); }
const char* engine_param_name(const Engine* e, int i) { return (i >= 0 && i < (int)e->names.size()) ? e->names[i].c_str() : ""; }
long long engine_param_numel(const Engine* e, int i) { return (i >= 0 && i < (int)e->numels.size()) ? e->numels[i] : -1; }
size_t engine_packed_bytes(const Engine* e) { return e->packed_bytes; }

// ---- workspace carving
struct Workspace {
  // backbone
  __half* col; __half* act[5]; float* convout; double* stats;
  long long col_plane, act_plane;
  // STE
  float* x; __half* ln; __half* qkv; float* xs; float* xt; __half* ao; __half* hid; float* tok;
  float* alpha; float* logits; float* vbuf;
  long long ln_plane, qkv_plane, ao_plane, hid_plane;
  // tail
  float* cls; float* tokscratch; __half* alpha_p;
  TailWs tail;
  size_t total;
};
static constexpr long long kColPerImg = 12544LL * kStemKPad;      // largest explicit-im2col matrix (stem)
static constexpr long long kActPerImg = 3136LL * 256;             // largest activation (stage-0 output)
void carve_tail(const Engine& e, int BT, Carver& c, TailWs& t) {
  const int HD = e.cfg.hidden_dim, F = e.feat_dim();
  t.h1 = (float*)c.take((size_t)BT * HD * 4);
  t.h2 = (float*)c.take((size_t)BT * HD * 4);
  t.base = (float*)c.take((size_t)BT * 192 * 4);
  t.xc = (float*)c.take((size_t)BT * (F + 160) * 4);                // iterative: [feat | pose6d | shape | cam] (spin.py:64)
  t.tail_plane = (long long)BT * std::max(std::max(1024, HD), F);
  t.tail_a = (__half*)c.take((size_t)t.tail_plane * 4);
  t.tail_b = (__half*)c.take((size_t)t.tail_plane * 4);
}

static void carve(const Engine& e, int BT, uint8_t* base, Workspace& w) {
  Carver cv(base);
  auto take = [&](size_t bytes) { return (uint8_t*)cv.take(bytes); };
  const long long rows = (long long)BT * 197;
  w.col_plane = kColPerImg * BT;
  w.act_plane = kActPerImg * BT;
  w.col = (__half*)take((size_t)w.col_plane * 2 * 2);
  for (int i = 0; i < 5; ++i) w.act[i] = (__half*)take((size_t)w.act_plane * 2 * 2);
  w.convout = (float*)take((size_t)BT * 12544 * 64 * 4);           // stem conv output is the largest fp32 map
  w.stats = (double*)take((size_t)64 * BT * 32 * 2 * 8);
  w.x = (float*)take((size_t)rows * 768 * 4);
  w.ln_plane = rows * 768; w.ln = (__half*)take((size_t)w.ln_plane * 4);
  w.qkv_plane = rows * 2304; w.qkv = (__half*)take((size_t)w.qkv_plane * 4);
  w.xs = (float*)take((size_t)rows * 768 * 4);
  w.xt = (float*)take((size_t)rows * 768 * 4);
  w.ao_plane = rows * 768; w.ao = (__half*)take((size_t)w.ao_plane * 4);
  w.hid_plane = rows * 3072; w.hid = (__half*)take((size_t)w.hid_plane * 4);
  w.tok = (float*)take((size_t)BT * 196 * 768 * 4);
  w.alpha = (float*)take((size_t)BT * 1536 * 4);
  w.logits = (float*)take((size_t)BT * 1536 * 4);
  w.vbuf = (float*)take((size_t)BT * 768 * 4);
  w.cls = (float*)take((size_t)BT * 768 * 4);
  w.tokscratch = (float*)take((size_t)BT * kTokenChunks * 1536 * 4);
  w.alpha_p = (__half*)take((size_t)BT * 1536 * 4);
  carve_tail(e, BT, cv, w.tail);
  w.total = cv.off;
}
size_t engine_workspace_bytes(const Engine* e, int BT) {
  if (e->cfg.encoder == ENC_CNN) return cnn_workspace_bytes(*e, BT);
  Workspace w;
  carve(*e, BT, nullptr, w);
  return w.total + 1024;
}

// ---- weight packing
int engine_pack(const Engine* e, const void* const* params, void* packed, cudaStream_t st) {
  MAED_CHECK_ARG(e && params && packed, "engine_pack: null argument");
  uint8_t* pk = (uint8_t*)packed;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto H = [&](size_t off) { return (__half*)(pk + off); };
  // ~90 independent small kernels and ~70 small copies: spread over side streams, joined before the dependent tail
  ForkJoin fj(st);
  if (e->cfg.encoder == ENC_CNN) {
    MAED_PROPAGATE(cnn_pack(*e, params, packed, st));
  } else {
    MAED_PROPAGATE(prep_conv_weight(P(e->i_stem_w), 64, 3, 7, 7, kStemKPad, 1, H(e->off_stem), 64LL * kStemKPad, fj.next()));
    int prev = 64, bi = 0;
    for (int s = 0; s < 3; ++s) {
      const int out = kStageOut[s], mid = out / 4;
      for (int b = 0; b < kStageDepth[s]; ++b, ++bi) {
        const Engine::BlockIdx& ix = e->bb[bi];
        const Engine::BlockOff& of = e->bb_off[bi];
        if (b == 0)
          MAED_PROPAGATE(prep_conv_weight(P(ix.ds_w), out, prev, 1, 1, prev, 1, H(of.ds), (long long)out * prev, fj.next()));
        MAED_PROPAGATE(prep_conv_weight(P(ix.c1_w), mid, prev, 1, 1, prev, 1, H(of.c1), (long long)mid * prev, fj.next()));
        MAED_PROPAGATE(prep_conv_weight(P(ix.c2_w), mid, mid, 3, 3, 9 * mid, 1, H(of.c2), (long long)mid * mid * 9, fj.next()));
        MAED_PROPAGATE(prep_conv_weight(P(ix.c3_w), out, mid, 1, 1, mid, 1, H(of.c3), (long long)out * mid, fj.next()));
        prev = out;
      }
    }
    MAED_PROPAGATE(split_f32(P(e->i_proj_w), H(e->off_proj), 768LL * 1024, 768LL * 1024, fj.next()));
    const long long CC = 768LL * 768;
    for (int i = 0; i < e->cfg.num_blocks; ++i) {
      const Engine::SteIdx& ix = e->blk[i];
      const Engine::SteOff& of = e->blk_off[i];
      MAED_PROPAGATE(split_f32(P(ix.qkv_w), H(of.qkv), 3 * CC, 3 * CC, fj.next()));
      MAED_PROPAGATE(split_f32(P(ix.proj_w), H(of.proj), CC, CC, fj.next()));
      MAED_PROPAGATE(split_f32(P(ix.fc1_w), H(of.fc1), 4 * CC, 4 * CC, fj.next()));
      MAED_PROPAGATE(split_f32(P(ix.fc2_w), H(of.fc2), 4 * CC, 4 * CC, fj.next()));
      if (e->cfg.mode == MODE_PARALLEL) MAED_PROPAGATE(split_f32(P(ix.ts_w), H(of.ts), 4 * CC, 4 * CC, fj.next()));
    }
    MAED_PROPAGATE(split_f32(P(e->i_pl_w), H(e->off_pl), CC, CC, fj.next()));
  }  // !cnn
  if (e->cfg.decoder == DEC_KTD) {
    // joint_regs.j.weight [6, HD + 6k] -> Wx rows (first HD columns), ancestor blocks (last 6k columns), biases
    const int HD = e->cfg.hidden_dim;
    float* wx = (float*)(pk + e->off_ktd_wx);
    float* bj = (float*)(pk + e->off_ktd_b);
    float* wa = (float*)(pk + e->off_ktd_anc);
    size_t aoff = 0;
    for (int j = 0; j < 24; ++j) {
      const int k = kAncCnt[j];
      const float* wj = P(e->i_joint0 + 2 * j);
      const float* bjs = P(e->i_joint0 + 2 * j + 1);
      const size_t src_pitch = (size_t)(HD + 6 * k) * 4;
      cudaStream_t js = fj.next();
      MAED_CUDA_CHECK(cudaMemcpy2DAsync(wx + (size_t)j * 6 * HD, (size_t)HD * 4, wj, src_pitch, (size_t)HD * 4, 6,
                                        cudaMemcpyDeviceToDevice, js));
      if (k > 0)
        MAED_CUDA_CHECK(cudaMemcpy2DAsync(wa + aoff, (size_t)6 * k * 4, wj + HD, src_pitch, (size_t)6 * k * 4, 6,
                                          cudaMemcpyDeviceToDevice, js));
      MAED_CUDA_CHECK(cudaMemcpyAsync(bj + j * 6, bjs, 24, cudaMemcpyDeviceToDevice, js));
      aoff += 36 * k;
    }
    MAED_PROPAGATE(fj.join());
    // tensor-core operands of the tail: fc1, fc2 and one [192, HD] head matrix = [24 joint bases (144) | shape (10) |
    // cam (3) | 35 zero rows] with its bias vector
    const long long F = e->feat_dim();
    MAED_PROPAGATE(split_f32(P(e->i_fc1_w), H(e->off_kfc1), HD * F, HD * F, st));
    MAED_PROPAGATE(split_f32(P(e->i_fc2_w), H(e->off_kfc2), (long long)HD * HD, (long long)HD * HD, st));
    float* hb = (float*)(pk + e->off_kheads_b);
    MAED_CUDA_CHECK(cudaMemsetAsync(pk + e->off_kheads, 0, (size_t)192 * HD * 4, st));
    MAED_CUDA_CHECK(cudaMemsetAsync(hb, 0, 192 * 4, st));
    const long long hp = 192LL * HD;
    MAED_PROPAGATE(split_f32(wx, H(e->off_kheads), hp, 144LL * HD, st));
    MAED_PROPAGATE(split_f32(P(e->i_shape_w), H(e->off_kheads) + 144LL * HD, hp, 10LL * HD, st));
    MAED_PROPAGATE(split_f32(P(e->i_cam_w), H(e->off_kheads) + 154LL * HD, hp, 3LL * HD, st));
    MAED_CUDA_CHECK(cudaMemcpyAsync(hb, bj, 144 * 4, cudaMemcpyDeviceToDevice, st));
    MAED_CUDA_CHECK(cudaMemcpyAsync(hb + 144, P(e->i_shape_b), 40, cudaMemcpyDeviceToDevice, st));
    MAED_CUDA_CHECK(cudaMemcpyAsync(hb + 154, P(e->i_cam_b), 12, cudaMemcpyDeviceToDevice, st));
  }
  return MAED_OK;
}

// ---- forward
static int conv_gn(const Engine& e, Workspace& w, int& gn_counter, int BT, const __half* A, long long a_plane, int a_is_col,
                   int Hin, int Cin, int Cout, int ksz, int stride, const __half* Wp, long long w_plane, const float* gamma,
                   const float* beta, int relu, const __half* res, long long res_plane, __half* out, long long out_plane,
                   cudaStream_t st) {
  // returns the conv+GN(+res)(+relu) result as planes; SAME padding (resnetv2.py:51-59)
  (void)a_is_col;
  const int Hout = (Hin + stride - 1) / stride;
  const int M = BT * Hout * Hout;
  GemmArgs g;
  g.nsplit = e.cfg.nsplit;
  g.B = Wp; g.b_plane = w_plane;
  g.M = M; g.N = Cout; g.out_mode = OUT_F32; g.out = w.convout; g.ldc = Cout;
  const int pad_total = std::max((Hout - 1) * stride + ksz - Hin, 0);
  if (ksz == 1 && stride == 1) {
    g.A = A; g.a_plane = a_plane; g.K = Cin;
  } else if (stride == 1) {
    g.A = A; g.a_plane = a_plane; g.K = ksz * ksz * Cin;
    g.conv = 1; g.n_img = BT; g.H = Hin; g.W = Hin; g.Cin = Cin; g.KH = ksz; g.KW = ksz;
    g.pad_h = pad_total / 2; g.pad_w = pad_total / 2;
  } else {
    MAED_PROPAGATE(im2col_nhwc(A, a_plane, BT, Hin, Hin, Cin, ksz, ksz, stride, pad_total / 2, pad_total / 2, Hout, Hout,
                               w.col, w.col_plane, st));
    g.A = w.col; g.a_plane = w.col_plane; g.K = ksz * ksz * Cin;
  }
  if (e.fuse_gn) {
    ConvGnArgs f;
    f.A = g.A; f.a_plane = g.a_plane; f.B = Wp; f.b_plane = w_plane;
    f.n_img = BT; f.H_out = Hout; f.W_out = Hout; f.C = Cout; f.K = g.K; f.nsplit = e.cfg.nsplit;
    f.conv = g.conv; f.Cin = Cin; f.KH = ksz; f.KW = ksz; f.pad_h = g.pad_h; f.pad_w = g.pad_w;
    f.gamma = gamma; f.beta = beta; f.eps = 1e-5f; f.relu = relu; f.res = res; f.res_plane = res_plane;
    f.out = out; f.out_plane = out_plane;
    const int rc = conv_gn_fused(f, st);
    if (rc == MAED_OK) return MAED_OK;
    if (rc != MAED_ERR_UNSUPPORTED) return rc;
  }
  MAED_PROPAGATE(launch_gemm(g, st));
  double* stats = w.stats + (size_t)gn_counter * BT * 64;
  ++gn_counter;
  MAED_PROPAGATE(gn_stats(w.convout, BT, Hout * Hout, Cout, stats, st));
  MAED_PROPAGATE(gn_apply(w.convout, stats, gamma, beta, BT, Hout * Hout, Cout, 1e-5f, relu, res, res_plane, out, out_plane, st));
  return MAED_OK;
}

int engine_forward(const Engine* ep, const void* const* params, const void* packed, const float* x_in, int N, int T,
                   void* workspace, size_t workspace_bytes, const EngineOutputs* outs, float* const* taps,
                   cudaStream_t st) {
  MAED_CHECK_ARG(ep && params && packed && x_in && workspace && outs, "engine_forward: null argument");
  const Engine& e = *ep;
  const EngineConfig& c = e.cfg;
  const int BT = N * T;
  MAED_CHECK_ARG(N >= 1 && T >= 1, "engine_forward: empty batch N=%d T=%d", N, T);
  if (c.encoder == ENC_CNN) return cnn_forward(e, params, packed, x_in, N, T, workspace, workspace_bytes, outs, taps, st);
  const bool has_temp = e.i_temp >= 0;
  MAED_CHECK_ARG(!has_temp || T <= c.temp_frames, "engine_forward: seqlen T=%d exceeds temp_embed frames %d "
                 "(reference vision_transformer.py:364,398 raises a broadcast error here)", T, c.temp_frames);
  MAED_CHECK_ARG(T <= 32, "engine_forward: T=%d > 32 unsupported by the temporal attention kernel", T);
  Workspace w;
  carve(e, BT, (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023), w);
  MAED_CHECK_ARG(w.total + 1024 <= workspace_bytes, "engine_forward: workspace too small (%zu < %zu)", workspace_bytes,
                 w.total + 1024);
  const uint8_t* pk = (const uint8_t*)packed;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto Hh = [&](size_t off) { return (const __half*)(pk + off); };
  const int ns = c.nsplit;
  const int om = ns == 3 ? OUT_F16_SPLIT : OUT_F16;

  // ------------------------------------------------------------------------------------- backbone
  MAED_CUDA_CHECK(cudaMemsetAsync(w.stats, 0, (size_t)64 * BT * 32 * 2 * 8, st));
  int gn = 0;
  {
    // stem: 7x7/2 SAME (pad 2 top/left, 3 bottom/right) -> GN+ReLU -> 3x3/2 SAME max-pool
    double* stats = w.stats + (size_t)gn * BT * 64;
    ++gn;
    MAED_PROPAGATE(stem_conv(x_in, BT, Hh(e.off_stem), 64LL * kStemKPad, kStemKPad, ns, w.convout, stats, st));
    MAED_PROPAGATE(gn_apply_maxpool(w.convout, stats, P(e.i_stem_g), P(e.i_stem_g + 1), BT, 112, 112, 64, 1e-5f, w.act[0],
                                    w.act_plane, st));
  }
  if (taps && taps[TAP_STEM]) MAED_PROPAGATE(planes_to_f32(w.act[0], ns == 3 ? w.act_plane : 0, (long long)BT * 3136 * 64, taps[TAP_STEM], st));
  __half* cur = w.act[0];
  __half* nxt = w.act[1];
  __half* t1 = w.act[2];
  __half* t2 = w.act[3];
  __half* sc = w.act[4];
  int prev = 64, Hc = 56, bi = 0;
  const long long ap = w.act_plane;
  for (int s = 0; s < 3; ++s) {
    const int out = kStageOut[s], mid = out / 4;
    for (int b = 0; b < kStageDepth[s]; ++b, ++bi) {
      const Engine::BlockIdx& ix = e.bb[bi];
      const Engine::BlockOff& of = e.bb_off[bi];
      const int stride = (s > 0 && b == 0) ? 2 : 1;
      const int Ho = Hc / stride;
      const __half* shortcut = cur;
      if (b == 0) {
        MAED_PROPAGATE(conv_gn(e, w, gn, BT, cur, ap, 0, Hc, prev, out, 1, stride, Hh(of.ds), (long long)out * prev, P(ix.ds_g),
                               P(ix.ds_g + 1), 0, nullptr, 0, sc, ap, st));
        shortcut = sc;
      }
      MAED_PROPAGATE(conv_gn(e, w, gn, BT, cur, ap, 0, Hc, prev, mid, 1, 1, Hh(of.c1), (long long)mid * prev, P(ix.c1_g),
                             P(ix.c1_g + 1), 1, nullptr, 0, t1, ap, st));
      MAED_PROPAGATE(conv_gn(e, w, gn, BT, t1, ap, 0, Hc, mid, mid, 3, stride, Hh(of.c2), (long long)mid * mid * 9, P(ix.c2_g),
                             P(ix.c2_g + 1), 1, nullptr, 0, t2, ap, st));
      MAED_PROPAGATE(conv_gn(e, w, gn, BT, t2, ap, 0, Ho, mid, out, 1, 1, Hh(of.c3), (long long)out * mid, P(ix.c3_g),
                             P(ix.c3_g + 1), 1, shortcut, ap, nxt, ap, st));
      std::swap(cur, nxt);
      prev = out;
      Hc = Ho;
    }
    if (taps && taps[TAP_STAGE0 + s])
      MAED_PROPAGATE(planes_to_f32(cur, ns == 3 ? ap : 0, (long long)BT * Hc * Hc * prev, taps[TAP_STAGE0 + s], st));
  }

  // ------------------------------------------------------------------------------- patch embedding
  const int ntok = 197, C = 768, heads = c.num_heads;
  const int rows = BT * ntok;
  {
    GemmArgs g;
    g.nsplit = ns;
    g.A = cur; g.a_plane = ap; g.B = Hh(e.off_proj); g.b_plane = 768LL * 1024;
    g.M = BT * 196; g.N = 768; g.K = 1024; g.bias = P(e.i_proj_b); g.out_mode = OUT_F32; g.out = w.tok; g.ldc = 768;
    MAED_PROPAGATE(launch_gemm(g, st));
    MAED_PROPAGATE(embed_assemble(w.tok, P(e.i_cls), P(e.i_pos), has_temp ? P(e.i_temp) : nullptr, BT, T, ntok, C, w.x, st));
  }
  if (taps && taps[TAP_EMBED]) MAED_CUDA_CHECK(cudaMemcpyAsync(taps[TAP_EMBED], w.x, (size_t)rows * C * 4, cudaMemcpyDeviceToDevice, st));

  // ------------------------------------------------------------------------------------ STE blocks
  const float scale = 0.125f;                               // head_dim ** -0.5 (vision_transformer.py:121)
  auto linear_tc = [&](const __half* A, long long a_plane, int M, int K, size_t w_off, int Nout, const float* bias, int act,
                       const float* residual, int out_mode, void* out, long long out_plane) -> int {
    GemmArgs g;
    g.nsplit = ns;
    g.A = A; g.a_plane = a_plane; g.B = Hh(w_off); g.b_plane = (long long)Nout * K;
    g.M = M; g.N = Nout; g.K = K; g.bias = bias; g.act = act; g.residual = residual; g.out_mode = out_mode; g.out = out;
    g.out_plane = out_plane; g.ldc = Nout;
    return launch_gemm(g, st);
  };
  for (int i = 0; i < c.num_blocks; ++i) {
    const Engine::SteIdx& ix = e.blk[i];
    const Engine::SteOff& of = e.blk_off[i];
    if (c.mode == MODE_TEMPORAL) {
      // x.mean(dim=1) of LN1(x) -> qkv -> attention across frames -> proj, broadcast over tokens (:167-173)
      MAED_PROPAGATE(layernorm_f32(w.x, C, P(ix.n1), P(ix.n1 + 1), rows, C, 1e-6f, w.xs, st));
      MAED_PROPAGATE(token_mean(w.xs, BT, ntok, C, w.vbuf, C, 0, st));
      MAED_PROPAGATE(split_f32(w.vbuf, w.ln, w.ln_plane, (long long)BT * C, st));
      MAED_PROPAGATE(linear_tc(w.ln, w.ln_plane, BT, C, of.qkv, 3 * C, P(ix.qkv_b), ACT_NONE, nullptr, om, w.qkv, w.qkv_plane));
      MAED_PROPAGATE(attn_temporal(w.qkv, ns == 3 ? w.qkv_plane : 0, N, T, 1, heads, scale, nullptr, w.ao, w.ao_plane, st));
      MAED_PROPAGATE(linear_tc(w.ao, w.ao_plane, BT, C, of.proj, C, P(ix.proj_b), ACT_NONE, nullptr, OUT_F32, w.vbuf, 0));
      MAED_PROPAGATE(broadcast_add(w.x, w.vbuf, BT, ntok, C, st));
    } else {
      MAED_PROPAGATE(layernorm_planes(w.x, C, P(ix.n1), P(ix.n1 + 1), rows, C, 1e-6f, w.ln, w.ln_plane, st));
      MAED_PROPAGATE(linear_tc(w.ln, w.ln_plane, rows, C, of.qkv, 3 * C, P(ix.qkv_b), ACT_NONE, nullptr, om, w.qkv, w.qkv_plane));
      const long long qp = ns == 3 ? w.qkv_plane : 0;
      if (c.mode == MODE_PARALLEL) {
        MAED_PROPAGATE(attn_temporal(w.qkv, qp, N, T, ntok, heads, scale, w.xt, nullptr, 0, st));
        MAED_PROPAGATE(attn_spatial(w.qkv, w.qkv_plane, BT, ntok, heads, scale, ns, w.xs, nullptr, 0, st));
        MAED_PROPAGATE(token_mean2_planes(w.xs, w.xt, BT, ntok, C, w.tokscratch, w.alpha_p, (long long)BT * 2 * C, st));
        MAED_PROPAGATE(linear_tc(w.alpha_p, (long long)BT * 2 * C, BT, 2 * C, of.ts, 2 * C, P(ix.ts_b), ACT_NONE, nullptr, OUT_F32,
                                 w.logits, 0));
        MAED_PROPAGATE(ts_blend(w.xs, w.xt, w.logits, BT, ntok, C, w.ao, w.ao_plane, st));
      } else if (c.mode == MODE_SERIES) {
        MAED_PROPAGATE(attn_spatial(w.qkv, w.qkv_plane, BT, ntok, heads, scale, ns, nullptr, w.ao, w.ao_plane, st));
        MAED_PROPAGATE(linear_tc(w.ao, w.ao_plane, rows, C, of.qkv, 3 * C, P(ix.qkv_b), ACT_NONE, nullptr, om, w.qkv, w.qkv_plane));
        MAED_PROPAGATE(attn_temporal(w.qkv, qp, N, T, ntok, heads, scale, nullptr, w.ao, w.ao_plane, st));
      } else if (c.mode == MODE_VANILLA) {
        MAED_PROPAGATE(attn_spatial(w.qkv, w.qkv_plane, BT, ntok, heads, scale, ns, nullptr, w.ao, w.ao_plane, st));
      } else {  // coupling: joint attention over the T*197 tokens of a clip
        MAED_PROPAGATE(attn_generic(w.qkv, qp, N, T * ntok, heads, scale, ntok, T, nullptr, w.ao, w.ao_plane, st));
      }
      MAED_PROPAGATE(linear_tc(w.ao, w.ao_plane, rows, C, of.proj, C, P(ix.proj_b), ACT_NONE, w.x, OUT_F32, w.x, 0));
    }
    MAED_PROPAGATE(layernorm_planes(w.x, C, P(ix.n2), P(ix.n2 + 1), rows, C, 1e-6f, w.ln, w.ln_plane, st));
    MAED_PROPAGATE(linear_tc(w.ln, w.ln_plane, rows, C, of.fc1, 4 * C, P(ix.fc1_b), ACT_GELU, nullptr, om, w.hid, w.hid_plane));
    MAED_PROPAGATE(linear_tc(w.hid, w.hid_plane, rows, 4 * C, of.fc2, C, P(ix.fc2_b), ACT_NONE, w.x, OUT_F32, w.x, 0));
    if (taps && i < 8 && taps[TAP_BLOCK0 + i])
      MAED_CUDA_CHECK(cudaMemcpyAsync(taps[TAP_BLOCK0 + i], w.x, (size_t)rows * C * 4, cudaMemcpyDeviceToDevice, st));
  }

  // ------------------------------------------------------------------------------------------ tail
  // The tail always runs in split precision (nsplit = 3), also in the fp16 fast mode: its GEMMs feed the outputs
  // without a damping residual path and are ~0.5 GFLOP.
  auto tail_tc = [&](const __half* A, long long a_plane, int K, size_t w_off, int Nout, const float* bias, int act,
                     int out_mode, void* out, long long out_plane) -> int {
    GemmArgs g;
    g.nsplit = 3;
    g.A = A; g.a_plane = a_plane; g.B = Hh(w_off); g.b_plane = (long long)Nout * K;
    g.M = BT; g.N = Nout; g.K = K; g.bias = bias; g.act = act; g.out_mode = out_mode; g.out = out; g.out_plane = out_plane;
    g.ldc = Nout;
    return launch_gemm(g, st);
  };
  MAED_PROPAGATE(layernorm_planes(w.x, (long long)ntok * C, P(e.i_norm), P(e.i_norm + 1), BT, C, 1e-6f, w.tail.tail_a,
                                  w.tail.tail_plane, st));
  MAED_PROPAGATE(tail_tc(w.tail.tail_a, w.tail.tail_plane, C, e.off_pl, C, P(e.i_pl_b), ACT_TANH, OUT_F32, outs->feat, 0));
  return run_decoder(e, params, pk, BT, w.tail, outs, st);
}

// Decoder from the encoder feature outs->feat [BT, F] (F = 768 'ste', 2048 'cnn'): KTD (ktd.py:69-88) as three
// split-precision tensor-core GEMMs + the kinematic-tree pass, or the iterative regressor (spin.py:51-74) in fp32;
// then rot6d -> rotmat -> angle-axis, theta, kp_2d.
int run_decoder(const Engine& e, const void* const* params, const uint8_t* pk, int BT, const TailWs& w, const EngineOutputs* outs,
                cudaStream_t st) {
  const EngineConfig& c = e.cfg;
  auto P = [&](int i) { return (const float*)params[i]; };
  auto Hh = [&](size_t off) { return (const __half*)(pk + off); };
  const int HD = c.hidden_dim;
  const int C = e.feat_dim();
  auto tail_tc = [&](const __half* A, long long a_plane, int K, size_t w_off, int Nout, const float* bias, int act,
                     int out_mode, void* out, long long out_plane) -> int {
    GemmArgs g;
    g.nsplit = 3;
    g.A = A; g.a_plane = a_plane; g.B = Hh(w_off); g.b_plane = (long long)Nout * K;
    g.M = BT; g.N = Nout; g.K = K; g.bias = bias; g.act = act; g.out_mode = out_mode; g.out = out; g.out_plane = out_plane;
    g.ldc = Nout;
    return launch_gemm(g, st);
  };
  if (c.decoder == DEC_KTD) {
    MAED_PROPAGATE(split_f32(outs->feat, w.tail_a, w.tail_plane, (long long)BT * C, st));
    MAED_PROPAGATE(tail_tc(w.tail_a, w.tail_plane, C, e.off_kfc1, HD, P(e.i_fc1_b), ACT_NONE, OUT_F16_SPLIT, w.tail_b, w.tail_plane));
    MAED_PROPAGATE(tail_tc(w.tail_b, w.tail_plane, HD, e.off_kfc2, HD, P(e.i_fc2_b), ACT_NONE, OUT_F16_SPLIT, w.tail_a, w.tail_plane));
    MAED_PROPAGATE(tail_tc(w.tail_a, w.tail_plane, HD, e.off_kheads, 192, (const float*)(pk + e.off_kheads_b), ACT_NONE, OUT_F32,
                           w.base, 0));
    MAED_PROPAGATE(ktd_tree(w.base, 192, (const float*)(pk + e.off_ktd_anc), BT, outs->pose6d, outs->shape, outs->cam, st));
  } else {
    const int KI = C + 157;
    MAED_PROPAGATE(broadcast_row(P(e.i_init_pose), 144, BT, outs->pose6d, st));
    MAED_PROPAGATE(broadcast_row(P(e.i_init_shape), 10, BT, outs->shape, st));
    MAED_PROPAGATE(broadcast_row(P(e.i_init_cam), 3, BT, outs->cam, st));
    for (int it = 0; it < 3; ++it) {                      // spin.py:63-72
      MAED_PROPAGATE(concat_cols(outs->feat, C, outs->pose6d, 144, outs->shape, 10, outs->cam, 3, BT, w.xc, st));
      MAED_PROPAGATE(linear_f32(w.xc, KI, P(e.i_fc1_w), KI, P(e.i_fc1_b), BT, HD, KI, 0, nullptr, 0, w.h1, HD, st));
      MAED_PROPAGATE(linear_f32(w.h1, HD, P(e.i_fc2_w), HD, P(e.i_fc2_b), BT, HD, HD, 0, nullptr, 0, w.h2, HD, st));
      MAED_PROPAGATE(linear_f32(w.h2, HD, P(e.i_decpose_w), HD, P(e.i_decpose_b), BT, 144, HD, 0, outs->pose6d, 144, outs->pose6d, 144, st));
      MAED_PROPAGATE(linear_f32(w.h2, HD, P(e.i_shape_w), HD, P(e.i_shape_b), BT, 10, HD, 0, outs->shape, 10, outs->shape, 10, st));
      MAED_PROPAGATE(linear_f32(w.h2, HD, P(e.i_cam_w), HD, P(e.i_cam_b), BT, 3, HD, 0, outs->cam, 3, outs->cam, 3, st));
    }
  }
  MAED_PROPAGATE(decode_outputs(outs->pose6d, outs->shape, outs->cam, BT, outs->kp3d, outs->n_joints, outs->rotmat, outs->theta,
                                outs->kp2d, st));
  MAED_PROPAGATE(decode_outputs(outs->pose6d, outs->shape, outs->cam, BT, outs->kp3d, outs->n_joints, outs->rotmat, outs->theta,
                                outs->kp2d, st));
  return MAED_OK;
}

}  // namespace maed
