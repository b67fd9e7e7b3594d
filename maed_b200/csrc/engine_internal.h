// Internal tables of the MAED engine shared by the forward (engine.cu) and the training path (train.cu):
// parameter indices in reference state_dict order and offsets of the derived (packed) tensor-core weights.
#pragma once
#include <string>
#include <vector>

#include "bwd_kernels.h"
#include "engine.h"

namespace maed {

static const int kAncCnt[24] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8};
static const int kStageDepth[3] = {3, 4, 9};
static const int kStageOut[3] = {256, 512, 1024};
static constexpr int kStemKPad = 152;        // 7*7*3 = 147 padded to a multiple of 8 (16-byte TMA rows)

struct Engine {
  EngineConfig cfg;
  std::vector<std::string> names;          // reference state_dict keys, in engine order
  std::vector<long long> numels;
  // indices into the parameter table
  int i_cls, i_pos, i_temp;
  int i_stem_w, i_stem_g;                   // conv weight ; norm weight (bias = +1)
  struct BlockIdx { int ds_w, ds_g, c1_w, c1_g, c2_w, c2_g, c3_w, c3_g; };
  std::vector<BlockIdx> bb;                 // 16 bottlenecks
  int i_proj_w, i_proj_b;
  struct SteIdx { int n1, qkv_w, qkv_b, ts_w, ts_b, proj_w, proj_b, n2, fc1_w, fc1_b, fc2_w, fc2_b; };
  std::vector<SteIdx> blk;
  int i_norm, i_pl_w, i_pl_b;
  int i_fc1_w, i_fc1_b, i_fc2_w, i_fc2_b, i_joint0, i_shape_w, i_shape_b, i_cam_w, i_cam_b;
  int i_decpose_w, i_decpose_b, i_init_pose, i_init_shape, i_init_cam;
  // packed weight offsets (bytes)
  size_t off_stem;
  struct BlockOff { size_t ds, c1, c2, c3; };
  std::vector<BlockOff> bb_off;
  size_t off_proj;
  struct SteOff { size_t qkv, proj, fc1, fc2, ts; };
  std::vector<SteOff> blk_off;
  size_t off_ktd_wx, off_ktd_b, off_ktd_anc;
  size_t off_pl, off_kfc1, off_kfc2, off_kheads, off_kheads_b;   // tensor-core planes of the tail (KTD)
  size_t packed_bytes;
  bool fuse_gn = true;                      // MAED_B200_FUSE_GN=0 selects the unfused conv / gn_stats / gn_apply kernels
  // 'cnn' encoder (torchvision ResNet-50, cnn_engine.cu): the 53 conv+BN pairs in forward order
  // (stem; per bottleneck: [downsample], conv1, conv2, conv3)
  struct CnnConv {
    int w, bn;                              // parameter indices: conv weight; bn.weight (bias, running_mean, running_var follow)
    int cin, cout, k, stride, k_pad;        // k_pad: GEMM K (k*k*cin, stem padded to 152)
    size_t off_w, off_b;                    // packed: BN-folded weight planes [cout][k_pad]; folded bias fp32 [cout]
    size_t off_w_raw;                       // packed: the un-folded weight planes (training: BatchNorm on batch statistics)
  };
  std::vector<CnnConv> cnn;
  size_t off_cnn_scratch = 0;               // fp32 scratch for one BN-scaled weight tensor (pack time only)
  BnExchange bn_exchange{nullptr, nullptr, nullptr, 0};   // SyncBatchNorm hook of the 'cnn' training path (train_set_exchange)
  // backward progress hook (train_set_progress): called on the host after the kernels of a stage have been ENQUEUED
  struct Progress { int (*fn)(void* user, int first_param, int end_param); void* user; } progress{nullptr, nullptr};
  int feat_dim() const { return cfg.encoder == ENC_CNN ? 2048 : 768; }
  int np() const { return cfg.nsplit == 3 ? 2 : 1; }
};

// ---- workspace carving shared by the two encoders
struct Carver {
  uint8_t* base; size_t off = 0;
  explicit Carver(uint8_t* b) : base(b) {}
  void* take(size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off = (off + bytes + 1023) / 1024 * 1024;
    return p;
  }
};
// buffers of the decoder tail (KTD: split-precision tensor-core GEMMs; iterative: fp32)
struct TailWs {
  float* h1; float* h2; float* base; float* xc;
  __half* tail_a; __half* tail_b; long long tail_plane;
};
void carve_tail(const Engine& e, int BT, Carver& c, TailWs& t);
// decoder (reference ktd.py:69-124 / spin.py:51-110) from the encoder feature in outs->feat [BT, feat_dim]
int run_decoder(const Engine& e, const void* const* params, const uint8_t* packed, int BT, const TailWs& t,
                const EngineOutputs* outs, cudaStream_t st);

// planes (hi + lo when plane != 0) -> fp32, n elements (debug taps)
int planes_to_f32(const __half* hi, long long plane, long long n, float* out, cudaStream_t st);

// dst[r, :] = src[0..C) for R rows
int broadcast_row(const float* src, int C, int R, float* dst, cudaStream_t st);

// ---- 'cnn' encoder (cnn_engine.cu)
void cnn_add_params(Engine& e, int (*add_param)(Engine&, const std::string&, long long));
void cnn_add_packed(Engine& e, size_t& off);
size_t cnn_workspace_bytes(const Engine& e, int BT);
int cnn_pack(const Engine& e, const void* const* params, void* packed, cudaStream_t st);
int cnn_forward(const Engine& e, const void* const* params, const void* packed, const float* x, int N, int T, void* workspace,
                size_t workspace_bytes, const EngineOutputs* outs, float* const* taps, cudaStream_t st);

}  // namespace maed
