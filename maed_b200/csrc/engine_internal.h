// Internal tables of the MAED engine shared by the forward (engine.cu) and the training path (train.cu):
// parameter indices in reference state_dict order and offsets of the derived (packed) tensor-core weights.
#pragma once
#include <string>
#include <vector>

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
  int feat_dim() const { return 768; }
  int np() const { return cfg.nsplit == 3 ? 2 : 1; }
};

}  // namespace maed
