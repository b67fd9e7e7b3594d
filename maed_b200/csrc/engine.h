// Forward engine of the B200-native MAED hot path (engine.cu).
#pragma once
#include "common.h"

namespace maed {

enum : int { MODE_VANILLA = 0, MODE_PARALLEL = 1, MODE_SERIES = 2, MODE_COUPLING = 3, MODE_TEMPORAL = 4 };
enum : int { DEC_KTD = 0, DEC_ITERATIVE = 1 };
enum : int { ENC_STE = 0, ENC_CNN = 1 };
// debug taps (fp32 copies of intermediates, NHWC for the backbone): indices into the `taps` array
enum : int { TAP_STEM = 0, TAP_STAGE0 = 1, TAP_STAGE1 = 2, TAP_STAGE2 = 3, TAP_EMBED = 4, TAP_BLOCK0 = 5, TAP_COUNT = 13 };

struct EngineConfig {
  int num_blocks;     // STE depth (reference MODEL.ENCODER.NUM_BLOCKS)
  int num_heads;      // 12 (head_dim 64)
  int mode;           // MODE_*  (reference st_mode)
  int decoder;        // DEC_*
  int hidden_dim;     // decoder hidden size (1024)
  int nsplit;         // 3: split-fp16 (parity mode), 1: plain fp16 (fast, fails the 1e-3 gate)
  int temp_frames;    // rows of temp_embed (16 in the reference)
  int encoder;        // ENC_STE (hybrid ResNetV2 + STE, feature 768) or ENC_CNN (torchvision ResNet-50, feature 2048;
                      // reference maed.py:35-37, inference only); num_blocks / num_heads / mode are ignored for ENC_CNN
};

struct EngineOutputs {
  float* feat;        // [BT, 768] (ENC_CNN: [BT, 2048])
  float* pose6d;      // [BT, 144]
  float* shape;       // [BT, 10]
  float* cam;         // [BT, 3]
  float* rotmat;      // [BT, 24, 3, 3]
  float* theta;       // [BT, 85]
  float* kp2d;        // [BT, n_joints, 2]
  const float* kp3d;  // [BT, n_joints, 3] or nullptr (zeros)
  int n_joints;
};

struct Engine;
int engine_create(const EngineConfig* cfg, Engine** out);
void engine_destroy(Engine* e);
int engine_num_params(const Engine* e);
const char* engine_param_name(const Engine* e, int i);
long long engine_param_numel(const Engine* e, int i);
size_t engine_packed_bytes(const Engine* e);
size_t engine_workspace_bytes(const Engine* e, int BT);
int engine_pack(const Engine* e, const void* const* params, void* packed, cudaStream_t st);
int engine_forward(const Engine* e, const void* const* params, const void* packed, const float* x, int N, int T,
                   void* workspace, size_t workspace_bytes, const EngineOutputs* outs, float* const* taps, cudaStream_t st);

}  // namespace maed
