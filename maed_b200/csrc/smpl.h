// SMPL forward (smpl.cu).  All pointers are fp32 / int32 device memory.
#pragma once
#include "common.h"

namespace maed {

struct SmplAssets {
  const float* v_template;         // [6890, 3]
  const float* shapedirs;          // [6890*3, 10]
  const float* posedirs;           // [207, 6890*3]
  const float* J_template;         // [24, 3]        = J_regressor @ v_template          (derived once)
  const float* J_shapedirs;        // [24*3, 10]     = J_regressor @ shapedirs           (derived once)
  const float* lbs_weights;        // [6890, 24]
  const float* J_regressor_extra;  // [9, 6890]      (lib/models/smpl.py:90-91)
  const int* parents;              // [24], -1 for the root
  const int* extra_vertex_ids;     // [21]  vertex-selected joints (smplx VertexJointSelector)
  const int* joint_map;            // [49]  indices into the 54 joints (lib/models/smpl.py:15-55,89,99)
};

size_t smpl_scratch_bytes(int BT);
// betas [BT,10], rotmat [BT,24,3,3] -> verts [BT,6890,3]; joints [BT,49,3], or [BT,n_reg,3] = J_regressor @ verts
int smpl_forward(const SmplAssets* a, const float* betas, const float* rotmat, int BT, const float* J_regressor, int n_reg,
                 float* verts, float* joints, void* scratch, size_t scratch_bytes, cudaStream_t st);

// gradients of (verts, joints) w.r.t. (betas, rotmat): d_verts [BT,6890,3] or nullptr, d_joints [BT,49 | n_reg,3] ->
// d_betas [BT,10], d_rotmat [BT,24,3,3] (overwritten).  scratch: smpl_backward_scratch_bytes(BT).
size_t smpl_backward_scratch_bytes(int BT);
int smpl_backward(const SmplAssets* a, const float* betas, const float* rotmat, int BT, const float* J_regressor, int n_reg,
                  const float* d_verts, const float* d_joints, float* d_betas, float* d_rotmat, void* scratch,
                  size_t scratch_bytes, cudaStream_t st);

}  // namespace maed
