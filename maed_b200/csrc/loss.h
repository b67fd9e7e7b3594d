// Fused training loss + gradients with respect to the predictions (loss.cu; reference lib/core/loss.py).
#pragma once
#include "common.h"

namespace maed {

struct LossWeights { float kp2d, kp3d, pose, shape, norm, accl; };

size_t loss_scratch_bytes(int M2, int M3);
// pred_kp2d [M2,J2,2] / gt_kp2d [M2,J2,3] (x, y, conf) or NULL; pred_kp3d [M3,J3,3] / gt_kp3d [M3,J3,4] or NULL;
// theta [M3,85]; valid [M3] bytes or NULL (all valid); T = frames per clip (M3 % T == 0; only the accl term uses it).
// losses[8] = kp2d, kp3d, shape, pose, norm, accl (weighted), total, n_valid.  d_* = d total / d pred_*.
int loss_forward_backward(const float* pred_kp2d, const float* gt_kp2d, int M2, int J2, const float* pred_kp3d,
                          const float* gt_kp3d, int M3, int J3, const float* pred_theta, const float* gt_theta,
                          const unsigned char* valid, int T, const LossWeights* w, float* losses, float* d_kp2d, float* d_kp3d,
                          float* d_theta, void* scratch, size_t scratch_bytes, cudaStream_t st);

}  // namespace maed
