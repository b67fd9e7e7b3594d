// Geometry tail of the TRAINING path as two fused nodes (the inference engine's forward is decoder.cu::decode_outputs):
//   pose:    pose6d [R,144] -> rotmat [R,24,3,3] (Gram-Schmidt, reference lib/utils/geometry.py:320-334) -> angle-axis
//            (geometry.py:58-223) -> theta[:, 3:75]; backward by forward-mode differentiation (dual numbers with six partials)
//            of exactly the arithmetic of the forward kernel, branch for branch;
//   project: (kp_3d [R,J,3] | zeros, cam [R,3]) -> kp_2d [R,J,2] (reference lib/models/spin.py:113-157), forward and backward.
// Replaces ~100 elementwise torch launches per training step between the decoder and the loss.
#include "device_utils.cuh"
#include "kernels.h"

namespace maed {

namespace {

using bw::block_sum;

struct D6 { float v, d[6]; };
__device__ __forceinline__ D6 var6(float v, int k) { D6 r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = (i == k) ? 1.f : 0.f; return r; }
__device__ __forceinline__ D6 cst6(float v) { D6 r; r.v = v; for (int i = 0; i < 6; ++i) r.d[i] = 0.f; return r; }
__device__ __forceinline__ D6 operator+(D6 a, D6 b) { D6 r; r.v = a.v + b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D6 operator-(D6 a, D6 b) { D6 r; r.v = a.v - b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D6 operator-(D6 a) { D6 r; r.v = -a.v; for (int i = 0; i < 6; ++i) r.d[i] = -a.d[i]; return r; }
__device__ __forceinline__ D6 operator*(D6 a, D6 b) { D6 r; r.v = a.v * b.v; for (int i = 0; i < 6; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D6 operator/(D6 a, D6 b) {
  D6 r; r.v = a.v / b.v;
  for (int i = 0; i < 6; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
__device__ __forceinline__ D6 operator*(float s, D6 a) { D6 r; r.v = s * a.v; for (int i = 0; i < 6; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ D6 operator+(float s, D6 a) { a.v += s; return a; }
__device__ __forceinline__ D6 dsqrt(D6 a) { D6 r; r.v = sqrtf(a.v); for (int i = 0; i < 6; ++i) r.d[i] = 0.5f * a.d[i] / r.v; return r; }
// atan2(y, x): d = (x dy - y dx) / (x^2 + y^2)
__device__ __forceinline__ D6 datan2(D6 y, D6 x) {
  D6 r; r.v = atan2f(y.v, x.v);
  const float den = x.v * x.v + y.v * y.v;
  for (int i = 0; i < 6; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) / den;
  return r;
}
// x / max(|x|, 1e-6): below the clamp the denominator is a constant (torch.nn.functional.normalize semantics)
__device__ __forceinline__ D6 clamped_norm(D6 x, D6 y, D6 z) {
  const D6 n = dsqrt(x * x + y * y + z * z);
  return n.v < 1e-6f ? cst6(1e-6f) : n;
}

// one thread per (frame, joint): d_pose6d = J_R^T d_rotmat + J_aa^T d_aa
__global__ void decode_pose_bwd_kernel(const float* __restrict__ pose6d, int R, const float* __restrict__ d_rotmat,
                                       const float* __restrict__ d_aa, int ld_aa, float* __restrict__ d_pose6d) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 24) return;
  const int r = idx / 24, j = idx % 24;
  const float* v = pose6d + (long long)r * 144 + j * 6;
  // x.view(3, 2): a1 = (v0, v2, v4), a2 = (v1, v3, v5)
  const D6 a1x = var6(v[0], 0), a1y = var6(v[2], 2), a1z = var6(v[4], 4), a2x = var6(v[1], 1), a2y = var6(v[3], 3), a2z = var6(v[5], 5);
  const D6 n1 = clamped_norm(a1x, a1y, a1z);
  const D6 b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const D6 dp = b1x * a2x + b1y * a2y + b1z * a2z;
  const D6 ux = a2x - dp * b1x, uy = a2y - dp * b1y, uz = a2z - dp * b1z;
  const D6 n2 = clamped_norm(ux, uy, uz);
  const D6 b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  const D6 b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  float g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (d_rotmat) {                                          // R[i][0] = b1[i], R[i][1] = b2[i], R[i][2] = b3[i]
    const float* dR = d_rotmat + (long long)idx * 9;
    const D6* Rm[9] = {&b1x, &b2x, &b3x, &b1y, &b2y, &b3y, &b1z, &b2z, &b3z};
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
      for (int k = 0; k < 6; ++k) g[k] += dR[i] * Rm[i]->d[k];
  }
  if (d_aa) {
    // m = R^T; quaternion by the four masked cases of geometry.py:143-223, then angle-axis (:90-140)
    const D6 m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z, m20 = b3x, m21 = b3y, m22 = b3z;
    D6 qw, qx, qy, qz, t;
    if (m22.v < 1e-6f) {
      if (m00.v > m11.v) { t = 1.f + (m00 - m11 - m22); qw = m12 - m21; qx = t; qy = m01 + m10; qz = m20 + m02; }
      else               { t = 1.f + (m11 - m00 - m22); qw = m20 - m02; qx = m01 + m10; qy = t; qz = m12 + m21; }
    } else {
      if (m00.v < -m11.v) { t = 1.f + (m22 - m00 - m11); qw = m01 - m10; qx = m20 + m02; qy = m12 + m21; qz = t; }
      else                { t = 1.f + (m00 + m11 + m22); qw = t; qx = m12 - m21; qy = m20 - m02; qz = m01 - m10; }
    }
    const D6 sc = cst6(0.5f) / dsqrt(t);
    qw = qw * sc; qx = qx * sc; qy = qy * sc; qz = qz * sc;
    const D6 s2 = qx * qx + qy * qy + qz * qz;
    D6 k = cst6(2.0f);                                     // s2 == 0 (exact identity): aa = 2 q, a finite derivative
    if (s2.v > 0.0f) {
      const D6 s = dsqrt(s2);
      const D6 two_theta = 2.0f * (qw.v < 0.0f ? datan2(-s, -qw) : datan2(s, qw));
      k = two_theta / s;
    }
    const D6 aa[3] = {qx * k, qy * k, qz * k};
    const float* da = d_aa + (long long)r * ld_aa + 3 * j;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (isnan(aa[c].v)) continue;                        // the forward zeroes NaN entries: no gradient through them
#pragma unroll
      for (int k6 = 0; k6 < 6; ++k6) g[k6] += da[c] * aa[c].d[k6];
    }
  }
  float* o = d_pose6d + (long long)r * 144 + j * 6;
#pragma unroll
  for (int k = 0; k < 6; ++k) o[k] = g[k];
}

// kp_2d[r, j] = (5000 / 112) * (X / Z, Y / Z), (X, Y, Z) = kp_3d[r, j] + (cam1, cam2, 2 * 5000 / (224 * cam0 + 1e-9))
__global__ void project_fwd_kernel(const float* __restrict__ kp3d, const float* __restrict__ cam, int J, float* __restrict__ kp2d) {
  const int r = blockIdx.x;
  const float tx = cam[r * 3 + 1], ty = cam[r * 3 + 2];
  const float tz = 2.0f * 5000.0f / (224.0f * cam[r * 3 + 0] + 1e-9f);
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float X = tx, Y = ty, Z = tz;
    if (kp3d) {
      const float* p = kp3d + ((long long)r * J + j) * 3;
      X += p[0]; Y += p[1]; Z += p[2];
    }
    kp2d[((long long)r * J + j) * 2 + 0] = (5000.0f * (X / Z)) / 112.0f;
    kp2d[((long long)r * J + j) * 2 + 1] = (5000.0f * (Y / Z)) / 112.0f;
  }
}
__global__ void project_bwd_kernel(const float* __restrict__ kp3d, const float* __restrict__ cam, int J,
                                   const float* __restrict__ d_kp2d, float* __restrict__ d_cam, float* __restrict__ d_kp3d) {
  __shared__ float buf[32];
  const int r = blockIdx.x;
  const float c0 = 224.0f * cam[r * 3 + 0] + 1e-9f;
  const float tx = cam[r * 3 + 1], ty = cam[r * 3 + 2], tz = 2.0f * 5000.0f / c0;
  const float f = 5000.0f / 112.0f;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float X = tx, Y = ty, Z = tz;
    if (kp3d) {
      const float* p = kp3d + ((long long)r * J + j) * 3;
      X += p[0]; Y += p[1]; Z += p[2];
    }
    const float gx = d_kp2d[((long long)r * J + j) * 2 + 0], gy = d_kp2d[((long long)r * J + j) * 2 + 1];
    const float dX = f * gx / Z, dY = f * gy / Z, dZ = -f * (gx * X + gy * Y) / (Z * Z);
    sx += dX; sy += dY; sz += dZ;
    if (d_kp3d) {
      float* o = d_kp3d + ((long long)r * J + j) * 3;
      o[0] = dX; o[1] = dY; o[2] = dZ;
    }
  }
  sx = block_sum(sx, buf); sy = block_sum(sy, buf); sz = block_sum(sz, buf);
  if (threadIdx.x == 0) {
    d_cam[r * 3 + 0] = sz * (-2.0f * 5000.0f * 224.0f / (c0 * c0));
    d_cam[r * 3 + 1] = sx;
    d_cam[r * 3 + 2] = sy;
  }
}

}  // namespace

int decode_pose_backward(const float* pose6d, int R, const float* d_rotmat, const float* d_aa, int ld_aa, float* d_pose6d,
                         cudaStream_t st) {
  MAED_CHECK_ARG(pose6d && d_pose6d && R >= 1, "decode_pose_backward: bad argument");
  decode_pose_bwd_kernel<<<cdiv(R * 24, 64), 64, 0, st>>>(pose6d, R, d_rotmat, d_aa, ld_aa, d_pose6d);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
int project_keypoints_forward(const float* kp3d, const float* cam, int R, int J, float* kp2d, cudaStream_t st) {
  MAED_CHECK_ARG(cam && kp2d && R >= 1 && J >= 1, "project_keypoints_forward: bad argument");
  project_fwd_kernel<<<R, 64, 0, st>>>(kp3d, cam, J, kp2d);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}
int project_keypoints_backward(const float* kp3d, const float* cam, int R, int J, const float* d_kp2d, float* d_cam, float* d_kp3d,
                               cudaStream_t st) {
  MAED_CHECK_ARG(cam && d_kp2d && d_cam && R >= 1 && J >= 1, "project_keypoints_backward: bad argument");
  project_bwd_kernel<<<R, 64, 0, st>>>(kp3d, cam, J, d_kp2d, d_cam, d_kp3d);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
