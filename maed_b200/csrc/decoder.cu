// Tail of the hot path: fp32 CUDA-core kernels for the parts that are not GEMM-shaped or too small to matter —
// the KTD kinematic-tree pass, the iterative regressor's linears (linear_f32) and the rotation conversions.
// (pre_logits and the KTD fc1 / fc2 / head GEMMs run as split-precision tcgen05 GEMMs from engine.cu; they are never run
// on plain fp16 operands because they feed the outputs without a damping residual path)
// References: lib/models/ktd.py:69-124, lib/models/spin.py:51-157, lib/utils/geometry.py:58-223,320-334.
#include "kernels.h"

namespace maed {

#define LAUNCH_CHECK()                      \
  do {                                      \
    count_launch();                         \
    MAED_CUDA_CHECK(cudaGetLastError());    \
  } while (0)

// ------------------------------------------------------------------------------------- small fp32 GEMM
// out[R,N] = act(x[R,K] W[N,K]^T + bias) + residual.  Block tile 32 (rows) x 64 (cols), K tile 32,
// 256 threads, each thread 2 rows x 4 cols.
static constexpr int LBM = 32, LBN = 64, LBK = 32;
__global__ void __launch_bounds__(256)
linear_f32_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ W, int ldw,
                  const float* __restrict__ bias, int R, int N, int K, int act, const float* __restrict__ residual, int ldr,
                  float* __restrict__ out, int ldo) {
  __shared__ float xs[LBK][LBM + 1];
  __shared__ float ws[LBK][LBN + 1];
  const int r0 = blockIdx.y * LBM, n0 = blockIdx.x * LBN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;     // tx -> 4 cols, ty -> 2 rows
  float acc[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int k0 = 0; k0 < K; k0 += LBK) {
    for (int i = threadIdx.x; i < LBM * LBK; i += 256) {
      const int r = i / LBK, k = i % LBK;
      xs[k][r] = (r0 + r < R && k0 + k < K) ? x[(long long)(r0 + r) * ldx + k0 + k] : 0.f;
    }
    for (int i = threadIdx.x; i < LBN * LBK; i += 256) {
      const int n = i / LBK, k = i % LBK;
      ws[k][n] = (n0 + n < N && k0 + k < K) ? W[(long long)(n0 + n) * ldw + k0 + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < LBK; ++k) {
      const float a0 = xs[k][ty * 2], a1 = xs[k][ty * 2 + 1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = ws[k][tx * 4 + j];
        acc[0][j] += a0 * b;
        acc[1][j] += a1 * b;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = r0 + ty * 2 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (act == 3) v = tanhf(v);
      if (residual) v += residual[(long long)r * ldr + n];
      out[(long long)r * ldo + n] = v;
    }
  }
}
int linear_f32(const float* x, int ldx, const float* W, int ldw, const float* bias, int R, int N, int K, int act,
               const float* residual, int ldr, float* out, int ldo, cudaStream_t st) {
  linear_f32_kernel<<<dim3(cdiv(N, LBN), cdiv(R, LBM)), 256, 0, st>>>(x, ldx, W, ldw, bias, R, N, K, act, residual, ldr, out,
                                                                       ldo);
  LAUNCH_CHECK();
  return MAED_OK;
}

// --------------------------------------------------------------------------------------- KTD tree pass
// Kinematic ancestors (reference lib/models/ktd.py:10-35).
__constant__ int c_anc_cnt[24] = {0, 1, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8};
__constant__ int c_anc[24][8] = {
    {0}, {0}, {0}, {0}, {0, 1}, {0, 2}, {0, 3}, {0, 1, 4}, {0, 2, 5}, {0, 3, 6}, {0, 1, 4, 7}, {0, 2, 5, 8},
    {0, 3, 6, 9}, {0, 3, 6, 9}, {0, 3, 6, 9}, {0, 3, 6, 9, 12}, {0, 3, 6, 9, 13}, {0, 3, 6, 9, 14},
    {0, 3, 6, 9, 13, 16}, {0, 3, 6, 9, 14, 17}, {0, 3, 6, 9, 13, 16, 18}, {0, 3, 6, 9, 14, 17, 19},
    {0, 3, 6, 9, 13, 16, 18, 20}, {0, 3, 6, 9, 14, 17, 19, 21}};
// w_anc: for joint j, a [6][6*cnt(j)] row-major block (columns 1024.. of joint_regs.j.weight), blocks
// concatenated in joint order.  One thread per frame; weights are warp-uniform broadcast loads.
__global__ void ktd_tree_kernel(const float* __restrict__ base, int ld, const float* __restrict__ w_anc, int R,
                                float* __restrict__ pose6d, float* __restrict__ shape, float* __restrict__ cam) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  if (shape) {
#pragma unroll
    for (int i = 0; i < 10; ++i) shape[r * 10 + i] = base[(long long)r * ld + 144 + i];
#pragma unroll
    for (int i = 0; i < 3; ++i) cam[r * 3 + i] = base[(long long)r * ld + 154 + i];
  }
  float pose[144];
  int woff = 0;
  for (int j = 0; j < 24; ++j) {
    const int cnt = c_anc_cnt[j];
    float o[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) o[i] = base[(long long)r * ld + j * 6 + i];
    for (int a = 0; a < cnt; ++a) {
      const int aj = c_anc[j][a];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        const float* wr = w_anc + woff + i * 6 * cnt + a * 6;
#pragma unroll
        for (int e = 0; e < 6; ++e) o[i] += __ldg(wr + e) * pose[aj * 6 + e];
      }
    }
    woff += 36 * cnt;
#pragma unroll
    for (int i = 0; i < 6; ++i) { pose[j * 6 + i] = o[i]; pose6d[(long long)r * 144 + j * 6 + i] = o[i]; }
  }
}
int ktd_tree(const float* base, int ld, const float* w_anc, int R, float* pose6d, float* shape, float* cam,
             cudaStream_t st) {
  ktd_tree_kernel<<<cdiv(R, 32), 32, 0, st>>>(base, ld, w_anc, R, pose6d, shape, cam);
  LAUNCH_CHECK();
  return MAED_OK;
}

// ------------------------------------------------------------------------------------ output decode
// One thread per (frame, joint): rot6d -> R (Gram-Schmidt, geometry.py:320-334) -> quaternion with the
// reference's 4-way case split on the transposed matrix (geometry.py:181-222) -> angle-axis (:90-140).
__global__ void decode_pose_kernel(const float* __restrict__ pose6d, int R, float* __restrict__ rotmat,
                                   float* __restrict__ theta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * 24) return;
  const int r = idx / 24, j = idx % 24;
  const float* v = pose6d + (long long)r * 144 + j * 6;
  const float a1x = v[0], a1y = v[2], a1z = v[4], a2x = v[1], a2y = v[3], a2z = v[5];
  const float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-6f);
  const float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  const float dp = b1x * a2x + b1y * a2y + b1z * a2z;
  const float ux = a2x - dp * b1x, uy = a2y - dp * b1y, uz = a2z - dp * b1z;
  const float n2 = fmaxf(sqrtf(ux * ux + uy * uy + uz * uz), 1e-6f);
  const float b2x = ux / n2, b2y = uy / n2, b2z = uz / n2;
  const float b3x = b1y * b2z - b1z * b2y, b3y = b1z * b2x - b1x * b2z, b3z = b1x * b2y - b1y * b2x;
  // R[i][0] = b1[i], R[i][1] = b2[i], R[i][2] = b3[i]
  float* Ro = rotmat + (long long)idx * 9;
  Ro[0] = b1x; Ro[1] = b2x; Ro[2] = b3x;
  Ro[3] = b1y; Ro[4] = b2y; Ro[5] = b3y;
  Ro[6] = b1z; Ro[7] = b2z; Ro[8] = b3z;
  // m = R^T:  m[i][j] = R[j][i]
  const float m00 = b1x, m01 = b1y, m02 = b1z;
  const float m10 = b2x, m11 = b2y, m12 = b2z;
  const float m20 = b3x, m21 = b3y, m22 = b3z;
  float qw, qx, qy, qz, t;
  if (m22 < 1e-6f) {
    if (m00 > m11) { t = 1 + m00 - m11 - m22; qw = m12 - m21; qx = t; qy = m01 + m10; qz = m20 + m02; }
    else           { t = 1 - m00 + m11 - m22; qw = m20 - m02; qx = m01 + m10; qy = t; qz = m12 + m21; }
  } else {
    if (m00 < -m11) { t = 1 - m00 - m11 + m22; qw = m01 - m10; qx = m20 + m02; qy = m12 + m21; qz = t; }
    else            { t = 1 + m00 + m11 + m22; qw = t; qx = m12 - m21; qy = m20 - m02; qz = m01 - m10; }
  }
  const float sc = 0.5f / sqrtf(t);
  qw *= sc; qx *= sc; qy *= sc; qz *= sc;
  const float s2 = qx * qx + qy * qy + qz * qz;
  const float s = sqrtf(s2);
  const float two_theta = 2.0f * (qw < 0.0f ? atan2f(-s, -qw) : atan2f(s, qw));
  const float k = s2 > 0.0f ? two_theta / s : 2.0f;
  float ax = qx * k, ay = qy * k, az = qz * k;
  if (isnan(ax)) ax = 0.f;
  if (isnan(ay)) ay = 0.f;
  if (isnan(az)) az = 0.f;
  float* th = theta + (long long)r * 85 + 3 + j * 3;
  th[0] = ax; th[1] = ay; th[2] = az;
}
// theta[:, :3] = cam, theta[:, 75:] = shape; kp_2d = projection(kp_3d, cam)  (spin.py:113-157)
__global__ void decode_misc_kernel(const float* __restrict__ shape, const float* __restrict__ cam, int R,
                                   const float* __restrict__ kp3d, int n_joints, float* __restrict__ theta,
                                   float* __restrict__ kp2d) {
  const int r = blockIdx.x;
  if (threadIdx.x < 3) theta[(long long)r * 85 + threadIdx.x] = cam[r * 3 + threadIdx.x];
  if (threadIdx.x < 10) theta[(long long)r * 85 + 75 + threadIdx.x] = shape[r * 10 + threadIdx.x];
  const float tx = cam[r * 3 + 1], ty = cam[r * 3 + 2];
  const float tz = 2.0f * 5000.0f / (224.0f * cam[r * 3 + 0] + 1e-9f);
  for (int j = threadIdx.x; j < n_joints; j += blockDim.x) {
    float X = tx, Y = ty, Z = tz;
    if (kp3d) {
      const float* p = kp3d + ((long long)r * n_joints + j) * 3;
      X += p[0]; Y += p[1]; Z += p[2];
    }
    kp2d[((long long)r * n_joints + j) * 2 + 0] = (5000.0f * (X / Z)) / 112.0f;
    kp2d[((long long)r * n_joints + j) * 2 + 1] = (5000.0f * (Y / Z)) / 112.0f;
  }
}
int decode_outputs(const float* pose6d, const float* shape, const float* cam, int R, const float* kp3d, int n_joints,
                   float* rotmat, float* theta, float* kp2d, cudaStream_t st) {
  decode_pose_kernel<<<cdiv(R * 24, 128), 128, 0, st>>>(pose6d, R, rotmat, theta);
  LAUNCH_CHECK();
  decode_misc_kernel<<<R, 64, 0, st>>>(shape, cam, R, kp3d, n_joints, theta, kp2d);
  LAUNCH_CHECK();
  return MAED_OK;
}

__global__ void concat_cols_kernel(const float* a, int ca, const float* b, int cb, const float* c, int cc, const float* d,
                                   int cd, int R, float* out) {
  const int W = ca + cb + cc + cd;
  const long long total = (long long)R * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % W);
    const long long r = i / W;
    float v;
    if (col < ca) v = a[r * ca + col];
    else if (col < ca + cb) v = b[r * cb + col - ca];
    else if (col < ca + cb + cc) v = c[r * cc + col - ca - cb];
    else v = d[r * cd + col - ca - cb - cc];
    out[i] = v;
  }
}
int concat_cols(const float* a, int ca, const float* b, int cb, const float* c, int cc, const float* d, int cd, int R,
                float* out, cudaStream_t st) {
  const long long total = (long long)R * (ca + cb + cc + cd);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 1024) blocks = 1024;
  concat_cols_kernel<<<blocks, 256, 0, st>>>(a, ca, b, cb, c, cc, d, cd, R, out);
  LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
