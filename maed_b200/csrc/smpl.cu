// SMPL body model forward (linear blend skinning) + joint regression: the `self.smpl(...)` call of the decoders
// (reference lib/models/ktd.py:100-114, lib/models/spin.py:92-104 -> lib/models/smpl.py:84-106 -> smplx.SMPL.forward /
// smplx.lbs.lbs).  `smplx==0.1.13` (requirements.txt:4) is not vendored under the reference and not installed, and the
// SMPL assets are licensed data: this file restates the PUBLISHED algorithm of smplx.lbs.lbs (shape blend shapes, joint
// regression, pose blend shapes, kinematic chain, skinning, vertex-selected joints) and is checked against
// oracle/smpl_oracle.py on a seeded synthetic asset pack — parity with smplx itself is UNPINNED.
//
// All HBM / L2-bound, fp32:
//   smpl_chain_kernel : per frame: J = J_template + J_shapedirs beta; 24-joint kinematic chain; A_j = [G_j | t_j - G_j J_j];
//                       pose feature (R_j - I, j = 1..23)
//   smpl_blend_kernel : v_posed = v_template + shapedirs beta + posedirs^T pose_feature          (20 670 values per frame)
//   smpl_skin_kernel  : verts = (sum_j w_vj A_j) [v_posed; 1]
//   smpl_joints_kernel: 24 chain joints + 21 vertex-selected + 9 regressed (J_regressor_extra) -> joint_map -> 49,
//                       or J_regressor (<= 17 rows) @ verts when the caller passes one (ktd.py:110-112)
#include "smpl.h"

#include "device_utils.cuh"

namespace maed {
using namespace bw;

static constexpr int kNV = 6890, kNJ = 24, kNE = kNV * 3, kPF = 207;

__global__ void smpl_chain_kernel(const float* __restrict__ betas, const float* __restrict__ rot, const float* __restrict__ Jt,
                                  const float* __restrict__ Jsd, const int* __restrict__ parents, int BT,
                                  float* __restrict__ A, float* __restrict__ jpos, float* __restrict__ pf) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= BT) return;
  float beta[10];
#pragma unroll
  for (int l = 0; l < 10; ++l) beta[l] = betas[(long long)b * 10 + l];
  float J[kNJ][3];
  for (int j = 0; j < kNJ; ++j)
    for (int c = 0; c < 3; ++c) {
      float s = Jt[j * 3 + c];
#pragma unroll
      for (int l = 0; l < 10; ++l) s += Jsd[(j * 3 + c) * 10 + l] * beta[l];
      J[j][c] = s;
    }
  const float* R = rot + (long long)b * kNJ * 9;
  float G[kNJ][12];                                    // global transform rows: [r00 r01 r02 t0 | r10 .. t1 | r20 .. t2]
  for (int j = 0; j < kNJ; ++j) {
    const float* Rj = R + j * 9;
    const int pa = parents[j];
    float rel[3];
    for (int c = 0; c < 3; ++c) rel[c] = J[j][c] - (pa >= 0 ? J[pa][c] : 0.f);
    if (pa < 0) {
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) G[j][r * 4 + c] = Rj[r * 3 + c];
        G[j][r * 4 + 3] = rel[r];
      }
    } else {
      for (int r = 0; r < 3; ++r) {
        const float g0 = G[pa][r * 4], g1 = G[pa][r * 4 + 1], g2 = G[pa][r * 4 + 2];
        for (int c = 0; c < 3; ++c) G[j][r * 4 + c] = g0 * Rj[c] + g1 * Rj[3 + c] + g2 * Rj[6 + c];
        G[j][r * 4 + 3] = g0 * rel[0] + g1 * rel[1] + g2 * rel[2] + G[pa][r * 4 + 3];
      }
    }
    if (j >= 1)
      for (int e = 0; e < 9; ++e) pf[(long long)b * kPF + (j - 1) * 9 + e] = Rj[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
  }
  for (int j = 0; j < kNJ; ++j) {
    float* Aj = A + ((long long)b * kNJ + j) * 12;
    for (int r = 0; r < 3; ++r) {
      const float g0 = G[j][r * 4], g1 = G[j][r * 4 + 1], g2 = G[j][r * 4 + 2], t = G[j][r * 4 + 3];
      Aj[r * 4] = g0; Aj[r * 4 + 1] = g1; Aj[r * 4 + 2] = g2;
      Aj[r * 4 + 3] = t - (g0 * J[j][0] + g1 * J[j][1] + g2 * J[j][2]);
      jpos[((long long)b * kNJ + j) * 3 + r] = t;
    }
  }
}

// grid (ceil(20670 / 256), BT)
__global__ void smpl_blend_kernel(const float* __restrict__ betas, const float* __restrict__ pf, const float* __restrict__ vt,
                                  const float* __restrict__ sd, const float* __restrict__ pd, float* __restrict__ vposed) {
  __shared__ float s_pf[kPF];
  __shared__ float s_beta[10];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kPF; i += blockDim.x) s_pf[i] = pf[(long long)b * kPF + i];
  if (threadIdx.x < 10) s_beta[threadIdx.x] = betas[(long long)b * 10 + threadIdx.x];
  __syncthreads();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= kNE) return;
  float s = vt[e];
#pragma unroll
  for (int l = 0; l < 10; ++l) s += sd[(long long)e * 10 + l] * s_beta[l];
  for (int k = 0; k < kPF; ++k) s += s_pf[k] * pd[(long long)k * kNE + e];
  vposed[(long long)b * kNE + e] = s;
}

// grid (ceil(6890 / 128), BT)
__global__ void smpl_skin_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ vposed,
                                 float* __restrict__ verts) {
  __shared__ float s_A[kNJ * 12];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kNJ * 12; i += blockDim.x) s_A[i] = A[(long long)b * kNJ * 12 + i];
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= kNV) return;
  float T[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = 0.f;
  for (int j = 0; j < kNJ; ++j) {
    const float w = W[(long long)v * kNJ + j];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] += w * s_A[j * 12 + i];
  }
  const float* p = vposed + ((long long)b * kNV + v) * 3;
  const float x = p[0], y = p[1], z = p[2];
  float* o = verts + ((long long)b * kNV + v) * 3;
  o[0] = T[0] * x + T[1] * y + T[2] * z + T[3];
  o[1] = T[4] * x + T[5] * y + T[6] * z + T[7];
  o[2] = T[8] * x + T[9] * y + T[10] * z + T[11];
}

// One block per frame.  nreg rows of `reg` ([nreg, 6890]) are applied to the frame's vertices (block reduction), then:
//   mode 0: out[49] = joint_map( 24 chain joints | 21 vertex-selected | 9 regressed )      (smpl.py:97-100)
//   mode 1: out[nreg] = reg @ verts                                                          (ktd.py:110-112)
static constexpr int kMaxReg = 17;
__global__ void __launch_bounds__(256)
smpl_joints_kernel(const float* __restrict__ verts, const float* __restrict__ jpos, const float* __restrict__ reg, int nreg,
                   const int* __restrict__ extra_ids, const int* __restrict__ joint_map, int mode, int n_out,
                   float* __restrict__ out) {
  __shared__ float s_part[8][kMaxReg * 3];
  __shared__ float s_j[54 * 3];
  const int b = blockIdx.x;
  const float* vb = verts + (long long)b * kNE;
  float acc[kMaxReg * 3];
#pragma unroll
  for (int i = 0; i < kMaxReg * 3; ++i) acc[i] = 0.f;
  for (int v = threadIdx.x; v < kNV; v += blockDim.x) {
    const float x = vb[v * 3], y = vb[v * 3 + 1], z = vb[v * 3 + 2];
#pragma unroll
    for (int k = 0; k < kMaxReg; ++k)
      if (k < nreg) {
        const float w = reg[(long long)k * kNV + v];
        acc[k * 3] += w * x; acc[k * 3 + 1] += w * y; acc[k * 3 + 2] += w * z;
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kMaxReg * 3; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) s_part[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < nreg * 3) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_part[w][threadIdx.x];
    if (mode == 1) out[((long long)b * n_out) * 3 + threadIdx.x] = s;
    else s_j[45 * 3 + threadIdx.x] = s;
  }
  if (mode == 0) {
    if (threadIdx.x < 24 * 3) s_j[threadIdx.x] = jpos[(long long)b * 72 + threadIdx.x];
    if (threadIdx.x >= 96 && threadIdx.x < 96 + 21 * 3) {
      const int t = threadIdx.x - 96;
      s_j[24 * 3 + t] = vb[extra_ids[t / 3] * 3 + t % 3];
    }
    __syncthreads();
    if (threadIdx.x < n_out * 3) out[((long long)b * n_out) * 3 + threadIdx.x] = s_j[joint_map[threadIdx.x / 3] * 3 + threadIdx.x % 3];
  }
}

size_t smpl_scratch_bytes(int BT) {
  return ((size_t)BT * kNJ * 12 + (size_t)BT * kNJ * 3 + (size_t)BT * kPF + (size_t)BT * kNE) * sizeof(float) + 1024;
}

int smpl_forward(const SmplAssets* a, const float* betas, const float* rotmat, int BT, const float* J_regressor, int n_reg,
                 float* verts, float* joints, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  MAED_CHECK_ARG(a && betas && rotmat && verts && joints && scratch, "smpl_forward: null argument");
  MAED_CHECK_ARG(BT >= 1, "smpl_forward: BT=%d", BT);
  MAED_CHECK_ARG(scratch_bytes >= smpl_scratch_bytes(BT), "smpl_forward: scratch too small");
  MAED_CHECK_ARG(!J_regressor || (n_reg >= 1 && n_reg <= kMaxReg), "smpl_forward: J_regressor rows %d unsupported (1..%d)", n_reg,
                 kMaxReg);
  float* A = (float*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  float* jpos = A + (size_t)BT * kNJ * 12;
  float* pf = jpos + (size_t)BT * kNJ * 3;
  float* vposed = pf + (size_t)BT * kPF;
  smpl_chain_kernel<<<cdiv(BT, 64), 64, 0, st>>>(betas, rotmat, a->J_template, a->J_shapedirs, a->parents, BT, A, jpos, pf);
  MAED_BW_LAUNCH_CHECK();
  smpl_blend_kernel<<<dim3(cdiv(kNE, 256), BT), 256, 0, st>>>(betas, pf, a->v_template, a->shapedirs, a->posedirs, vposed);
  MAED_BW_LAUNCH_CHECK();
  smpl_skin_kernel<<<dim3(cdiv(kNV, 128), BT), 128, 0, st>>>(A, a->lbs_weights, vposed, verts);
  MAED_BW_LAUNCH_CHECK();
  if (J_regressor)
    smpl_joints_kernel<<<BT, 256, 0, st>>>(verts, jpos, J_regressor, n_reg, a->extra_vertex_ids, a->joint_map, 1, n_reg, joints);
  else
    smpl_joints_kernel<<<BT, 256, 0, st>>>(verts, jpos, a->J_regressor_extra, 9, a->extra_vertex_ids, a->joint_map, 0, 49, joints);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ====================================================================================================== backward
// Gradients of (verts, joints) with respect to (betas, rotmat): what the keypoint losses of the reference need
// (lib/core/loss.py:178-192 act on kp_2d / kp_3d = joints of the body model, ktd.py:100-114).  The forward intermediates
// (A, vposed) are recomputed into the scratch; every reduction is a fixed-order two-stage sum (bit-reproducible).
//   smpl_joints_bwd_kernel : d_joints (49 mapped | n_reg regressed) -> gradient of the 24 chain joints + contribution to d_verts
//   smpl_skin_bwd_kernel   : d_verts -> d_vposed and per-block partials of dA_j = sum_v w_vj (dv (x) [vposed; 1])
//   smpl_blend_bwd_kernel  : d_pose_feature = posedirs d_vposed, d_beta (blend part) = shapedirs^T d_vposed
//   smpl_chain_bwd_kernel  : dA, d_jpos, d_pose_feature -> d_rotmat, d_betas (kinematic chain walked leaves to root)
static constexpr int kSkinBlk = 128, kSkinBlocks = (kNV + kSkinBlk - 1) / kSkinBlk;      // 54 blocks per frame

// one block per frame: dvt[b] = d_verts[b] (or 0) + reg^T d_reg + vertex-selected joints; d_jpos[b, 24, 3]
__global__ void __launch_bounds__(256)
smpl_joints_bwd_kernel(const float* __restrict__ d_verts, const float* __restrict__ d_joints, const float* __restrict__ reg,
                       int nreg, const int* __restrict__ extra_ids, const int* __restrict__ joint_map, int mode, int n_out,
                       float* __restrict__ dvt, float* __restrict__ d_jpos) {
  __shared__ float s_d54[(45 + kMaxReg) * 3];             // 54 joints (mode 0) or 45 unused slots + up to 17 regressed rows
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < (45 + kMaxReg) * 3; i += blockDim.x) s_d54[i] = 0.f;
  __syncthreads();
  if (mode == 0) {
    // joint_map repeats indices (e.g. hips, knees, wrists appear twice): one thread per (joint of the 54, coordinate) sums its
    // occurrences in a fixed order
    if (threadIdx.x < 54 * 3) {
      const int j = threadIdx.x / 3, c = threadIdx.x % 3;
      float s = 0.f;
      for (int o = 0; o < n_out; ++o)
        if (joint_map[o] == j) s += d_joints[((long long)b * n_out + o) * 3 + c];
      s_d54[threadIdx.x] = s;
    }
  } else if (threadIdx.x < nreg * 3) {
    s_d54[45 * 3 + threadIdx.x] = d_joints[(long long)b * n_out * 3 + threadIdx.x];     // regressed rows live at slot 45..
  }
  __syncthreads();
  if (threadIdx.x < 72) d_jpos[(long long)b * 72 + threadIdx.x] = mode == 0 ? s_d54[threadIdx.x] : 0.f;
  float* ob = dvt + (long long)b * kNE;
  for (int v = threadIdx.x; v < kNV; v += blockDim.x) {
    float x = 0.f, y = 0.f, z = 0.f;
    if (d_verts) { const float* p = d_verts + ((long long)b * kNV + v) * 3; x = p[0]; y = p[1]; z = p[2]; }
    for (int k = 0; k < nreg; ++k) {
      const float w = reg[(long long)k * kNV + v];
      x += w * s_d54[(45 + k) * 3]; y += w * s_d54[(45 + k) * 3 + 1]; z += w * s_d54[(45 + k) * 3 + 2];
    }
    ob[v * 3] = x; ob[v * 3 + 1] = y; ob[v * 3 + 2] = z;
  }
  __syncthreads();
  if (mode == 0 && threadIdx.x == 0)                     // 21 vertex-selected joints (ids may repeat: serial, fixed order)
    for (int t = 0; t < 21; ++t)
      for (int c = 0; c < 3; ++c) ob[extra_ids[t] * 3 + c] += s_d54[(24 + t) * 3 + c];
}

// grid (kSkinBlocks, BT), 128 threads
__global__ void __launch_bounds__(kSkinBlk)
smpl_skin_bwd_kernel(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ vposed,
                     const float* __restrict__ dvt, float* __restrict__ d_vposed, float* __restrict__ dA_part) {
  __shared__ float s_A[kNJ * 12];
  __shared__ float s_red[kSkinBlk / 32][12];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < kNJ * 12; i += blockDim.x) s_A[i] = A[(long long)b * kNJ * 12 + i];
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = v < kNV;
  float dT[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) dT[i] = 0.f;
  if (ok) {
    float T[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) T[i] = 0.f;
    for (int j = 0; j < kNJ; ++j) {
      const float w = W[(long long)v * kNJ + j];
#pragma unroll
      for (int i = 0; i < 12; ++i) T[i] += w * s_A[j * 12 + i];
    }
    const float* g = dvt + ((long long)b * kNV + v) * 3;
    const float* p = vposed + ((long long)b * kNV + v) * 3;
    const float gx = g[0], gy = g[1], gz = g[2], x = p[0], y = p[1], z = p[2];
    float* o = d_vposed + ((long long)b * kNV + v) * 3;            // d vposed = T_R^T dv
    o[0] = T[0] * gx + T[4] * gy + T[8] * gz;
    o[1] = T[1] * gx + T[5] * gy + T[9] * gz;
    o[2] = T[2] * gx + T[6] * gy + T[10] * gz;
    dT[0] = gx * x; dT[1] = gx * y; dT[2] = gx * z; dT[3] = gx;     // dT = dv (x) [vposed; 1]
    dT[4] = gy * x; dT[5] = gy * y; dT[6] = gy * z; dT[7] = gy;
    dT[8] = gz * x; dT[9] = gz * y; dT[10] = gz * z; dT[11] = gz;
  }
  // dA_j partial of this block = sum over its vertices of w_vj dT_v: warp shuffles, then the 4 warps in a fixed order
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* out = dA_part + ((long long)b * kSkinBlocks + blockIdx.x) * kNJ * 12;
  for (int j = 0; j < kNJ; ++j) {
    const float w = ok ? W[(long long)v * kNJ + j] : 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float s = warp_sum(w * dT[i]);
      if (lane == 0) s_red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
      float s = 0.f;
      for (int k = 0; k < kSkinBlk / 32; ++k) s += s_red[k][threadIdx.x];
      out[j * 12 + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

// grid (207 + 10, ceil(BT / 8)), 256 threads: out[b, k] = sum_e d_vposed[b, e] * w_k[e]; w_k = posedirs[k, :] | shapedirs[:, k - 207]
__global__ void __launch_bounds__(256)
smpl_blend_bwd_kernel(const float* __restrict__ d_vposed, const float* __restrict__ pd, const float* __restrict__ sd, int BT,
                      float* __restrict__ d_pf, float* __restrict__ d_beta_blend) {
  __shared__ float s_red[8][8];
  const int k = blockIdx.x, b0 = blockIdx.y * 8;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int e = threadIdx.x; e < kNE; e += blockDim.x) {
    const float w = k < kPF ? pd[(long long)k * kNE + e] : sd[(long long)e * 10 + (k - kPF)];
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (b0 + i < BT) acc[i] += w * d_vposed[(long long)(b0 + i) * kNE + e];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) s_red[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8 && b0 + threadIdx.x < BT) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += s_red[w][threadIdx.x];
    if (k < kPF) d_pf[(long long)(b0 + threadIdx.x) * kPF + k] = s;
    else d_beta_blend[(long long)(b0 + threadIdx.x) * 10 + (k - kPF)] = s;
  }
}

// one thread per frame
__global__ void smpl_chain_bwd_kernel(const float* __restrict__ betas, const float* __restrict__ rot, const float* __restrict__ Jt,
                                      const float* __restrict__ Jsd, const int* __restrict__ parents, int BT,
                                      const float* __restrict__ dA_part, const float* __restrict__ d_jpos,
                                      const float* __restrict__ d_pf, const float* __restrict__ d_beta_blend,
                                      float* __restrict__ d_rot, float* __restrict__ d_betas) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= BT) return;
  float beta[10];
#pragma unroll
  for (int l = 0; l < 10; ++l) beta[l] = betas[(long long)b * 10 + l];
  float J[kNJ][3], GR[kNJ][9], dGR[kNJ][9], dGt[kNJ][3], dJ[kNJ][3];
  for (int j = 0; j < kNJ; ++j)
    for (int c = 0; c < 3; ++c) {
      float s = Jt[j * 3 + c];
#pragma unroll
      for (int l = 0; l < 10; ++l) s += Jsd[(j * 3 + c) * 10 + l] * beta[l];
      J[j][c] = s;
    }
  const float* R = rot + (long long)b * kNJ * 9;
  for (int j = 0; j < kNJ; ++j) {                         // forward: global rotations only (translations are not needed)
    const float* Rj = R + j * 9;
    const int pa = parents[j];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c)
        GR[j][r * 3 + c] = pa < 0 ? Rj[r * 3 + c]
                                  : GR[pa][r * 3] * Rj[c] + GR[pa][r * 3 + 1] * Rj[3 + c] + GR[pa][r * 3 + 2] * Rj[6 + c];
  }
  // dA (sum of the per-block partials, fixed order) -> dG: A_R = G_R, A_t = G_t - G_R J
  for (int j = 0; j < kNJ; ++j) {
    float dA[12];
    for (int i = 0; i < 12; ++i) {
      float s = 0.f;
      for (int k = 0; k < kSkinBlocks; ++k) s += dA_part[(((long long)b * kSkinBlocks + k) * kNJ + j) * 12 + i];
      dA[i] = s;
    }
    for (int r = 0; r < 3; ++r) {
      const float dt = dA[r * 4 + 3];
      for (int c = 0; c < 3; ++c) dGR[j][r * 3 + c] = dA[r * 4 + c] - dt * J[j][c];
      dGt[j][r] = dt + d_jpos[((long long)b * kNJ + j) * 3 + r];
    }
    for (int c = 0; c < 3; ++c)
      dJ[j][c] = -(GR[j][c] * dA[3] + GR[j][3 + c] * dA[7] + GR[j][6 + c] * dA[11]);
  }
  float* dR = d_rot + (long long)b * kNJ * 9;
  for (int j = kNJ - 1; j >= 0; --j) {
    const float* Rj = R + j * 9;
    const int pa = parents[j];
    if (pa < 0) {
      for (int e = 0; e < 9; ++e) dR[j * 9 + e] = dGR[j][e];
      for (int c = 0; c < 3; ++c) dJ[j][c] += dGt[j][c];
      continue;
    }
    float rel[3], drel[3];
    for (int c = 0; c < 3; ++c) rel[c] = J[j][c] - J[pa][c];
    for (int a = 0; a < 3; ++a) {
      for (int c = 0; c < 3; ++c)                                   // dR_j = G_R_p^T dG_R_j (+ pose-feature gradient)
        dR[j * 9 + a * 3 + c] = GR[pa][a] * dGR[j][c] + GR[pa][3 + a] * dGR[j][3 + c] + GR[pa][6 + a] * dGR[j][6 + c] +
                                d_pf[(long long)b * kPF + (j - 1) * 9 + a * 3 + c];
      drel[a] = GR[pa][a] * dGt[j][0] + GR[pa][3 + a] * dGt[j][1] + GR[pa][6 + a] * dGt[j][2];
    }
    for (int r = 0; r < 3; ++r) {
      for (int a = 0; a < 3; ++a)                                   // dG_R_p += dG_R_j R_j^T + dG_t_j (x) rel_j
        dGR[pa][r * 3 + a] += dGR[j][r * 3] * Rj[a * 3] + dGR[j][r * 3 + 1] * Rj[a * 3 + 1] + dGR[j][r * 3 + 2] * Rj[a * 3 + 2] +
                              dGt[j][r] * rel[a];
      dGt[pa][r] += dGt[j][r];
    }
    for (int c = 0; c < 3; ++c) { dJ[j][c] += drel[c]; dJ[pa][c] -= drel[c]; }
  }
  for (int l = 0; l < 10; ++l) {
    float s = d_beta_blend[(long long)b * 10 + l];
    for (int j = 0; j < kNJ; ++j)
      for (int c = 0; c < 3; ++c) s += Jsd[(j * 3 + c) * 10 + l] * dJ[j][c];
    d_betas[(long long)b * 10 + l] = s;
  }
}

size_t smpl_backward_scratch_bytes(int BT) {
  return smpl_scratch_bytes(BT) + ((size_t)2 * BT * kNE + (size_t)BT * kSkinBlocks * kNJ * 12 + (size_t)BT * (72 + kPF + 10)) *
                                      sizeof(float) + 1024;
}

int smpl_backward(const SmplAssets* a, const float* betas, const float* rotmat, int BT, const float* J_regressor, int n_reg,
                  const float* d_verts, const float* d_joints, float* d_betas, float* d_rotmat, void* scratch,
                  size_t scratch_bytes, cudaStream_t st) {
  MAED_CHECK_ARG(a && betas && rotmat && d_joints && d_betas && d_rotmat && scratch, "smpl_backward: null argument");
  MAED_CHECK_ARG(BT >= 1, "smpl_backward: BT=%d", BT);
  MAED_CHECK_ARG(scratch_bytes >= smpl_backward_scratch_bytes(BT), "smpl_backward: scratch too small");
  MAED_CHECK_ARG(!J_regressor || (n_reg >= 1 && n_reg <= kMaxReg), "smpl_backward: J_regressor rows %d unsupported (1..%d)", n_reg,
                 kMaxReg);
  float* A = (float*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
  float* jpos = A + (size_t)BT * kNJ * 12;
  float* pf = jpos + (size_t)BT * kNJ * 3;
  float* vposed = pf + (size_t)BT * kPF;
  float* dvt = vposed + (size_t)BT * kNE;
  float* d_vposed = dvt + (size_t)BT * kNE;
  float* dA_part = d_vposed + (size_t)BT * kNE;
  float* d_jpos = dA_part + (size_t)BT * kSkinBlocks * kNJ * 12;
  float* d_pf = d_jpos + (size_t)BT * 72;
  float* d_beta_blend = d_pf + (size_t)BT * kPF;
  // forward intermediates (A, vposed) again
  smpl_chain_kernel<<<cdiv(BT, 64), 64, 0, st>>>(betas, rotmat, a->J_template, a->J_shapedirs, a->parents, BT, A, jpos, pf);
  MAED_BW_LAUNCH_CHECK();
  smpl_blend_kernel<<<dim3(cdiv(kNE, 256), BT), 256, 0, st>>>(betas, pf, a->v_template, a->shapedirs, a->posedirs, vposed);
  MAED_BW_LAUNCH_CHECK();
  if (J_regressor)
    smpl_joints_bwd_kernel<<<BT, 256, 0, st>>>(d_verts, d_joints, J_regressor, n_reg, a->extra_vertex_ids, a->joint_map, 1, n_reg, dvt,
                                               d_jpos);
  else
    smpl_joints_bwd_kernel<<<BT, 256, 0, st>>>(d_verts, d_joints, a->J_regressor_extra, 9, a->extra_vertex_ids, a->joint_map, 0, 49,
                                               dvt, d_jpos);
  MAED_BW_LAUNCH_CHECK();
  smpl_skin_bwd_kernel<<<dim3(kSkinBlocks, BT), kSkinBlk, 0, st>>>(A, a->lbs_weights, vposed, dvt, d_vposed, dA_part);
  MAED_BW_LAUNCH_CHECK();
  smpl_blend_bwd_kernel<<<dim3(kPF + 10, cdiv(BT, 8)), 256, 0, st>>>(d_vposed, a->posedirs, a->shapedirs, BT, d_pf, d_beta_blend);
  MAED_BW_LAUNCH_CHECK();
  smpl_chain_bwd_kernel<<<cdiv(BT, 32), 32, 0, st>>>(betas, rotmat, a->J_template, a->J_shapedirs, a->parents, BT, dA_part, d_jpos,
                                                     d_pf, d_beta_blend, d_rotmat, d_betas);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
