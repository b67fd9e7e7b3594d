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

}  // namespace maed
