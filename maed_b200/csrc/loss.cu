// Fused training loss of MAED (reference lib/core/loss.py:21-117 terms, :159-210 LossVideo, :214-283 LossImage): every
// term AND its gradient with respect to the predictions in three small launches, no host synchronisation
// (the reference runs ~60 elementwise kernels and syncs the host with .item() per iteration, lib/core/trainer.py:209-226).
//
//   partial   grid = frames : per-frame sums of every term (fixed-order block reductions)
//   finalize  1 block       : sums over frames in double, fixed order -> the six weighted terms, their total, n_valid
//   grads     grid = frames : d total / d kp_2d, d kp_3d, d theta using the global scalars (n_valid, the theta norm)
//
// Terms (weights w_*):
//   kp2d  : mean_{M2,J,2}  conf * (pred - gt)^2                                             loss.py:21-38
//   kp3d  : mean_{M3,J,3}  conf * ((pred - pelvis_pred) - (gt - pelvis_gt))^2               loss.py:40-62
//           pelvis = mean of joints 27 and 28
//   pose  : mean_{valid,24,9} (R(pred_aa) - R(gt_aa))^2, R = batch_rodrigues                loss.py:64-93, geometry.py:12-58
//   shape : mean_{valid,10} (pred - gt)^2
//   norm  : ||theta[:, 3:]||_F / M3                                                         loss.py:203
//   accl  : mean_{N,T-2,J,3} (conf_a * (accl_pred - accl_gt))^2, conf_a = conf[t+2]^4       loss.py:95-117
// The derivative of batch_rodrigues is taken by forward-mode differentiation (dual numbers with three partials) of the
// very formula the reference evaluates, so that it follows the reference's autograd also at its singular points
// (|aa| -> 0, where the reference divides by |aa + 1e-8|).
#include "device_utils.cuh"
#include "loss.h"

namespace maed {

namespace {

using bw::block_sum;

struct D3 { float v, d[3]; };
__device__ __forceinline__ D3 mk(float v, int k) { D3 r; r.v = v; r.d[0] = r.d[1] = r.d[2] = 0.f; if (k >= 0) r.d[k] = 1.f; return r; }
__device__ __forceinline__ D3 operator+(D3 a, D3 b) { D3 r; r.v = a.v + b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
__device__ __forceinline__ D3 operator-(D3 a, D3 b) { D3 r; r.v = a.v - b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
__device__ __forceinline__ D3 operator*(D3 a, D3 b) { D3 r; r.v = a.v * b.v; for (int i = 0; i < 3; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
__device__ __forceinline__ D3 operator/(D3 a, D3 b) {
  D3 r; r.v = a.v / b.v;
  for (int i = 0; i < 3; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v;
  return r;
}
__device__ __forceinline__ D3 operator*(float s, D3 a) { D3 r; r.v = s * a.v; for (int i = 0; i < 3; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ D3 operator+(D3 a, float s) { a.v += s; return a; }
__device__ __forceinline__ D3 tsqrt(D3 a) { D3 r; r.v = sqrtf(a.v); for (int i = 0; i < 3; ++i) r.d[i] = 0.5f * a.d[i] / r.v; return r; }
__device__ __forceinline__ D3 tsin(D3 a) { D3 r; const float c = cosf(a.v); r.v = sinf(a.v); for (int i = 0; i < 3; ++i) r.d[i] = c * a.d[i]; return r; }
__device__ __forceinline__ D3 tcos(D3 a) { D3 r; const float s = -sinf(a.v); r.v = cosf(a.v); for (int i = 0; i < 3; ++i) r.d[i] = s * a.d[i]; return r; }
__device__ __forceinline__ float tsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ float tsin(float a) { return sinf(a); }
__device__ __forceinline__ float tcos(float a) { return cosf(a); }

// geometry.py:12-24 batch_rodrigues + :27-58 quat2mat, one joint; R row-major
template <class S>
__device__ __forceinline__ void rodrigues(S ax, S ay, S az, S* R) {
  const float e = 1e-8f;
  const S sx = ax + e, sy = ay + e, sz = az + e;
  const S n = tsqrt(sx * sx + sy * sy + sz * sz);            // norm(axisang + 1e-8)
  const S half = 0.5f * n;
  const S c = tcos(half), s = tsin(half);
  S w = c, x = s * (ax / n), y = s * (ay / n), z = s * (az / n);
  const S qn = tsqrt(w * w + x * x + y * y + z * z);         // quat2mat normalises again
  w = w / qn; x = x / qn; y = y / qn; z = z / qn;
  const S w2 = w * w, x2 = x * x, y2 = y * y, z2 = z * z;
  const S wx = w * x, wy = w * y, wz = w * z, xy = x * y, xz = x * z, yz = y * z;
  R[0] = w2 + x2 - y2 - z2; R[1] = 2.f * xy - 2.f * wz;    R[2] = 2.f * wy + 2.f * xz;
  R[3] = 2.f * wz + 2.f * xy; R[4] = w2 - x2 + y2 - z2;    R[5] = 2.f * yz - 2.f * wx;
  R[6] = 2.f * xz - 2.f * wy; R[7] = 2.f * wx + 2.f * yz;  R[8] = w2 - x2 - y2 + z2;
}

struct LossArgs {
  const float* p2; const float* g2; int M2, J2;
  const float* p3; const float* g3; int M3, J3, pel_a, pel_b;
  const float* pt; const float* gt; const unsigned char* valid;
  int T;
  LossWeights w;
  float* partial;          // [frames][kPartial]
  float* scalars;          // kScalars floats written by finalize (device copy of `losses` + helpers)
  float* losses;           // [8] output
  float* d2; float* d3; float* dt;
  int frames;
};
constexpr int kPartial = 8;   // kp2d, kp3d, pose, shape, normsq, accl, valid, unused
constexpr int kThreads = 128;

__device__ __forceinline__ bool frame_valid(const LossArgs& a, int f) { return a.valid ? a.valid[f] != 0 : true; }

// second difference of joint coordinate c at window s of clip frame range (pred or gt with row stride `ld`)
__device__ __forceinline__ float accl_at(const float* base, int ld, int J, int f, int j, int c) {
  const float* p = base + ((long long)f * J + j) * ld + c;
  const long long step = (long long)J * ld;
  return p[2 * step] - 2.f * p[step] + p[0];
}

__global__ void loss_partial_kernel(const LossArgs a) {
  __shared__ float buf[32];
  const int f = blockIdx.x;
  float s2 = 0.f, s3 = 0.f, sp = 0.f, ss = 0.f, sn = 0.f, sa = 0.f;
  if (f < a.M2 && a.g2) {
    for (int i = threadIdx.x; i < a.J2 * 2; i += blockDim.x) {
      const int j = i >> 1, c = i & 1;
      const float* g = a.g2 + ((long long)f * a.J2 + j) * 3;
      const float d = a.p2[((long long)f * a.J2 + j) * 2 + c] - g[c];
      s2 += g[2] * d * d;
    }
  }
  if (f < a.M3) {
    if (a.g3) {
      const float* P = a.p3 + (long long)f * a.J3 * 3;
      const float* G = a.g3 + (long long)f * a.J3 * 4;
      for (int i = threadIdx.x; i < a.J3 * 3; i += blockDim.x) {
        const int j = i / 3, c = i % 3;
        const float pp = 0.5f * (P[a.pel_a * 3 + c] + P[a.pel_b * 3 + c]);
        const float pg = 0.5f * (G[a.pel_a * 4 + c] + G[a.pel_b * 4 + c]);
        const float d = (P[j * 3 + c] - pp) - (G[j * 4 + c] - pg);
        s3 += G[j * 4 + 3] * d * d;
      }
      if (a.w.accl > 0.f && a.T >= 3 && (f % a.T) + 2 < a.T) {        // window starting at this frame
        for (int i = threadIdx.x; i < a.J3 * 3; i += blockDim.x) {
          const int j = i / 3, c = i % 3;
          const float cf = G[((long long)2 * a.J3 + j) * 4 + 3];      // conf of frame f + 2
          const float c4 = cf * cf * cf * cf;
          const float d = c4 * (accl_at(a.p3, 3, a.J3, f, j, c) - accl_at(a.g3, 4, a.J3, f, j, c));
          sa += d * d;
        }
      }
    }
    const float* th = a.pt + (long long)f * 85;
    for (int i = threadIdx.x; i < 82; i += blockDim.x) sn += th[3 + i] * th[3 + i];
    if (frame_valid(a, f)) {
      const float* tg = a.gt + (long long)f * 85;
      if (threadIdx.x < 24) {
        float Rp[9], Rg[9];
        rodrigues<float>(th[3 + 3 * threadIdx.x], th[4 + 3 * threadIdx.x], th[5 + 3 * threadIdx.x], Rp);
        rodrigues<float>(tg[3 + 3 * threadIdx.x], tg[4 + 3 * threadIdx.x], tg[5 + 3 * threadIdx.x], Rg);
#pragma unroll
        for (int k = 0; k < 9; ++k) sp += (Rp[k] - Rg[k]) * (Rp[k] - Rg[k]);
      } else if (threadIdx.x >= 32 && threadIdx.x < 42) {
        const float d = th[75 + threadIdx.x - 32] - tg[75 + threadIdx.x - 32];
        ss = d * d;
      }
    }
  }
  s2 = block_sum(s2, buf); s3 = block_sum(s3, buf); sp = block_sum(sp, buf);
  ss = block_sum(ss, buf); sn = block_sum(sn, buf); sa = block_sum(sa, buf);
  if (threadIdx.x == 0) {
    float* o = a.partial + (long long)f * kPartial;
    o[0] = s2; o[1] = s3; o[2] = sp; o[3] = ss; o[4] = sn; o[5] = sa;
    o[6] = (f < a.M3 && frame_valid(a, f)) ? 1.f : 0.f; o[7] = 0.f;
  }
}

// scalars: [0..5] weighted terms (kp2d, kp3d, shape, pose, norm, accl — the reference's loss_dict order), [6] total,
// [7] n_valid, [8] sqrt(sum theta^2)
__global__ void loss_finalize_kernel(const LossArgs a) {
  __shared__ double sh[kPartial][kThreads];
  double acc[kPartial];
  for (int k = 0; k < kPartial; ++k) acc[k] = 0.0;
  // fixed assignment frame -> thread and a fixed tree below: bit-reproducible
  for (int f = threadIdx.x; f < a.frames; f += kThreads)
    for (int k = 0; k < kPartial; ++k) acc[k] += (double)a.partial[(long long)f * kPartial + k];
  for (int k = 0; k < kPartial; ++k) sh[k][threadIdx.x] = acc[k];
  __syncthreads();
  for (int o = kThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o)
      for (int k = 0; k < kPartial; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double nv = sh[6][0];
    const double l2 = (a.g2 && a.M2 > 0) ? a.w.kp2d * sh[0][0] / ((double)a.M2 * a.J2 * 2) : 0.0;
    const double l3 = (a.g3 && a.M3 > 0) ? a.w.kp3d * sh[1][0] / ((double)a.M3 * a.J3 * 3) : 0.0;
    const bool smpl = a.w.pose > 0.f && a.w.shape > 0.f && nv > 0.0;        // loss.py:194 / :84-93
    const double lp = smpl ? a.w.pose * sh[2][0] / (nv * 216.0) : 0.0;
    const double ls = smpl ? a.w.shape * sh[3][0] / (nv * 10.0) : 0.0;
    const double nrm = sqrt(sh[4][0]);
    const double ln = (a.w.norm > 0.f && a.M3 > 0) ? a.w.norm * nrm / (double)a.M3 : 0.0;
    double la = 0.0;
    if (a.w.accl > 0.f && a.g3 && a.T >= 3 && a.M3 > 0)
      la = a.w.accl * sh[5][0] / ((double)(a.M3 / a.T) * (a.T - 2) * a.J3 * 3);
    const float v[9] = {(float)l2, (float)l3, (float)ls, (float)lp, (float)ln, (float)la,
                        (float)(l2 + l3 + ls + lp + ln + la), (float)nv, (float)nrm};
    for (int k = 0; k < 9; ++k) a.scalars[k] = v[k];
    for (int k = 0; k < 8; ++k) a.losses[k] = v[k];
  }
}

__global__ void loss_grads_kernel(const LossArgs a) {
  const int f = blockIdx.x;
  const float nv = a.scalars[7], nrm = a.scalars[8];
  if (f < a.M2) {
    const float k2 = a.g2 ? 2.f * a.w.kp2d / ((float)a.M2 * a.J2 * 2) : 0.f;
    for (int i = threadIdx.x; i < a.J2 * 2; i += blockDim.x) {
      const int j = i >> 1, c = i & 1;
      const long long o = ((long long)f * a.J2 + j) * 2 + c;
      float g = 0.f;
      if (a.g2) {
        const float* gg = a.g2 + ((long long)f * a.J2 + j) * 3;
        g = k2 * gg[2] * (a.p2[o] - gg[c]);
      }
      a.d2[o] = g;
    }
  }
  if (f >= a.M3) return;
  // ---- kp_3d: centred residuals r_j = conf_j * d_j; grad_j = k3 * (r_j - [j in pelvis] * 0.5 * sum_i r_i)
  __shared__ float buf[32];
  __shared__ float rsum[3];
  const float* P = a.p3 + (long long)f * a.J3 * 3;
  if (a.g3) {
    const float* G = a.g3 + (long long)f * a.J3 * 4;
    const float k3 = 2.f * a.w.kp3d / ((float)a.M3 * a.J3 * 3);
    float part[3] = {0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < a.J3 * 3; i += blockDim.x) {
      const int j = i / 3, c = i % 3;
      const float pp = 0.5f * (P[a.pel_a * 3 + c] + P[a.pel_b * 3 + c]);
      const float pg = 0.5f * (G[a.pel_a * 4 + c] + G[a.pel_b * 4 + c]);
      const float r = G[j * 4 + 3] * ((P[j * 3 + c] - pp) - (G[j * 4 + c] - pg));
      part[c] += r;
    }
    for (int c = 0; c < 3; ++c) {
      const float t = block_sum(part[c], buf);
      if (threadIdx.x == 0) rsum[c] = t;
    }
    __syncthreads();
    const bool accl = a.w.accl > 0.f && a.T >= 3;
    const float ka = accl ? 2.f * a.w.accl / ((float)(a.M3 / a.T) * (a.T - 2) * a.J3 * 3) : 0.f;
    const int t = f % a.T;
    for (int i = threadIdx.x; i < a.J3 * 3; i += blockDim.x) {
      const int j = i / 3, c = i % 3;
      const float pp = 0.5f * (P[a.pel_a * 3 + c] + P[a.pel_b * 3 + c]);
      const float pg = 0.5f * (G[a.pel_a * 4 + c] + G[a.pel_b * 4 + c]);
      const float r = G[j * 4 + 3] * ((P[j * 3 + c] - pp) - (G[j * 4 + c] - pg));
      float g = k3 * (r - ((j == a.pel_a ? 0.5f : 0.f) + (j == a.pel_b ? 0.5f : 0.f)) * rsum[c]);
      if (accl) {
        // this frame is position s+2 (coef 1), s+1 (coef -2), s (coef 1) of the windows s = t-2, t-1, t
        const float coef[3] = {1.f, -2.f, 1.f};
        for (int q = 0; q < 3; ++q) {
          const int s = t - 2 + q;
          if (s < 0 || s + 2 >= a.T) continue;
          const int fs = f - t + s;
          const float cf = a.g3[(((long long)fs + 2) * a.J3 + j) * 4 + 3];
          const float c8 = cf * cf * cf * cf * cf * cf * cf * cf;
          g += ka * coef[q] * c8 * (accl_at(a.p3, 3, a.J3, fs, j, c) - accl_at(a.g3, 4, a.J3, fs, j, c));
        }
      }
      a.d3[(long long)f * a.J3 * 3 + i] = g;
    }
  } else {
    for (int i = threadIdx.x; i < a.J3 * 3; i += blockDim.x) a.d3[(long long)f * a.J3 * 3 + i] = 0.f;
  }
  // ---- theta: cam untouched (0); pose / shape terms on valid frames; norm term on every frame
  const float* th = a.pt + (long long)f * 85;
  const float* tg = a.gt + (long long)f * 85;
  float* dth = a.dt + (long long)f * 85;
  const bool smpl = a.w.pose > 0.f && a.w.shape > 0.f && nv > 0.f && frame_valid(a, f);
  const float kn = (a.w.norm > 0.f && nrm > 0.f) ? a.w.norm / (nrm * (float)a.M3) : 0.f;
  if (threadIdx.x < 3) dth[threadIdx.x] = 0.f;
  if (threadIdx.x < 24) {
    const int j = threadIdx.x;
    float g[3] = {0.f, 0.f, 0.f};
    if (smpl) {
      D3 Rp[9];
      float Rg[9];
      rodrigues<D3>(mk(th[3 + 3 * j], 0), mk(th[4 + 3 * j], 1), mk(th[5 + 3 * j], 2), Rp);
      rodrigues<float>(tg[3 + 3 * j], tg[4 + 3 * j], tg[5 + 3 * j], Rg);
      const float kp = 2.f * a.w.pose / (nv * 216.f);
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const float r = kp * (Rp[k].v - Rg[k]);
        g[0] += r * Rp[k].d[0]; g[1] += r * Rp[k].d[1]; g[2] += r * Rp[k].d[2];
      }
    }
    for (int c = 0; c < 3; ++c) dth[3 + 3 * j + c] = g[c] + kn * th[3 + 3 * j + c];
  } else if (threadIdx.x >= 32 && threadIdx.x < 42) {
    const int i = 75 + threadIdx.x - 32;
    float g = kn * th[i];
    if (smpl) g += 2.f * a.w.shape / (nv * 10.f) * (th[i] - tg[i]);
    dth[i] = g;
  }
}

}  // namespace

size_t loss_scratch_bytes(int M2, int M3) {
  const int frames = M2 > M3 ? M2 : M3;
  return (size_t)(frames > 0 ? frames : 1) * kPartial * 4 + 64;
}

int loss_forward_backward(const float* pred_kp2d, const float* gt_kp2d, int M2, int J2, const float* pred_kp3d,
                          const float* gt_kp3d, int M3, int J3, const float* pred_theta, const float* gt_theta,
                          const unsigned char* valid, int T, const LossWeights* w, float* losses, float* d_kp2d, float* d_kp3d,
                          float* d_theta, void* scratch, size_t scratch_bytes, cudaStream_t st) {
  MAED_CHECK_ARG(w && losses && scratch, "loss: null argument");
  MAED_CHECK_ARG(M2 >= 0 && M3 >= 0 && M2 + M3 > 0, "loss: empty batch (M2=%d, M3=%d)", M2, M3);
  MAED_CHECK_ARG(M2 == 0 || (pred_kp2d && d_kp2d && J2 >= 1), "loss: kp_2d predictions / gradient buffer missing");
  MAED_CHECK_ARG(M3 == 0 || (pred_kp3d && d_kp3d && pred_theta && gt_theta && d_theta), "loss: kp_3d / theta tensors missing");
  MAED_CHECK_ARG(M3 == 0 || (J3 > 28), "loss: kp_3d needs the 49-joint layout (pelvis = joints 27, 28); J3=%d", J3);
  MAED_CHECK_ARG(T >= 1 && (M3 % T) == 0, "loss: M3=%d is not a multiple of the clip length T=%d", M3, T);
  MAED_CHECK_ARG(scratch_bytes >= loss_scratch_bytes(M2, M3), "loss: scratch too small");
  LossArgs a;
  a.p2 = pred_kp2d; a.g2 = gt_kp2d; a.M2 = M2; a.J2 = J2;
  a.p3 = pred_kp3d; a.g3 = gt_kp3d; a.M3 = M3; a.J3 = J3; a.pel_a = 27; a.pel_b = 28;
  a.pt = pred_theta; a.gt = gt_theta; a.valid = valid; a.T = T; a.w = *w;
  a.frames = M2 > M3 ? M2 : M3;
  a.scalars = (float*)scratch;
  a.partial = (float*)scratch + 16;
  a.losses = losses; a.d2 = d_kp2d; a.d3 = d_kp3d; a.dt = d_theta;
  loss_partial_kernel<<<a.frames, kThreads, 0, st>>>(a);
  MAED_BW_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, kThreads, 0, st>>>(a);
  MAED_BW_LAUNCH_CHECK();
  loss_grads_kernel<<<a.frames, kThreads, 0, st>>>(a);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
