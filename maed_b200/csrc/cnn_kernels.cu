// Memory-bound kernels of the 'cnn' encoder (torchvision ResNet-50, reference lib/models/maed.py:35-37): BatchNorm folding
// at pack time and the 3x3/2 max-pool behind the stem.  HBM-bound: 16-byte accesses along the channel dimension, grids capped
// at a few waves of the SMs.
#include "device_utils.cuh"

namespace maed {

using bw::grid_for;
using bw::store_split4;

// ------------------------------------------------------------------------------------ BatchNorm folding (eval mode)
// y = (conv(x, w) - mean) / sqrt(var + eps) * gamma + beta  ==  conv(x, w * s) + (beta - mean * s),  s = gamma / sqrt(var + eps)
// One block per output channel; E = Cin*KH*KW weights per channel.
__global__ void fold_bn_kernel(const float* __restrict__ w, long long E, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, float* __restrict__ w_out, float* __restrict__ bias_out) {
  const int co = blockIdx.x;
  const float s = gamma[co] / sqrtf(var[co] + eps);
  const float* wc = w + (long long)co * E;
  float* oc = w_out + (long long)co * E;
  for (long long i = threadIdx.x; i < E; i += blockDim.x) oc[i] = wc[i] * s;
  if (threadIdx.x == 0) bias_out[co] = beta[co] - mean[co] * s;
}
int fold_bn(const float* w, int Cout, long long E, const float* gamma, const float* beta, const float* mean, const float* var,
            float eps, float* w_out, float* bias_out, cudaStream_t st) {
  MAED_CHECK_ARG(w && gamma && beta && mean && var && w_out && bias_out, "fold_bn: null argument");
  MAED_CHECK_ARG(Cout >= 1 && E >= 1, "fold_bn: bad shape Cout=%d E=%lld", Cout, E);
  fold_bn_kernel<<<Cout, 256, 0, st>>>(w, E, gamma, beta, mean, var, eps, w_out, bias_out);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

// ----------------------------------------------------------------------------- nn.MaxPool2d(3, stride 2, padding 1)
// fp32 NHWC in -> fp32 NHWC out (the identity of the first bottleneck) and planes (its tensor-core operand).  Padding
// positions never win (PyTorch pads with -inf): taps outside the image are skipped.
__global__ void maxpool3x3s2_kernel(const float* __restrict__ x, int H, int W, int C, int OH, int OW, long long total4,
                                    float* __restrict__ out_f32, __half* __restrict__ out_hi, long long plane) {
  const int c4n = C >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4n) * 4;
    const long long m = i / c4n;
    const int ow = (int)(m % OW);
    const int oh = (int)((m / OW) % OH);
    const long long n = m / ((long long)OW * OH);
    float4 best = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ih = oh * 2 + r - 1;
      if (ih < 0 || ih >= H) continue;
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iw = ow * 2 + s - 1;
        if (iw < 0 || iw >= W) continue;
        const float4 v = *reinterpret_cast<const float4*>(x + ((n * H + ih) * W + iw) * C + c);
        best.x = fmaxf(best.x, v.x); best.y = fmaxf(best.y, v.y); best.z = fmaxf(best.z, v.z); best.w = fmaxf(best.w, v.w);
      }
    }
    const long long o = m * C + c;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = best;
    if (out_hi) store_split4(out_hi + o, plane, best);
  }
}
int maxpool3x3s2(const float* x, int n_img, int H, int W, int C, float* out_f32, __half* out_hi, long long plane,
                 cudaStream_t st) {
  MAED_CHECK_ARG(x && (out_f32 || out_hi), "maxpool3x3s2: null argument");
  MAED_CHECK_ARG(C % 4 == 0 && H >= 1 && W >= 1 && n_img >= 1, "maxpool3x3s2: bad shape n=%d H=%d W=%d C=%d", n_img, H, W, C);
  MAED_CHECK_ARG(!out_hi || plane % 4 == 0, "maxpool3x3s2: plane stride must be a multiple of 4");
  const int OH = (H + 2 - 3) / 2 + 1, OW = (W + 2 - 3) / 2 + 1;
  const long long total4 = (long long)n_img * OH * OW * (C / 4);
  maxpool3x3s2_kernel<<<grid_for(total4, 256), 256, 0, st>>>(x, H, W, C, OH, OW, total4, out_f32, out_hi, plane);
  MAED_BW_LAUNCH_CHECK();
  return MAED_OK;
}

}  // namespace maed
