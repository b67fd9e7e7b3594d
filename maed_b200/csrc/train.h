// Training path of the MAED engine (train.cu): forward with a saved-activation tape, backward to every parameter.
#pragma once
#include "engine.h"

namespace maed {

struct TrainOutputs {
  float* feat;        // [BT, 768] or nullptr
  float* pose6d;      // [BT, 144]
  float* shape;       // [BT, 10]
  float* cam;         // [BT, 3]
};

size_t train_pack_bytes(const Engine* e);
size_t train_workspace_bytes(const Engine* e, int BT);
// derived weights of the data-gradient GEMMs (transposed / flipped planes); redo after every parameter update
int train_pack(const Engine* e, const void* const* params, void* tpack, cudaStream_t st);
// dropout_p = 0 reproduces the reference in eval() mode (the parity configuration); > 0: KTD dropout (ktd.py:54-56)
int train_forward(const Engine* e, const void* const* params, const void* packed, const float* x, int N, int T,
                  void* workspace, size_t workspace_bytes, float dropout_p, unsigned long long seed, const TrainOutputs* outs,
                  cudaStream_t st);
// grads[i]: fp32 device buffer of engine_param_numel(e, i) elements, overwritten with dL/dparam_i (un-scaled).
// d_pose6d / d_shape / d_cam: gradient of the loss w.r.t. the outputs of the matching train_forward call.
int train_backward(const Engine* e, const void* const* params, const void* packed, const void* tpack, const float* x, int N,
                   int T, void* workspace, size_t workspace_bytes, const float* d_pose6d, const float* d_shape,
                   const float* d_cam, float loss_scale, float dropout_p, float* const* grads, cudaStream_t st);

// SyncBatchNorm hook of the 'cnn' training path: fn(user, n) must add the first n doubles of `buf` (device memory, at least
// 2 * 2048 + 1 doubles) up over the data-parallel ranks in stream order; fn == nullptr removes the hook (per-rank statistics).
int train_set_exchange(Engine* e, int (*fn)(void*, int), void* user, double* buf, int capacity);

// Backward progress hook for overlapping the data-parallel gradient exchange with the rest of the backward: during
// train_backward, fn(user, first, end) is called on the host (in stream order: every kernel writing those gradients has been
// enqueued) when the gradients of the engine parameters with table index in [first, end) are final.  'ste': the tail range
// [blocks.k ..., table end) after block k (k = num_blocks - 1 ... 0), then [0, blocks.0) at the end.  fn == nullptr: no hook.
int train_set_progress(Engine* e, int (*fn)(void*, int, int), void* user);

}  // namespace maed
