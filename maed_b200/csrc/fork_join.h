// Fork / join over a few side streams for bursts of small, mutually independent kernels (the per-step weight packs: ~170
// launches of 5-20 us each, every one far too small to fill 148 SMs; back to back on one stream they cost 1.4 ms per train
// step, mostly launch latency).  fork: the side streams wait for what `main` has queued so far; next(): round-robin side
// stream; join: `main` waits for all of them.  Streams and events are created once per (host thread, device) and kept.
// The CUDA-on-CPU test build has no streams: everything stays on `main`.
#pragma once
#include <cuda_runtime.h>

#include "common.h"

namespace maed {

class ForkJoin {
 public:
  static constexpr int kSide = 4;
  explicit ForkJoin(cudaStream_t main) : main_(main) {
#ifndef MAED_EMU
    Pool& p = pool();
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return;
    if (p.dev != dev) {                                   // first use on this thread / device: create (never destroyed)
      p.dev = -1;
      for (int i = 0; i < kSide; ++i) {
        if (cudaStreamCreateWithFlags(&p.side[i], cudaStreamNonBlocking) != cudaSuccess) return;
        if (cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming) != cudaSuccess) return;
      }
      if (cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming) != cudaSuccess) return;
      p.dev = dev;
    }
    if (cudaEventRecord(p.fork, main_) != cudaSuccess) return;
    for (int i = 0; i < kSide; ++i)
      if (cudaStreamWaitEvent(p.side[i], p.fork, 0) != cudaSuccess) return;
    pool_ = &p;
#endif
  }
  ~ForkJoin() { join(); }
  ForkJoin(const ForkJoin&) = delete;
  ForkJoin& operator=(const ForkJoin&) = delete;

  cudaStream_t next() {
#ifndef MAED_EMU
    if (pool_) return pool_->side[(n_++) % kSide];
#endif
    return main_;
  }
  int join() {
#ifndef MAED_EMU
    if (pool_) {
      Pool* p = pool_;
      pool_ = nullptr;
      for (int i = 0; i < kSide; ++i) {
        MAED_CUDA_CHECK(cudaEventRecord(p->done[i], p->side[i]));
        MAED_CUDA_CHECK(cudaStreamWaitEvent(main_, p->done[i], 0));
      }
    }
#endif
    return MAED_OK;
  }

 private:
#ifndef MAED_EMU
  struct Pool { int dev = -1; cudaStream_t side[kSide]; cudaEvent_t done[kSide]; cudaEvent_t fork; };
  static Pool& pool() { static thread_local Pool p; return p; }
  Pool* pool_ = nullptr;
  unsigned n_ = 0;
#endif
  cudaStream_t main_;
};

}  // namespace maed
