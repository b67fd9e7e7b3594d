// CUDA-on-CPU execution shim — TEST INFRASTRUCTURE ONLY (tests/emu/README.md).
//
// Lets g++ compile the *unmodified* CUDA-core translation units of maed_b200/csrc (kernels.cu, decoder.cu,
// bwd_kernels*.cu, attention_bwd.cu, smpl.cu) and the host orchestration (engine.cu, train.cu, capi.cu) into
// tests/emu/_build/libmaed_emu.so, so that the index logic of every kernel and the buffer plumbing of the training
// path can be checked against torch autograd / the reference's gradient digests on a machine without a GPU.
// The product library (maed_b200/libmaed_b200.so) never sees this header; nothing under maed_b200/ loads the emulator.
//
// Execution model: one kernel launch = loop over blocks (spread over a few OS threads); the threads of a block are
// fibers on one OS thread, scheduled round-robin.  __syncthreads() / warp shuffles / __syncwarp() yield until every live
// thread of the block / warp has arrived, so barrier-divergence deadlocks are detected (abort with a message) instead of
// hanging.  `__shared__` variables become `static thread_local` (one copy per OS thread = per running block).
// The force-include order matters: this file must come first (-include) so that __shared__ / __constant__ keep the
// definitions below when the CUDA headers are read in host mode.
#pragma once
#define MAED_EMU 1
#define __shared__ static thread_local
#define __constant__ static const

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <functional>
#include <tuple>
#include <type_traits>
#include <utility>

#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__

namespace emu {

struct ThreadState;
// per-OS-thread view of the running CUDA thread
extern thread_local uint3 t_threadIdx;
extern thread_local uint3 t_blockIdx;
extern thread_local dim3 t_blockDim;
extern thread_local dim3 t_gridDim;
extern thread_local uint8_t* t_dyn_smem;

void block_sync();                       // __syncthreads
void warp_sync();                        // __syncwarp / both halves of a shuffle
uint64_t* warp_slots();                  // 2 x 32 exchange slots of the running thread's warp
unsigned shfl_parity();                  // per-thread shuffle counter (post-incremented)
int lane_id();
int warp_lanes();                        // live width of the running thread's warp (blockDim tail)
void run_grid(dim3 grid, dim3 block, size_t smem, const std::function<void()>& thread_fn);
void prof_kernel(const void* fn, double seconds);      // per-kernel wall time (MAED_EMU_PROFILE=1, dumped at exit)
double prof_now();

// dynamic shared memory above 48 KB needs cudaFuncSetAttribute(MaxDynamicSharedMemorySize) on the real runtime: recorded per
// kernel by emu_cudaFuncSetAttribute and enforced at every launch
void set_smem_optin(const void* kernel, int bytes);
void check_smem_optin(const void* kernel, size_t bytes);

struct Launch {
  dim3 g, b;
  size_t smem;
  Launch(dim3 grid, dim3 block, size_t smem_bytes = 0, cudaStream_t = nullptr) : g(grid), b(block), smem(smem_bytes) {}
  template <class... P, class... A>
  void call(void (*k)(P...), A&&... a) {
    std::tuple<std::decay_t<P>...> args(std::forward<A>(a)...);      // kernel parameters are passed by value
    const double t0 = prof_now();
    check_smem_optin(reinterpret_cast<const void*>(k), smem);
    run_grid(g, b, smem, [&]() { std::apply(k, args); });
    prof_kernel(reinterpret_cast<const void*>(k), prof_now() - t0);
  }
};

template <class T>
inline T shfl_from(T v, int src) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  // two slot sets used alternately: a lane can run at most one shuffle ahead of the slowest lane (the next warp_sync stops
  // it), so the set it overwrites then has been read by everybody — one barrier per shuffle instead of two
  uint64_t* s = warp_slots() + 32 * (shfl_parity() & 1);
  const int lane = lane_id();
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  s[lane] = raw;
  warp_sync();
  T r = v;
  if (src >= 0 && src < warp_lanes()) memcpy(&r, &s[src], sizeof(T));
  return r;
}

}  // namespace emu

#define threadIdx (::emu::t_threadIdx)
#define blockIdx (::emu::t_blockIdx)
#define blockDim (::emu::t_blockDim)
#define gridDim (::emu::t_gridDim)

inline void __syncthreads() { ::emu::block_sync(); }
inline void __syncwarp(unsigned = 0xffffffffu) { ::emu::warp_sync(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return ::emu::shfl_from(v, ::emu::lane_id() ^ m); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return ::emu::shfl_from(v, ::emu::lane_id() + (int)d); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return ::emu::shfl_from(v, ::emu::lane_id() - (int)d); }
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return ::emu::shfl_from(v, src & 31); }

template <class T> inline T __ldg(const T* p) { return *p; }

// atomics: blocks run concurrently on several OS threads, so these are real atomics
inline float atomicAdd(float* p, float v) {
  uint32_t* ip = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu;
  float f;
  do {
    memcpy(&f, &old, 4);
    const float s = f + v;
    memcpy(&neu, &s, 4);
  } while (!__atomic_compare_exchange_n(ip, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
inline double atomicAdd(double* p, double v) {
  uint64_t* ip = reinterpret_cast<uint64_t*>(p);
  uint64_t old = __atomic_load_n(ip, __ATOMIC_RELAXED), neu;
  double f;
  do {
    memcpy(&f, &old, 8);
    const double s = f + v;
    memcpy(&neu, &s, 8);
  } while (!__atomic_compare_exchange_n(ip, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

// CUDA's global min / max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline float min(float a, float b) { return fminf(a, b); }
inline float max(float a, float b) { return fmaxf(a, b); }

// device math that glibc lacks
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
inline float __expf(float x) { return expf(x); }
inline float __fdividef(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __saturatef(float a) { return a < 0.f ? 0.f : (a > 1.f ? 1.f : a); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; memcpy(&v, &f, 4); return v; }
inline long long clock64() { return 0; }

// ---- the handful of runtime entry points the translation units use (renamed so that nothing collides with the real
// libcudart that torch has loaded into the test process)
inline cudaError_t emu_cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t emu_cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
inline cudaError_t emu_cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
inline cudaError_t emu_cudaGetLastError() { return cudaSuccess; }
inline const char* emu_cudaGetErrorString(cudaError_t) { return "emulated"; }
template <class F> inline cudaError_t emu_cudaFuncSetAttribute(F f, cudaFuncAttribute a, int v) {
  if (a == cudaFuncAttributeMaxDynamicSharedMemorySize) ::emu::set_smem_optin(reinterpret_cast<const void*>(f), v);
  return cudaSuccess;
}
#define cudaMemsetAsync emu_cudaMemsetAsync
#define cudaMemcpyAsync emu_cudaMemcpyAsync
#define cudaMemcpy2DAsync emu_cudaMemcpy2DAsync
#define cudaGetLastError emu_cudaGetLastError
#define cudaGetErrorString emu_cudaGetErrorString
#define cudaFuncSetAttribute emu_cudaFuncSetAttribute
