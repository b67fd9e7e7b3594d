// Shared host-side helpers for libmaed_b200.so: error reporting, driver entry points, device info.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace maed {

// Status codes returned by every C-ABI entry point (include/maed_b200.h).
enum : int { MAED_OK = 0, MAED_ERR_CUDA = 1, MAED_ERR_ARG = 2, MAED_ERR_UNSUPPORTED = 3, MAED_ERR_DRIVER = 4 };

void set_error(const char* fmt, ...);
const char* last_error();

#define MAED_CUDA_CHECK(expr)                                                                       \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::maed::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return ::maed::MAED_ERR_CUDA;                                                                 \
    }                                                                                               \
  } while (0)

#define MAED_CHECK_ARG(cond, ...)           \
  do {                                      \
    if (!(cond)) {                          \
      ::maed::set_error(__VA_ARGS__);       \
      return ::maed::MAED_ERR_ARG;          \
    }                                       \
  } while (0)

#define MAED_PROPAGATE(expr)               \
  do {                                     \
    int _s = (expr);                       \
    if (_s != ::maed::MAED_OK) return _s;  \
  } while (0)

// cuTensorMapEncodeTiled resolved at run time through cudaGetDriverEntryPoint, so the library has no link
// dependency on libcuda.so and loads (for symbol checks) on a machine without a GPU driver.
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
int get_encode_tiled(PFN_encodeTiled* fn);

// fp16 tensor map with 128-byte swizzle; dims/box innermost first; strides in bytes for dims 1..rank-1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, int swizzle_bytes = 128);

int sm_count();
inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace maed
