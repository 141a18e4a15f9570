// Host-side helpers shared by the libpetsyn translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/petsyn.h"

namespace petsyn {

// thread-local last-error message (petsyn_last_error)
char* last_error_buf();
int32_t fail(int32_t code, const char* fmt, ...);

#define PETSYN_CHECK_CUDA(expr)                                                                   \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::petsyn::fail(PETSYN_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                  \
  } while (0)

#define PETSYN_REQUIRE(cond, ...)                                  \
  do {                                                             \
    if (!(cond)) return ::petsyn::fail(PETSYN_EINVAL, __VA_ARGS__); \
  } while (0)

void count_launch();

inline int32_t check_launch(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(PETSYN_ECUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return PETSYN_OK;
}

// Encode a tiled TMA descriptor (cuTensorMapEncodeTiled resolved at run time through the CUDA runtime, so the
// library has no link-time dependency on libcuda and loads on a GPU-less box).
//   dims[rank]           extent of each dimension, innermost first
//   strides_bytes[rank-1] byte stride of dimensions 1..rank-1 (dimension 0 is contiguous)
//   box[rank]            box extent per dimension
//   swizzle_bytes        0, 32, 64 or 128
int32_t encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace petsyn
