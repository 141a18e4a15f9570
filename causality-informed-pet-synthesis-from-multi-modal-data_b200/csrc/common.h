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

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch.  A step is ~1 000 kernels of 5-50 us that depend on each other in stream order; with a plain
// launch every kernel pays the grid drain of its predecessor plus its own launch latency (~2-3 us).  Kernels launched through
// launch_pdl() may become resident while their predecessor is still running; they do their on-chip prologue (barrier init, TMEM
// allocation, descriptor prefetch) and then block in pdl_sync() until the predecessor grid has completed and its memory is
// visible.  Rules: (1) a kernel launched through launch_pdl() MUST call pdl_sync() before its first global-memory access;
// (2) pdl_sync() waits first and releases its own dependents second, so a kernel never runs ahead of the grid TWO launches
// before it.  Measured on configs[1] (B200, whole step replayed as a CUDA graph): 15.74 ms with the attribute against 15.61 ms
// without -- the ~2 us per launch that separate the step from the sum of its kernels are not launch latency that an early
// start can hide (the kernels' own tails and the concurrent weight-gradient stream are) -- so the attribute is OPT-IN
// (PETSYN_PDL=1); pdl_sync() is a no-op in a kernel launched without it.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#ifdef __CUDACC__
// wait for the grids this one depends on (all their memory operations are visible afterwards), then let the next grid in the
// stream become resident; a no-op for a kernel launched without the attribute
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
#endif

}  // namespace petsyn
