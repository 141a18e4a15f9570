#include "common.h"
#include "det_reduce.cuh"

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace petsyn {

uint64_t launch_count();
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("PETSYN_PDL");
    on = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 1 : 0;
  }
  return on == 1;
}

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int32_t fail(int32_t code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn resolve_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int32_t encode_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int rank, const void* base, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = resolve_encode();
  if (!fn) return fail(PETSYN_ECUDA, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swizzle_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swizzle_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swizzle_bytes == 128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return fail(PETSYN_ECUDA,
                "cuTensorMapEncodeTiled failed (%d): rank %d base %p dims [%llu %llu %llu %llu %llu] strides [%llu "
                "%llu %llu %llu] box [%u %u %u %u %u] swizzle %d",
                (int)r, rank, base, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                (unsigned long long)(rank > 4 ? gdim[4] : 0), (unsigned long long)(rank > 1 ? gstr[0] : 0),
                (unsigned long long)(rank > 2 ? gstr[1] : 0), (unsigned long long)(rank > 3 ? gstr[2] : 0),
                (unsigned long long)(rank > 4 ? gstr[3] : 0), bdim[0], rank > 1 ? bdim[1] : 0, rank > 2 ? bdim[2] : 0,
                rank > 3 ? bdim[3] : 0, rank > 4 ? bdim[4] : 0, swizzle_bytes);
  }
  return PETSYN_OK;
}

// per-device workspace of the reproducible reductions (det_reduce.cuh): allocated at the first launch that needs it
int32_t det_workspace(DetWs* out) {
  static std::mutex mu;
  static DetWs ws[64];
  static int mode = -1;           // 1: deterministic (default), 0: float atomics (PETSYN_NONDETERMINISTIC=1)
  std::lock_guard<std::mutex> lock(mu);
  if (mode < 0) {
    const char* e = getenv("PETSYN_NONDETERMINISTIC");
    mode = (e != nullptr && e[0] != '\0' && e[0] != '0') ? 0 : 1;
  }
  if (mode == 0) { *out = DetWs(); return PETSYN_OK; }
  int dev = 0;
  PETSYN_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) return fail(PETSYN_EINVAL, "device index %d out of range", dev);
  if (ws[dev].acc == nullptr) {
    void* p = nullptr;
    const size_t acc_bytes = (size_t)kDetRows * kDetRowValues * sizeof(double);
    cudaError_t err = cudaMalloc(&p, acc_bytes + kDetRows * sizeof(unsigned int));
    if (err != cudaSuccess)
      return fail(PETSYN_ECUDA, "allocating the reduction workspace failed: %s (the first norm / loss launch of a device "
                                "must not happen inside a CUDA-graph capture)", cudaGetErrorString(err));
    PETSYN_CHECK_CUDA(cudaMemset(p, 0, acc_bytes + kDetRows * sizeof(unsigned int)));
    ws[dev].acc = reinterpret_cast<double*>(p);
    ws[dev].counters = reinterpret_cast<unsigned int*>(ws[dev].acc + (size_t)kDetRows * kDetRowValues);
  }
  *out = ws[dev];
  return PETSYN_OK;
}

}  // namespace petsyn

extern "C" {
int32_t petsyn_version(void) { return PETSYN_VERSION; }
const char* petsyn_last_error(void) { return petsyn::last_error_buf(); }
uint64_t petsyn_launch_count(void) { return petsyn::launch_count(); }
}
