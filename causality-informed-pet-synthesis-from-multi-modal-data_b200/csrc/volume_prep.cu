// Input-side data format of the path: what pair_PET_T1dataset._preprocess_img does to every raw
// volume before the generator sees it (unet/utils/dataset.py:70-105):
//     SpatialPad(crop_size) -> CenterSpatialCrop(crop_size) -> img / torch.max(img)
// Both MONAI transforms together are one gather through a per-axis window (out[o] = src[o + off] or 0), so the raw volume is
// read where it lies after the H2D copy and the network input is written once:
//     pass 1  window maximum           (reads the part of the raw volume inside the window: 4 B / voxel)
//     pass 2  gather, divide, store    (the same voxels, now L2-resident, + 4 B / voxel written)
// HBM-bound integer/address work, nothing for the tensor cores.  The division is IEEE fp32 (no fast-math), so the result is
// bit-identical to the reference's torch expression.
#include <cfloat>

#include "common.h"

namespace petsyn {
namespace {

constexpr int kMaxVolumes = 16;

struct Window {
  const float* src;
  int sd, sh, sw;     // raw extent
  int od, oh, ow;     // out voxel o reads src voxel o + off (may lie outside: zero padding)
  int padded;         // any output voxel outside the raw volume?
};

struct Batch {
  Window v[kMaxVolumes];
  int D, H, W;
};

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  // order-preserving integer views of IEEE floats: non-negative floats order as signed ints, negative ones reversed as
  // unsigned ints
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v + 0.f));      // -0 -> +0
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void init_max_kernel(float* vmax, int n) {
  if (threadIdx.x < n) vmax[threadIdx.x] = -INFINITY;
}

__global__ void __launch_bounds__(256) window_max_kernel(const Batch b, float* __restrict__ vmax) {
  const Window& w = b.v[blockIdx.y];
  // window ∩ raw volume, in raw coordinates
  const int z0 = max(w.od, 0), z1 = min(w.od + b.D, w.sd);
  const int y0 = max(w.oh, 0), y1 = min(w.oh + b.H, w.sh);
  const int x0 = max(w.ow, 0), x1 = min(w.ow + b.W, w.sw);
  const int nx = x1 - x0, ny = y1 - y0, nz = z1 - z0;
  float m = w.padded ? 0.f : -INFINITY;
  if (nx > 0 && ny > 0 && nz > 0) {
    const int64_t total = (int64_t)nx * ny * nz;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int x = (int)(i % nx);
      const int64_t r = i / nx;
      const int y = (int)(r % ny), z = (int)(r / ny);
      m = fmaxf(m, __ldg(w.src + ((int64_t)(z0 + z) * w.sh + (y0 + y)) * w.sw + (x0 + x)));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float warp_max[8];
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, warp_max[i]);
    atomic_max_float(vmax + blockIdx.y, m);
  }
}

__global__ void __launch_bounds__(256) crop_scale_kernel(const Batch b, const float* __restrict__ vmax,
                                                         float* __restrict__ dst) {
  const Window& w = b.v[blockIdx.y];
  const float m = __ldg(vmax + blockIdx.y);
  const int64_t total = (int64_t)b.D * b.H * b.W;
  float* out = dst + (int64_t)blockIdx.y * total;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % b.W);
    const int64_t r = i / b.W;
    const int y = (int)(r % b.H), z = (int)(r / b.H);
    const int sx = x + w.ow, sy = y + w.oh, sz = z + w.od;
    float v = 0.f;
    if (sx >= 0 && sx < w.sw && sy >= 0 && sy < w.sh && sz >= 0 && sz < w.sd)
      v = __ldg(w.src + ((int64_t)sz * w.sh + sy) * w.sw + sx);
    out[i] = __fdiv_rn(v, m);
  }
}

// SpatialPad (method "symmetric": before = width / 2) followed by CenterSpatialCrop (start = size / 2 - roi / 2, clipped at 0)
inline int window_offset(int raw, int roi) {
  const int width = roi > raw ? roi - raw : 0;
  const int before = width / 2;
  const int padded = raw + width;
  int start = padded / 2 - roi / 2;
  if (start < 0) start = 0;
  return start - before;
}

}  // namespace
}  // namespace petsyn

using namespace petsyn;

extern "C" int32_t petsyn_volume_window_offset(int32_t raw, int32_t roi) { return window_offset(raw, roi); }

extern "C" int32_t petsyn_volume_prepare(const petsyn_volume_src* srcs, int32_t n, float* dst, int32_t d, int32_t h,
                                         int32_t w, float* vmax, void* stream) {
  PETSYN_REQUIRE(srcs != nullptr && dst != nullptr && vmax != nullptr, "volume_prepare: null pointer");
  PETSYN_REQUIRE(n >= 1 && n <= kMaxVolumes, "volume_prepare: 1..%d volumes per call (got %d)", kMaxVolumes, n);
  PETSYN_REQUIRE(d > 0 && h > 0 && w > 0, "volume_prepare: bad crop size %dx%dx%d", d, h, w);
  Batch b;
  b.D = d; b.H = h; b.W = w;
  for (int i = 0; i < n; ++i) {
    const petsyn_volume_src& s = srcs[i];
    PETSYN_REQUIRE(s.data != nullptr && s.d > 0 && s.h > 0 && s.w > 0, "volume_prepare: volume %d is empty", i);
    Window& v = b.v[i];
    v.src = s.data;
    v.sd = s.d; v.sh = s.h; v.sw = s.w;
    v.od = window_offset(s.d, d);
    v.oh = window_offset(s.h, h);
    v.ow = window_offset(s.w, w);
    v.padded = (s.d < d || s.h < h || s.w < w) ? 1 : 0;
  }
  cudaStream_t st = as_stream(stream);
  init_max_kernel<<<1, 32, 0, st>>>(vmax, n);
  int32_t rc = check_launch("init_max_kernel");
  if (rc) return rc;
  const int64_t total = (int64_t)d * h * w;
  // a multiple of the SM count, split over the volumes of the batch
  int blocks = (int)((total + 256 * 8 - 1) / (256 * 8));
  const int cap = (148 * 8 + n - 1) / n;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  dim3 grid((unsigned)blocks, (unsigned)n);
  window_max_kernel<<<grid, 256, 0, st>>>(b, vmax);
  rc = check_launch("window_max_kernel");
  if (rc) return rc;
  crop_scale_kernel<<<grid, 256, 0, st>>>(b, vmax, dst);
  return check_launch("crop_scale_kernel");
}
