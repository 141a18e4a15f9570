// Layout kernels for the first and last layers of the pix2pix 3D U-Net (unet_model.py:54-65).  With Cin = 1 (stem) or
// Cout = 1 (head) a direct convolution is bandwidth-bound SIMT work; instead these kernels re-express both layers as
// 1x1x1 GEMMs (im2col patches for the stem; per-source-voxel projection + gather for the head) that run on the same
// tcgen05 implicit-GEMM kernels as every other convolution.  What remains here is pure data movement: coalesced
// 16-byte stores, L2-resident gathers.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"

namespace petsyn {

// one thread = one output voxel x 8 consecutive taps (16-byte store); tap = (kd*4 + kh)*4 + kw
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ patches,
                                                          int N, int D, int H, int W) {
  const int OD = D / 2, OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * OD * OH * OW * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int grp = (int)(i & 7);           // taps grp*8 .. grp*8+7  ->  kd = grp>>1, kh = (grp&1)*2 + {0,1}, kw = 0..3
    int64_t q = i >> 3;
    const int ow = (int)(q % OW); q /= OW;
    const int oh = (int)(q % OH); q /= OH;
    const int od = (int)(q % OD); q /= OD;
    const int n = (int)q;
    const int id = 2 * od + (grp >> 1) - 1;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ih = 2 * oh + (grp & 1) * 2 + (j >> 2) - 1, iw = 2 * ow + (j & 3) - 1;
      const bool ok = id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W;
      v[j] = ok ? __ldg(x + (((int64_t)n * D + id) * H + ih) * W + iw) : 0.f;
    }
    uint4 o;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
    reinterpret_cast<uint4*>(patches)[i] = o;
  }
}

// dx[i] = sum over (o, tap) with 2o + k - 1 == i of dpatches[o][tap]: two (o, k) pairs per axis -> 8 terms
__global__ void __launch_bounds__(256) stem_col2im_kernel(const __nv_bfloat16* __restrict__ dp, float* __restrict__ dx,
                                                          int N, int D, int H, int W) {
  const int OD = D / 2, OH = H / 2, OW = W / 2;
  const int64_t total = (int64_t)N * D * H * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i;
    const int iw = (int)(q % W); q /= W;
    const int ih = (int)(q % H); q /= H;
    const int id = (int)(q % D); q /= D;
    const int n = (int)q;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int kd = ((id + 1) & 1) + 2 * a, od = (id + 1 - kd) / 2;     // 2*od + kd - 1 == id
      if (od < 0 || od >= OD) continue;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const int kh = ((ih + 1) & 1) + 2 * b, oh = (ih + 1 - kh) / 2;
        if (oh < 0 || oh >= OH) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int kw = ((iw + 1) & 1) + 2 * c, ow = (iw + 1 - kw) / 2;
          if (ow < 0 || ow >= OW) continue;
          const int64_t o = (((int64_t)n * OD + od) * OH + oh) * OW + ow;
          acc += __bfloat162float(dp[o * 64 + (kd * 4 + kh) * 4 + kw]);
        }
      }
    }
    dx[i] = acc;
  }
}

// y[o] = tanh( sum_k proj[src(o,k)][k] ), src = (o + k - 1) >> 1 on each axis, zero outside the upsampled volume
__global__ void __launch_bounds__(256) head_gather_tanh_kernel(const float* __restrict__ proj, float* __restrict__ y,
                                                               int N, int D, int H, int W) {
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)N * OD * OH * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t q = i;
    const int ow = (int)(q % OW); q /= OW;
    const int oh = (int)(q % OH); q /= OH;
    const int od = (int)(q % OD); q /= OD;
    const int n = (int)q;
    float acc = 0.f;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const int ud = od + kd - 1;
      if (ud < 0 || ud >= OD) continue;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int uh = oh + kh - 1;
        if (uh < 0 || uh >= OH) continue;
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const int uw = ow + kw - 1;
          if (uw < 0 || uw >= OW) continue;
          const int64_t s = (((int64_t)n * D + (ud >> 1)) * H + (uh >> 1)) * W + (uw >> 1);
          acc += __ldg(proj + s * 32 + (kd * 3 + kh) * 3 + kw);
        }
      }
    }
    y[i] = tanhf(acc);
  }
}

// dproj[s][k] = sum over outputs o with src(o,k) == s of dy[o] * (1 - y[o]^2); bf16 rows of 64 (k >= 27 zero).
// One thread = one source voxel x 8 consecutive k (16-byte store).
__global__ void __launch_bounds__(256) head_scatter_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                               __nv_bfloat16* __restrict__ dproj, int N, int D, int H,
                                                               int W) {
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)N * D * H * W * 8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int grp = (int)(i & 7);
    int64_t q = i >> 3;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (grp < 4) {
      const int sw_ = (int)(q % W); q /= W;
      const int sh = (int)(q % H); q /= H;
      const int sd = (int)(q % D); q /= D;
      const int n = (int)q;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = grp * 8 + j;
        if (k >= 27) break;
        const int kd = k / 9, kh = (k / 3) % 3, kw = k % 3;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          const int od = 2 * sd + a + 1 - kd;
          if (od < 0 || od >= OD) continue;
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const int oh = 2 * sh + b + 1 - kh;
            if (oh < 0 || oh >= OH) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int ow = 2 * sw_ + c + 1 - kw;
              if (ow < 0 || ow >= OW) continue;
              const int64_t o = (((int64_t)n * OD + od) * OH + oh) * OW + ow;
              const float yv = __ldg(y + o);
              acc += __ldg(dy + o) * (1.f - yv * yv);
            }
          }
        }
        v[j] = acc;
      }
    }
    uint4 o4;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(v[4], v[5]), b3 = __floats2bfloat162_rn(v[6], v[7]);
    o4.x = *reinterpret_cast<uint32_t*>(&b0); o4.y = *reinterpret_cast<uint32_t*>(&b1);
    o4.z = *reinterpret_cast<uint32_t*>(&b2); o4.w = *reinterpret_cast<uint32_t*>(&b3);
    reinterpret_cast<uint4*>(dproj)[i] = o4;
  }
}

// ---- one-channel heads / tails of the BMGAN networks -------------------------------------------------------------
// generator input: cat([t1, z.view(N,8,1,1,1).expand(...)], 1) (bmgan_model.py:76-79) as NDHWC bf16 padded to cpad
__global__ void __launch_bounds__(256) concat_latent_kernel(const float* __restrict__ x, const float* __restrict__ zvec,
                                                            __nv_bfloat16* __restrict__ out, int64_t rows_per_sample,
                                                            int64_t rows, int nz, int cpad) {
  // one thread per (row, 8-channel chunk): 16-byte stores, consecutive threads write consecutive chunks
  const int cpr = cpad >> 3;
  const int64_t total = rows * cpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c0 = (int)(i - r * cpr) * 8;
    const int n = (int)(r / rows_per_sample);
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      f[j] = c == 0 ? __ldg(x + r) : (c <= nz ? __ldg(zvec + n * nz + c - 1) : 0.f);
    }
    uint4 o;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(out + r * cpad + c0) = o;
  }
}
// y[r] = src[r, 0]   (channel 0 of a padded fp32 conv output -> contiguous N,1,D,H,W)
__global__ void __launch_bounds__(256) take_channel0_kernel(const float* __restrict__ src, float* __restrict__ y,
                                                            int64_t rows, int cpad) {
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x)
    y[r] = src[r * cpad];
}
// dz[r, 0] = dy[r] * (tanh_out ? 1 - y[r]^2 : 1); dz[r, 1:] = 0   (bf16, padded)
__global__ void __launch_bounds__(256) put_channel0_grad_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                                __nv_bfloat16* __restrict__ dz, int64_t rows, int cpad,
                                                                int tanh_out) {
  const int cpr = cpad >> 3;
  const int64_t total = rows * cpr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpr;
    const int c0 = (int)(i - r * cpr) * 8;
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (c0 == 0) {
      float g = __ldg(dy + r);
      if (tanh_out) { const float v = __ldg(y + r); g *= 1.f - v * v; }
      __nv_bfloat162 b0 = __floats2bfloat162_rn(g, 0.f);
      o.x = *reinterpret_cast<uint32_t*>(&b0);
    }
    *reinterpret_cast<uint4*>(dz + r * cpad + c0) = o;
  }
}

}  // namespace petsyn

using namespace petsyn;

extern "C" {

int32_t petsyn_stem_im2col_k4s2(const float* x, void* patches, int32_t n, int32_t d, int32_t h, int32_t w,
                                void* stream) {
  PETSYN_REQUIRE(x && patches, "null argument");
  PETSYN_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && (d % 2 | h % 2 | w % 2) == 0, "stem: dims must be positive and even");
  const int64_t total = (int64_t)n * (d / 2) * (h / 2) * (w / 2) * 8;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  stem_im2col_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(patches), n, d, h, w);
  return check_launch("stem_im2col_kernel");
}

int32_t petsyn_stem_col2im_k4s2(const void* dpatches, float* dx, int32_t n, int32_t d, int32_t h, int32_t w,
                                void* stream) {
  PETSYN_REQUIRE(dpatches && dx, "null argument");
  PETSYN_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0 && (d % 2 | h % 2 | w % 2) == 0, "stem: dims must be positive and even");
  const int64_t total = (int64_t)n * d * h * w;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  stem_col2im_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dpatches), dx, n, d, h,
                                                            w);
  return check_launch("stem_col2im_kernel");
}

int32_t petsyn_concat_latent(const float* x, const float* zvec, void* out, int64_t rows_per_sample, int32_t n,
                             int32_t nz, int32_t cpad, void* stream) {
  PETSYN_REQUIRE(x && zvec && out, "null argument");
  PETSYN_REQUIRE(n > 0 && rows_per_sample > 0 && nz >= 0 && 1 + nz <= cpad && cpad % 8 == 0, "bad channel layout");
  const int64_t rows = rows_per_sample * n;
  const int blocks = (int)std::min<int64_t>((rows * (cpad / 8) + 255) / 256, 148 * 16);
  concat_latent_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, zvec, reinterpret_cast<__nv_bfloat16*>(out),
                                                              rows_per_sample, rows, nz, cpad);
  return check_launch("concat_latent_kernel");
}

int32_t petsyn_take_channel0(const float* src, float* y, int64_t rows, int32_t cpad, void* stream) {
  PETSYN_REQUIRE(src && y && rows > 0 && cpad > 0, "bad argument");
  const int blocks = (int)std::min<int64_t>((rows + 255) / 256, 148 * 16);
  take_channel0_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, y, rows, cpad);
  return check_launch("take_channel0_kernel");
}

int32_t petsyn_put_channel0_grad(const float* y, const float* dy, void* dz, int64_t rows, int32_t cpad,
                                 int32_t tanh_out, void* stream) {
  PETSYN_REQUIRE(dy && dz && rows > 0 && cpad > 0 && (!tanh_out || y), "bad argument");
  PETSYN_REQUIRE(cpad % 8 == 0, "padded channel count must be a multiple of 8");
  const int blocks = (int)std::min<int64_t>((rows * (cpad / 8) + 255) / 256, 148 * 16);
  put_channel0_grad_kernel<<<blocks, 256, 0, as_stream(stream)>>>(y, dy, reinterpret_cast<__nv_bfloat16*>(dz), rows, cpad,
                                                                  tanh_out);
  return check_launch("put_channel0_grad_kernel");
}

int32_t petsyn_head_gather_tanh(const float* proj, float* y, int32_t n, int32_t d, int32_t h, int32_t w, void* stream) {
  PETSYN_REQUIRE(proj && y, "null argument");
  PETSYN_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, "head: non-positive dims");
  const int64_t total = (int64_t)n * d * h * w * 8;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  head_gather_tanh_kernel<<<blocks, 256, 0, as_stream(stream)>>>(proj, y, n, d, h, w);
  return check_launch("head_gather_tanh_kernel");
}

int32_t petsyn_head_scatter_bwd(const float* y, const float* dy, void* dproj, int32_t n, int32_t d, int32_t h,
                                int32_t w, void* stream) {
  PETSYN_REQUIRE(y && dy && dproj, "null argument");
  PETSYN_REQUIRE(n > 0 && d > 0 && h > 0 && w > 0, "head: non-positive dims");
  const int64_t total = (int64_t)n * d * h * w * 8;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  head_scatter_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(y, dy, reinterpret_cast<__nv_bfloat16*>(dproj), n, d, h,
                                                                 w);
  return check_launch("head_scatter_bwd_kernel");
}

}  // extern "C"
