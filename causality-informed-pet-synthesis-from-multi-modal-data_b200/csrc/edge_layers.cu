// First and last layers of the pix2pix 3D U-Net (unet_model.py:54-65).  With Cin = 1 (stem) or Cout = 1 (head) these are
// HBM/L2-bandwidth-bound by construction (arithmetic intensity ~60 and ~200 FLOP/B), so they are plain SIMT kernels
// with coalesced 16-byte accesses, shared-memory weight staging and register accumulation -- not tensor-core GEMMs.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"

namespace petsyn {

// ------------------------------------------------------------------------------------------------ stem: Conv3d(1->C, k4 s2 p1)
constexpr int kStemVox = 32;   // output voxels per tile

__device__ __forceinline__ void stem_load_patches(const float* __restrict__ x, float (*patch)[65], int64_t tile0,
                                                  int64_t nvox, int D, int H, int W) {
  const int OD = D / 2, OH = H / 2, OW = W / 2;
  for (int e = threadIdx.x; e < kStemVox * 64; e += blockDim.x) {
    const int v = e >> 6, tap = e & 63;
    const int64_t o = tile0 + v;
    float val = 0.f;
    if (o < nvox) {
      int64_t q = o;
      const int ow = (int)(q % OW); q /= OW;
      const int oh = (int)(q % OH); q /= OH;
      const int od = (int)(q % OD); q /= OD;
      const int n = (int)q;
      const int id = 2 * od + (tap >> 4) - 1, ih = 2 * oh + ((tap >> 2) & 3) - 1, iw = 2 * ow + (tap & 3) - 1;
      if (id >= 0 && id < D && ih >= 0 && ih < H && iw >= 0 && iw < W)
        val = __ldg(x + (((int64_t)n * D + id) * H + ih) * W + iw);
    }
    patch[v][tap] = val;
  }
}

// y[vox][c] = sum_tap patch[vox][tap] * w[c][tap]
__global__ void __launch_bounds__(128) stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                       __nv_bfloat16* __restrict__ y, int N, int D, int H, int W,
                                                       int C) {
  extern __shared__ float smem_f[];
  float* sw = smem_f;                                          // [64][C] (tap-major)
  float(*patch)[65] = reinterpret_cast<float(*)[65]>(smem_f + 64 * C);   // [32][65]
  for (int e = threadIdx.x; e < 64 * C; e += blockDim.x) {
    const int c = e / 64, tap = e % 64;
    sw[tap * C + c] = w[e];
  }
  const int64_t nvox = (int64_t)N * (D / 2) * (H / 2) * (W / 2);
  const int groups = C / 8;                  // channel groups of 8
  const int vpp = blockDim.x / groups;       // voxels per pass
  const int g = threadIdx.x % groups, vl = threadIdx.x / groups;
  for (int64_t tile0 = (int64_t)blockIdx.x * kStemVox; tile0 < nvox; tile0 += (int64_t)gridDim.x * kStemVox) {
    __syncthreads();
    stem_load_patches(x, patch, tile0, nvox, D, H, W);
    __syncthreads();
    if (vl < vpp) {
      for (int v = vl; v < kStemVox; v += vpp) {
        if (tile0 + v >= nvox) break;
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 8
        for (int tap = 0; tap < 64; ++tap) {
          const float pv = patch[v][tap];
          const float4 w0 = *reinterpret_cast<const float4*>(sw + tap * C + g * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(sw + tap * C + g * 8 + 4);
          acc[0] += pv * w0.x; acc[1] += pv * w0.y; acc[2] += pv * w0.z; acc[3] += pv * w0.w;
          acc[4] += pv * w1.x; acc[5] += pv * w1.y; acc[6] += pv * w1.z; acc[7] += pv * w1.w;
        }
        uint4 o;
        __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[0], acc[1]), b1 = __floats2bfloat162_rn(acc[2], acc[3]);
        __nv_bfloat162 b2 = __floats2bfloat162_rn(acc[4], acc[5]), b3 = __floats2bfloat162_rn(acc[6], acc[7]);
        o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
        o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
        *reinterpret_cast<uint4*>(y + (tile0 + v) * C + g * 8) = o;
      }
    }
  }
}

// dw[c][tap] += sum_vox dy[vox][c] * patch[vox][tap]; thread owns an 8(c) x 8(tap) block of dw for its c/tap groups
__global__ void __launch_bounds__(128) stem_wgrad_kernel(const float* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ dy, float* __restrict__ dw,
                                                         int N, int D, int H, int W, int C) {
  extern __shared__ float smem_f[];
  float(*patch)[65] = reinterpret_cast<float(*)[65]>(smem_f);          // [32][65]
  float* sdy = smem_f + kStemVox * 65;                                  // [32][C]
  const int64_t nvox = (int64_t)N * (D / 2) * (H / 2) * (W / 2);
  const int cgroups = C / 8;                      // 8-channel groups
  const int nblk = cgroups * 8;                   // (c-group, tap-group) blocks of 8x8
  const int per_thread = (nblk + blockDim.x - 1) / blockDim.x;   // 1 for C = 128
  for (int rep = 0; rep < per_thread; ++rep) {
    const int blk = threadIdx.x + rep * blockDim.x;
    const bool active = blk < nblk;
    const int cg = active ? blk % cgroups : 0, tg = active ? blk / cgroups : 0;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int64_t tile0 = (int64_t)blockIdx.x * kStemVox; tile0 < nvox; tile0 += (int64_t)gridDim.x * kStemVox) {
      __syncthreads();
      stem_load_patches(x, patch, tile0, nvox, D, H, W);
      for (int e = threadIdx.x; e < kStemVox * C; e += blockDim.x) {
        const int v = e / C;
        sdy[e] = (tile0 + v < nvox) ? __bfloat162float(dy[(tile0 + v) * C + e % C]) : 0.f;
      }
      __syncthreads();
      if (active) {
        for (int v = 0; v < kStemVox; ++v) {
          float dv[8], pv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) dv[i] = sdy[v * C + cg * 8 + i];
#pragma unroll
          for (int j = 0; j < 8; ++j) pv[j] = patch[v][tg * 8 + j];
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] += dv[i] * pv[j];
        }
      }
    }
    if (active) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dw + (cg * 8 + i) * 64 + tg * 8 + j, acc[i][j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ head: Up x2 + Conv3d(C->1,k3,p1) + tanh
// proj[s][k] = sum_c x[s][c] * w[k][c]  (k < 27, row pitch 32): the 27-tap conv on the upsampled grid becomes one
// small projection per SOURCE voxel followed by a 27-term gather per output voxel.
__global__ void __launch_bounds__(256) head_proj_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                        float* __restrict__ proj, int64_t rows, int C) {
  extern __shared__ float sw[];   // [27][C]; w is [1][C][27] in memory
  for (int e = threadIdx.x; e < 27 * C; e += blockDim.x) {
    const int c = e / 27, k = e % 27;
    sw[k * C + c] = w[e];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int64_t s = (int64_t)blockIdx.x * nwarps + warp; s < rows; s += (int64_t)gridDim.x * nwarps) {
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.f;
    for (int c0 = lane * 8; c0 < C; c0 += 256) {
      const uint4 raw = *reinterpret_cast<const uint4*>(x + s * C + c0);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
      float xv[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(b2[i]);
        xv[2 * i] = f.x; xv[2 * i + 1] = f.y;
      }
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const float4 w0 = *reinterpret_cast<const float4*>(sw + k * C + c0);
        const float4 w1 = *reinterpret_cast<const float4*>(sw + k * C + c0 + 4);
        acc[k] += xv[0] * w0.x + xv[1] * w0.y + xv[2] * w0.z + xv[3] * w0.w + xv[4] * w1.x + xv[5] * w1.y +
                  xv[6] * w1.z + xv[7] * w1.w;
      }
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      float v = acc[k];
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (lane == k) proj[s * 32 + k] = v;
    }
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

// dproj[s][k] = sum over outputs o with src(o,k) == s of dy[o] * (1 - y[o]^2)
__global__ void __launch_bounds__(256) head_scatter_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy,
                                                               float* __restrict__ dproj, int N, int D, int H, int W) {
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  const int64_t total = (int64_t)N * D * H * W * 32;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i & 31);
    int64_t q = i >> 5;
    float acc = 0.f;
    if (k < 27) {
      const int sw_ = (int)(q % W); q /= W;
      const int sh = (int)(q % H); q /= H;
      const int sd = (int)(q % D); q /= D;
      const int n = (int)q;
      const int kd = k / 9, kh = (k / 3) % 3, kw = k % 3;
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
    }
    dproj[i] = acc;
  }
}

// dx[s][c] = sum_k dproj[s][k] w[k][c];  dw[c][k] += sum_s dproj[s][k] x[s][c].  Thread <-> channel.
__global__ void __launch_bounds__(256) head_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ w,
                                                       const float* __restrict__ dproj, __nv_bfloat16* __restrict__ dx,
                                                       float* __restrict__ dw, int64_t rows, int C,
                                                       int64_t rows_per_block) {
  __shared__ float sdp[32][32];   // 32 source voxels x 32 taps
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  for (int cbase = 0; cbase < C; cbase += blockDim.x) {   // uniform trip count: the loop body has barriers
    const int c = cbase + threadIdx.x;
    const bool active = c < C;
    float wk[27], acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) { wk[k] = active ? w[c * 27 + k] : 0.f; acc[k] = 0.f; }
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
      __syncthreads();
      for (int e = threadIdx.x; e < 32 * 32; e += blockDim.x) {
        const int64_t r = r0 + (e >> 5);
        sdp[e >> 5][e & 31] = (r < r_end) ? dproj[r * 32 + (e & 31)] : 0.f;
      }
      __syncthreads();
      const int nr = active ? (int)min((long long)32, (long long)(r_end - r0)) : 0;
      for (int v = 0; v < nr; ++v) {
        const float xv = __bfloat162float(x[(r0 + v) * C + c]);
        float d = 0.f;
#pragma unroll
        for (int k = 0; k < 27; ++k) {
          const float dp = sdp[v][k];
          d += dp * wk[k];
          acc[k] += dp * xv;
        }
        dx[(r0 + v) * C + c] = __float2bfloat16(d);
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < 27; ++k) atomicAdd(dw + c * 27 + k, acc[k]);
    }
  }
}

}  // namespace petsyn

using namespace petsyn;

extern "C" {

int32_t petsyn_stem_conv_k4s2_fwd(const float* x, const float* w, void* y, int32_t n, int32_t d, int32_t h, int32_t w_,
                                  int32_t cout, void* stream) {
  PETSYN_REQUIRE(x && w && y, "null argument");
  PETSYN_REQUIRE(cout % 8 == 0 && cout >= 8 && cout <= 512, "stem: cout must be a multiple of 8 in [8, 512]");
  PETSYN_REQUIRE((d % 2 | h % 2 | w_ % 2) == 0, "stem: dims must be even");
  const size_t smem = (size_t)(64 * cout + kStemVox * 65) * sizeof(float);
  PETSYN_CHECK_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nvox = (int64_t)n * (d / 2) * (h / 2) * (w_ / 2);
  const int blocks = (int)std::min<int64_t>((nvox + kStemVox - 1) / kStemVox, 148 * 4);
  stem_fwd_kernel<<<blocks, 128, smem, as_stream(stream)>>>(x, w, reinterpret_cast<__nv_bfloat16*>(y), n, d, h, w_, cout);
  return check_launch("stem_fwd_kernel");
}

int32_t petsyn_stem_conv_k4s2_wgrad(const float* x, const void* dy, float* dw, int32_t n, int32_t d, int32_t h,
                                    int32_t w_, int32_t cout, void* stream) {
  PETSYN_REQUIRE(x && dy && dw, "null argument");
  PETSYN_REQUIRE(cout % 8 == 0, "stem: cout must be a multiple of 8");
  cudaStream_t st = as_stream(stream);
  PETSYN_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)cout * 64 * sizeof(float), st));
  const size_t smem = (size_t)(kStemVox * 65 + kStemVox * cout) * sizeof(float);
  PETSYN_CHECK_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t nvox = (int64_t)n * (d / 2) * (h / 2) * (w_ / 2);
  const int blocks = (int)std::min<int64_t>((nvox + kStemVox - 1) / kStemVox, 148 * 4);
  stem_wgrad_kernel<<<blocks, 128, smem, st>>>(x, reinterpret_cast<const __nv_bfloat16*>(dy), dw, n, d, h, w_, cout);
  return check_launch("stem_wgrad_kernel");
}

int32_t petsyn_head_upconv_tanh_fwd(const void* x, const float* w, float* proj, float* y, int32_t n, int32_t d,
                                    int32_t h, int32_t w_, int32_t cin, void* stream) {
  PETSYN_REQUIRE(x && w && proj && y, "null argument");
  PETSYN_REQUIRE(cin % 8 == 0, "head: cin must be a multiple of 8");
  cudaStream_t st = as_stream(stream);
  const int64_t rows = (int64_t)n * d * h * w_;
  const size_t smem = (size_t)27 * cin * sizeof(float);
  PETSYN_CHECK_CUDA(cudaFuncSetAttribute(head_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int blocks = (int)std::min<int64_t>((rows + 7) / 8, 148 * 8);
  head_proj_kernel<<<blocks, 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, proj, rows, cin);
  int32_t rc = check_launch("head_proj_kernel");
  if (rc) return rc;
  const int64_t total = rows * 8;
  const int blocks2 = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  head_gather_tanh_kernel<<<blocks2, 256, 0, st>>>(proj, y, n, d, h, w_);
  return check_launch("head_gather_tanh_kernel");
}

int32_t petsyn_head_upconv_tanh_bwd(const void* x, const float* w, const float* y, const float* dy, float* dproj,
                                    void* dx, float* dw, int32_t n, int32_t d, int32_t h, int32_t w_, int32_t cin,
                                    void* stream) {
  PETSYN_REQUIRE(x && w && y && dy && dproj && dx && dw, "null argument");
  PETSYN_REQUIRE(cin % 8 == 0, "head: cin must be a multiple of 8");
  cudaStream_t st = as_stream(stream);
  const int64_t rows = (int64_t)n * d * h * w_;
  const int64_t total = rows * 32;
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  head_scatter_bwd_kernel<<<blocks, 256, 0, st>>>(y, dy, dproj, n, d, h, w_);
  int32_t rc = check_launch("head_scatter_bwd_kernel");
  if (rc) return rc;
  PETSYN_CHECK_CUDA(cudaMemsetAsync(dw, 0, (size_t)cin * 27 * sizeof(float), st));
  const int nblk = (int)std::min<int64_t>((rows + 31) / 32, 148 * 4);
  const int64_t rpb = ((rows + nblk - 1) / nblk + 31) / 32 * 32;
  const int nblk2 = (int)((rows + rpb - 1) / rpb);
  head_bwd_kernel<<<nblk2, std::min(cin, 256), 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(x), w, dproj,
                                                        reinterpret_cast<__nv_bfloat16*>(dx), dw, rows, cin, rpb);
  return check_launch("head_bwd_kernel");
}

}  // extern "C"
