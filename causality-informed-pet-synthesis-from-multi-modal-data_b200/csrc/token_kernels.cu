// Kernels for the covariate-conditioned generator (AttenUNet, unet/utils/atten_unet_model.py): 2x resampling inside
// the up/down ResnetBlocks (:646-654), and the token-stream ops of the level-3 SpatialTransformer (:65-343): LayerNorm,
// GEGLU, and the covariate injection that cross-attention over a length-1 context reduces to.  All SIMT:
// these are bandwidth/latency-bound ops on tensors of a few MB; the GEMMs around them (proj_in/out, to_q/k/v, to_out,
// MLP linears) run on the tcgen05 conv kernel as k=1 convolutions and self-attention lives in attention_mma.cu.
#include <cuda_bf16.h>

#include <algorithm>

#include <cstdlib>

#include "common.h"
#include "det_reduce.cuh"

namespace petsyn {
namespace tk {

__device__ __forceinline__ float bf(const __nv_bfloat16 v) { return __bfloat162float(v); }

// ------------------------------------------------------------------------------------------------ resampling
// dst[n, od, oh, ow, :] (+)= scale * sum_{2x2x2} src[n, 2od+a, 2oh+b, 2ow+c, :]   (AvgPool3d(2,2): scale = 1/8)
__global__ void __launch_bounds__(256) resample_down_kernel(const __nv_bfloat16* __restrict__ src, int css, int cos,
                                                            __nv_bfloat16* __restrict__ dst, int csd, int cod, int N,
                                                            int OD, int OH, int OW, int C, float scale, int accumulate) {
  pdl_sync();
  const int cpt = C / 8;
  const int64_t total = (int64_t)N * OD * OH * OW * cpt;
  const int ID = 2 * OD, IH = 2 * OH, IW = 2 * OW;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cpt);
    int64_t q = i / cpt;
    const int ow = (int)(q % OW); q /= OW;
    const int oh = (int)(q % OH); q /= OH;
    const int od = (int)(q % OD); q /= OD;
    const int n = (int)q;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int id = 2 * od + (t >> 2), ih = 2 * oh + ((t >> 1) & 1), iw = 2 * ow + (t & 1);
      const int64_t r = (((int64_t)n * ID + id) * IH + ih) * IW + iw;
      const uint4 raw = *reinterpret_cast<const uint4*>(src + r * css + cos + cg * 8);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(b2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    const int64_t ro = (((int64_t)n * OD + od) * OH + oh) * OW + ow;
    __nv_bfloat16* p = dst + ro * csd + cod + cg * 8;
    if (accumulate) {
      const uint4 raw = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(b2[j]);
        acc[2 * j] = acc[2 * j] * scale + f.x;
        acc[2 * j + 1] = acc[2 * j + 1] * scale + f.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] *= scale;
    }
    uint4 o;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[0], acc[1]), b1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 b2o = __floats2bfloat162_rn(acc[4], acc[5]), b3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2o); o.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(p) = o;
  }
}

// dst[n, od, oh, ow, :] (+)= scale * src[n, od/2, oh/2, ow/2, :]   (nearest x2 upsampling: scale = 1)
__global__ void __launch_bounds__(256) resample_up_kernel(const __nv_bfloat16* __restrict__ src, int css, int cos,
                                                          __nv_bfloat16* __restrict__ dst, int csd, int cod, int N,
                                                          int OD, int OH, int OW, int C, float scale, int accumulate) {
  pdl_sync();
  const int cpt = C / 8;
  const int64_t total = (int64_t)N * OD * OH * OW * cpt;
  const int ID = OD / 2, IH = OH / 2, IW = OW / 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % cpt);
    int64_t q = i / cpt;
    const int ow = (int)(q % OW); q /= OW;
    const int oh = (int)(q % OH); q /= OH;
    const int od = (int)(q % OD); q /= OD;
    const int n = (int)q;
    const int64_t r = (((int64_t)n * ID + (od >> 1)) * IH + (oh >> 1)) * IW + (ow >> 1);
    const uint4 raw = *reinterpret_cast<const uint4*>(src + r * css + cos + cg * 8);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(b2[j]);
      acc[2 * j] = f.x * scale;
      acc[2 * j + 1] = f.y * scale;
    }
    const int64_t ro = (((int64_t)n * OD + od) * OH + oh) * OW + ow;
    __nv_bfloat16* p = dst + ro * csd + cod + cg * 8;
    if (accumulate) {
      const uint4 rawd = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&rawd);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(d2[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint4 o;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(acc[0], acc[1]), b1 = __floats2bfloat162_rn(acc[2], acc[3]);
    __nv_bfloat162 b2o = __floats2bfloat162_rn(acc[4], acc[5]), b3 = __floats2bfloat162_rn(acc[6], acc[7]);
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2o); o.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(p) = o;
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm (warp per row)
// x, y: bf16 [rows, C] contiguous, C <= 1024 and C % 8 == 0 (lane l owns channels l, l + 32, ...; a last partial pass when
// C is not a multiple of 32)
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                            float* __restrict__ rstd, int64_t rows, int C, float eps) {
  pdl_sync();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int per = (C + 31) / 32;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    float v[32];
    float s = 0.f;
    for (int j = 0; j < per; ++j) {
      const int c = j * 32 + lane;
      v[j] = c < C ? bf(x[r * C + c]) : 0.f;
      s += v[j];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)C;
    float q = 0.f;
    for (int j = 0; j < per; ++j) { const float d = (j * 32 + lane < C) ? v[j] - mu : 0.f; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q / (float)C + eps);
    for (int j = 0; j < per; ++j) {
      const int c = j * 32 + lane;
      if (c < C) y[r * C + c] = __float2bfloat16((v[j] - mu) * rs * gamma[c] + beta[c]);
    }
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
  }
}

// dx (+)= rstd * (g*gamma - mean(g*gamma) - xhat * mean(g*gamma*xhat)); dgamma += sum g*xhat, dbeta += sum g (fp32 atomics)
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const __nv_bfloat16* __restrict__ dy,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ rstd,
                                                            __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int64_t rows, int C,
                                                            int accumulate, const DetWs ws, int overwrite) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];   // [warps][2][C] per-warp partials of dgamma / dbeta (>= 1024 floats)
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const int per = (C + 31) / 32;
  float pg[32], pb[32];
  for (int j = 0; j < per; ++j) pg[j] = pb[j] = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const float mu = mean[r], rs = rstd[r];
    float xh[32], gg[32];
    float s0 = 0.f, s1 = 0.f;
    for (int j = 0; j < per; ++j) {
      const int c = j * 32 + lane;
      const bool in = c < C;
      const float g = in ? bf(dy[r * C + c]) : 0.f;
      xh[j] = in ? (bf(x[r * C + c]) - mu) * rs : 0.f;
      gg[j] = in ? g * gamma[c] : 0.f;
      s0 += gg[j];
      s1 += gg[j] * xh[j];
      pg[j] += g * xh[j];
      pb[j] += g;
    }
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    s0 /= (float)C; s1 /= (float)C;
    for (int j = 0; j < per; ++j) {
      const int c = j * 32 + lane;
      if (c >= C) continue;
      float v = rs * (gg[j] - s0 - xh[j] * s1);
      if (accumulate) v += bf(dx[r * C + c]);
      dx[r * C + c] = __float2bfloat16(v);
    }
  }
  {
    float* mine = sm + (threadIdx.x >> 5) * 2 * C;
    for (int j = 0; j < per; ++j) {
      const int c = j * 32 + lane;
      if (c >= C) continue;
      mine[c] = pg[j];
      mine[C + c] = pb[j];
    }
  }
  __syncthreads();
  // warps in warp order, CTAs in slot order (det_reduce.cuh): reproducible dgamma / dbeta
  det_cta_reduce(
      ws, 2 * C, sm,
      [&](int e) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += sm[w * 2 * C + e];
        return s;
      },
      [&](int e, float t, bool atomic) {
        float* p = e < C ? dgamma + e : dbeta + (e - C);
        if (atomic) atomicAdd(p, t); else if (overwrite) *p = t; else *p += t;
      });
}

// The same for C = NCH * 128: lane l owns channels [4 l, 4 l + 4) of every 128-channel chunk -- 8-byte accesses, everything
// in registers (the transformer width of the reference's configs is 128)
template <int NCH>
__global__ void __launch_bounds__(256) layernorm_bwd_vec_kernel(const __nv_bfloat16* __restrict__ x,
                                                                const __nv_bfloat16* __restrict__ dy,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma,
                                                                float* __restrict__ dbeta, int64_t rows, int accumulate,
                                                                const DetWs ws, int overwrite) {
  pdl_sync();
  constexpr int C = NCH * 128;
  extern __shared__ __align__(16) float sm[];   // [warps][2][C]
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  float ga[NCH][4], pg[NCH][4], pb[NCH][4];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const float4 g4 = *reinterpret_cast<const float4*>(gamma + k * 128 + lane * 4);
    ga[k][0] = g4.x; ga[k][1] = g4.y; ga[k][2] = g4.z; ga[k][3] = g4.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) pg[k][i] = pb[k][i] = 0.f;
  }
  auto ld4 = [](const __nv_bfloat16* p, float (&f)[4]) {
    const uint2 raw = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  };
  for (int64_t r = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const float mu = mean[r], rs = rstd[r];
    float xh[NCH][4], gg[NCH][4];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      float xv[4], gv[4];
      ld4(x + r * C + k * 128 + lane * 4, xv);
      ld4(dy + r * C + k * 128 + lane * 4, gv);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        xh[k][i] = (xv[i] - mu) * rs;
        gg[k][i] = gv[i] * ga[k][i];
        s0 += gg[k][i];
        s1 += gg[k][i] * xh[k][i];
        pg[k][i] += gv[i] * xh[k][i];
        pb[k][i] += gv[i];
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    s0 *= 1.f / (float)C; s1 *= 1.f / (float)C;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = rs * (gg[k][i] - s0 - xh[k][i] * s1);
      __nv_bfloat16* p = dx + r * C + k * 128 + lane * 4;
      if (accumulate) {
        float old[4];
        ld4(p, old);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] += old[i];
      }
      uint2 o;
      __nv_bfloat162 b0 = __floats2bfloat162_rn(v[0], v[1]), b1 = __floats2bfloat162_rn(v[2], v[3]);
      o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
      *reinterpret_cast<uint2*>(p) = o;
    }
  }
  {
    float* mine = sm + (threadIdx.x >> 5) * 2 * C;
#pragma unroll
    for (int k = 0; k < NCH; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        mine[k * 128 + lane * 4 + i] = pg[k][i];
        mine[C + k * 128 + lane * 4 + i] = pb[k][i];
      }
  }
  __syncthreads();
  det_cta_reduce(
      ws, 2 * C, sm,
      [&](int e) {
        float s = 0.f;
        for (int w = 0; w < wpb; ++w) s += sm[w * 2 * C + e];
        return s;
      },
      [&](int e, float t, bool atomic) {
        float* p = e < C ? dgamma + e : dbeta + (e - C);
        if (atomic) atomicAdd(p, t); else if (overwrite) *p = t; else *p += t;
      });
}

// ------------------------------------------------------------------------------------------------ GEGLU
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
// h: [rows, 2F] = (x | gate); out [rows, F] = x * gelu(gate)     (MONAI MLPBlock act="GEGLU")
// One thread = 8 consecutive channels (16-byte accesses); F must be a multiple of 8.
__device__ __forceinline__ void unpack8bf(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = __bfloat1622float2(b2[i]);
    f[2 * i] = v.x;
    f[2 * i + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8bf(const float (&f)[8]) {
  uint4 o;
  __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
  __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
  o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
  o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
  return o;
}
__global__ void __launch_bounds__(256) geglu_fwd_kernel(const __nv_bfloat16* __restrict__ h,
                                                        __nv_bfloat16* __restrict__ out, int64_t rows, int F) {
  pdl_sync();
  const int f8 = F / 8;
  const int64_t total = rows * f8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / f8;
    const int c = (int)(i - r * f8) * 8;
    float x[8], g[8], o[8];
    unpack8bf(*reinterpret_cast<const uint4*>(h + r * 2 * F + c), x);
    unpack8bf(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + c), g);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = x[q] * gelu_f(g[q]);
    *reinterpret_cast<uint4*>(out + r * F + c) = pack8bf(o);
  }
}
__global__ void __launch_bounds__(256) geglu_bwd_kernel(const __nv_bfloat16* __restrict__ h,
                                                        const __nv_bfloat16* __restrict__ dout,
                                                        __nv_bfloat16* __restrict__ dh, int64_t rows, int F) {
  pdl_sync();
  const int f8 = F / 8;
  const int64_t total = rows * f8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / f8;
    const int c = (int)(i - r * f8) * 8;
    float x[8], g[8], d[8], dx[8], dg[8];
    unpack8bf(*reinterpret_cast<const uint4*>(h + r * 2 * F + c), x);
    unpack8bf(*reinterpret_cast<const uint4*>(h + r * 2 * F + F + c), g);
    unpack8bf(*reinterpret_cast<const uint4*>(dout + r * F + c), d);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      dx[q] = d[q] * gelu_f(g[q]);
      dg[q] = d[q] * x[q] * gelu_grad(g[q]);
    }
    *reinterpret_cast<uint4*>(dh + r * 2 * F + c) = pack8bf(dx);
    *reinterpret_cast<uint4*>(dh + r * 2 * F + F + c) = pack8bf(dg);
  }
}

// ------------------------------------------------------------------------------------------------ covariate injection
// Cross-attention over a length-1 context: softmax over one key == 1, so attn2(x, ctx) = to_out(to_v(ctx)) for every
// token (atten_unet_model.py:156-175).  bias[n, :] = Wo (Wv ctx[n]) + bo, then t[n, l, :] += bias[n, :].
// grid (sample, groups of 8 output channels), 256 threads: every CTA recomputes v = Wv ctx[n] (C x Cctx products), then a warp
// per output channel strides its lanes over the row of Wo (coalesced) and folds them with a butterfly
__global__ void __launch_bounds__(256) covariate_bias_kernel(const float* __restrict__ ctx, const float* __restrict__ wv,
                                                             const float* __restrict__ wo, const float* __restrict__ bo,
                                                             float* __restrict__ vbuf, float* __restrict__ bias, int N,
                                                             int Cctx, int C) {
  pdl_sync();
  extern __shared__ float sv[];   // [C]
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < Cctx; ++j) s += wv[c * Cctx + j] * ctx[n * Cctx + j];
    sv[c] = s;
    if (blockIdx.y == 0) vbuf[n * C + c] = s;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.y * 8 + warp;
  if (c < C) {
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += wo[c * C + j] * sv[j];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) bias[n * C + c] = s + bo[c];
  }
}
__global__ void __launch_bounds__(256) add_sample_bias_kernel(__nv_bfloat16* __restrict__ t, const float* __restrict__ bias,
                                                              int64_t rows_per_sample, int64_t rows, int C) {
  pdl_sync();
  const int64_t total = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / C;
    const int c = (int)(i - r * C);
    t[i] = __float2bfloat16(bf(t[i]) + bias[(r / rows_per_sample) * C + c]);
  }
}
// dbias[n, c] = sum over the sample's rows of dt[., c]
__global__ void __launch_bounds__(256) sample_colsum_kernel(const __nv_bfloat16* __restrict__ dt, float* __restrict__ out,
                                                            int64_t rows_per_sample, int C, const DetWs ws) {
  pdl_sync();
  __shared__ __align__(16) float part[1024];
  const int n = blockIdx.y;
  const int c = threadIdx.x % C;
  const int rl = threadIdx.x / C, rpp = blockDim.x / C;
  float acc = 0.f;
  if (rl < rpp)
    for (int64_t r = (int64_t)blockIdx.x * rpp + rl; r < rows_per_sample; r += (int64_t)gridDim.x * rpp)
      acc += bf(dt[((int64_t)n * rows_per_sample + r) * C + c]);
  part[threadIdx.x] = rl < rpp ? acc : 0.f;
  __syncthreads();
  det_cta_reduce(
      ws, C, part,
      [&](int e) {
        float s = 0.f;
        for (int r = 0; r < rpp; ++r) s += part[r * C + e];
        return s;
      },
      [&](int e, float t, bool atomic) {
        if (atomic) atomicAdd(out + n * C + e, t); else out[n * C + e] += t;
      });
}
// gradients of the tiny linears: dbo = sum_n dbias; dWo = sum_n dbias (x) v; dv = Wo^T dbias; dWv = sum_n dv (x) ctx
__global__ void __launch_bounds__(128) covariate_bias_bwd_kernel(const float* __restrict__ ctx, const float* __restrict__ wo,
                                                                 const float* __restrict__ vbuf,
                                                                 const float* __restrict__ dbias, float* __restrict__ dwv,
                                                                 float* __restrict__ dwo, float* __restrict__ dbo, int N,
                                                                 int Cctx, int C) {
  pdl_sync();
  // one block per channel c: row c of dWo, dbo[c], dv[:, c] and row c of dWv -- no dependency between blocks
  extern __shared__ float sdv[];  // [N]: dv[n, c] = sum_c' Wo[c', c] * dbias[n, c']
  const int c = blockIdx.x;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dbias[n * C + c] * vbuf[n * C + j];
    dwo[c * C + j] = s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += dbias[n * C + c];
    dbo[c] = s;
  }
  // dv[n, c]: warp w handles samples w, w + nwarps, ...; lanes split the c' sum
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int n = warp; n < N; n += nwarps) {
    float s = 0.f;
    for (int cc = lane; cc < C; cc += 32) s += wo[cc * C + c] * dbias[n * C + cc];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sdv[n] = s;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cctx; j += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += sdv[n] * ctx[n * Cctx + j];
    dwv[c * Cctx + j] = s;
  }
}

static int blocks_for(int64_t total, int per_block = 256, int cap = 148 * 16) {
  if (const char* e = getenv("PETSYN_TK_WAVES")) cap = 148 * std::max(1, atoi(e));      // tuning experiments only
  return (int)std::max<int64_t>(1, std::min<int64_t>((total + per_block - 1) / per_block, cap));
}

}  // namespace tk
}  // namespace petsyn

using namespace petsyn;
using namespace petsyn::tk;
#define BFP(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBFP(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int32_t petsyn_resample2(const void* src, int32_t src_cstride, int32_t src_coff, void* dst, int32_t dst_cstride,
                         int32_t dst_coff, int32_t n, int32_t od, int32_t oh, int32_t ow, int32_t c, int32_t up,
                         float scale, int32_t accumulate, void* stream) {
  PETSYN_REQUIRE(src && dst && n > 0 && od > 0 && oh > 0 && ow > 0, "bad argument");
  PETSYN_REQUIRE(c % 8 == 0 && (src_cstride | src_coff | dst_cstride | dst_coff) % 8 == 0,
                 "channels, pitches and offsets must be multiples of 8");
  PETSYN_REQUIRE(!up || ((od | oh | ow) & 1) == 0, "upsampled dims must be even");
  const int64_t total = (int64_t)n * od * oh * ow * (c / 8);
  if (up)
    PETSYN_CHECK_CUDA(launch_pdl(resample_up_kernel, dim3(blocks_for(total)), dim3(256), 0, as_stream(stream), CBFP(src), src_cstride, src_coff, BFP(dst),
                                                                         dst_cstride, dst_coff, n, od, oh, ow, c, scale,
                                                                         accumulate));
  else
    PETSYN_CHECK_CUDA(launch_pdl(resample_down_kernel, dim3(blocks_for(total)), dim3(256), 0, as_stream(stream), CBFP(src), src_cstride, src_coff, BFP(dst),
                                                                           dst_cstride, dst_coff, n, od, oh, ow, c, scale,
                                                                           accumulate));
  return check_launch("resample kernel");
}

int32_t petsyn_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             int64_t rows, int32_t c, float eps, void* stream) {
  PETSYN_REQUIRE(x && gamma && beta && y && mean && rstd && rows > 0, "bad argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 1024, "LayerNorm width must be a multiple of 8, at most 1024");
  PETSYN_CHECK_CUDA(launch_pdl(layernorm_fwd_kernel, dim3(blocks_for(rows, 8)), dim3(256), 0, as_stream(stream), CBFP(x), gamma, beta, BFP(y), mean, rstd, rows,
                                                                           c, eps));
  return check_launch("layernorm_fwd_kernel");
}

int32_t petsyn_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                             void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t accumulate_dx,
                             void* stream) {
  PETSYN_REQUIRE(x && dy && gamma && mean && rstd && dx && dgamma && dbeta && rows > 0, "bad argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 1024, "LayerNorm width must be a multiple of 8, at most 1024");
  cudaStream_t st = as_stream(stream);
  DetWs ws;
  {
    int32_t rcw = det_workspace(&ws);
    if (rcw) return rcw;
  }
  // reproducible mode: the last CTA of the reduction writes the totals (no clearing needed); float-atomic mode adds into zeros
  const int overwrite = ws.acc != nullptr ? 1 : 0;
  if (!overwrite) {
    PETSYN_CHECK_CUDA(cudaMemsetAsync(dgamma, 0, c * sizeof(float), st));
    PETSYN_CHECK_CUDA(cudaMemsetAsync(dbeta, 0, c * sizeof(float), st));
  }
  const size_t ln_smem = std::max<size_t>((size_t)8 * 2 * c, 1024) * sizeof(float);
  if (ln_smem > 48 * 1024)
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(layernorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ln_smem));
  if (c == 128) {
    PETSYN_CHECK_CUDA(launch_pdl(layernorm_bwd_vec_kernel<1>, dim3(blocks_for(rows, 8, 148)), dim3(256), ln_smem, st, CBFP(x), CBFP(dy),
                                 gamma, mean, rstd, BFP(dx), dgamma, dbeta, rows, accumulate_dx, ws, overwrite));
  } else if (c == 256) {
    PETSYN_CHECK_CUDA(launch_pdl(layernorm_bwd_vec_kernel<2>, dim3(blocks_for(rows, 8, 148)), dim3(256), ln_smem, st, CBFP(x), CBFP(dy),
                                 gamma, mean, rstd, BFP(dx), dgamma, dbeta, rows, accumulate_dx, ws, overwrite));
  } else {
    PETSYN_CHECK_CUDA(launch_pdl(layernorm_bwd_kernel, dim3(blocks_for(rows, 8, 148)), dim3(256), ln_smem, st, 
      CBFP(x), CBFP(dy), gamma, mean, rstd, BFP(dx), dgamma, dbeta, rows, c, accumulate_dx, ws, overwrite));
  }
  return check_launch("layernorm_bwd_kernel");
}

int32_t petsyn_geglu_fwd(const void* h, void* out, int64_t rows, int32_t f, void* stream) {
  PETSYN_REQUIRE(h && out && rows > 0 && f > 0 && f % 8 == 0, "bad argument (the gated width must be a multiple of 8)");
  PETSYN_CHECK_CUDA(launch_pdl(geglu_fwd_kernel, dim3(blocks_for(rows * (f / 8))), dim3(256), 0, as_stream(stream), CBFP(h), BFP(out), rows, f));
  return check_launch("geglu_fwd_kernel");
}

int32_t petsyn_geglu_bwd(const void* h, const void* dout, void* dh, int64_t rows, int32_t f, void* stream) {
  PETSYN_REQUIRE(h && dout && dh && rows > 0 && f > 0 && f % 8 == 0, "bad argument (the gated width must be a multiple of 8)");
  PETSYN_CHECK_CUDA(launch_pdl(geglu_bwd_kernel, dim3(blocks_for(rows * (f / 8))), dim3(256), 0, as_stream(stream), CBFP(h), CBFP(dout), BFP(dh), rows, f));
  return check_launch("geglu_bwd_kernel");
}

int32_t petsyn_covariate_bias_fwd(const float* ctx, const float* wv, const float* wo, const float* bo, float* vbuf,
                                  float* bias, void* tokens, int32_t n, int32_t cctx, int32_t c,
                                  int64_t rows_per_sample, void* stream) {
  PETSYN_REQUIRE(ctx && wv && wo && bo && vbuf && bias && tokens && n > 0 && cctx > 0 && c > 0, "bad argument");
  cudaStream_t st = as_stream(stream);
  PETSYN_CHECK_CUDA(launch_pdl(covariate_bias_kernel, dim3(n, (c + 7) / 8), dim3(256), c * sizeof(float), st, ctx, wv, wo, bo, vbuf, bias, n, cctx, c));
  int32_t rc = check_launch("covariate_bias_kernel");
  if (rc) return rc;
  PETSYN_CHECK_CUDA(launch_pdl(add_sample_bias_kernel, dim3(blocks_for(rows_per_sample * n * c)), dim3(256), 0, st, BFP(tokens), bias, rows_per_sample,
                                                                            rows_per_sample * n, c));
  return check_launch("add_sample_bias_kernel");
}

int32_t petsyn_covariate_bias_bwd(const float* ctx, const float* wo, const float* vbuf, const void* dtokens,
                                  float* dbias, float* dwv, float* dwo, float* dbo, int32_t n, int32_t cctx, int32_t c,
                                  int64_t rows_per_sample, void* stream) {
  PETSYN_REQUIRE(ctx && wo && vbuf && dtokens && dbias && dwv && dwo && dbo && n > 0, "bad argument");
  PETSYN_REQUIRE(c <= 256 && 256 % c == 0 && (size_t)n * c * sizeof(float) <= 48 * 1024, "unsupported width/batch");
  cudaStream_t st = as_stream(stream);
  PETSYN_CHECK_CUDA(cudaMemsetAsync(dbias, 0, (size_t)n * c * sizeof(float), st));
  dim3 grid((unsigned)std::min<int64_t>(64, (rows_per_sample + 1) / 2), (unsigned)n);
  DetWs ws;
  {
    int32_t rcw = det_workspace(&ws);
    if (rcw) return rcw;
  }
  PETSYN_CHECK_CUDA(launch_pdl(sample_colsum_kernel, dim3(grid), dim3(256), 0, st, CBFP(dtokens), dbias, rows_per_sample, c, ws));
  int32_t rc = check_launch("sample_colsum_kernel");
  if (rc) return rc;
  PETSYN_CHECK_CUDA(launch_pdl(covariate_bias_bwd_kernel, dim3(c), dim3(128), (size_t)n * sizeof(float), st, ctx, wo, vbuf, dbias, dwv, dwo, dbo, n, cctx, c));
  return check_launch("covariate_bias_bwd_kernel");
}

}  // extern "C"
