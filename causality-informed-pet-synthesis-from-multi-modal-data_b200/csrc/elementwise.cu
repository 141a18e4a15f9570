// Bandwidth-bound kernels: normalisation statistics, fused normalise + activation + skip-concat writes and their
// backward passes, losses, Adam.  All activations are NDHWC bf16; each thread moves 16 bytes (8 channels) per access,
// a warp covers consecutive channels of consecutive voxels (fully coalesced), reductions go registers -> shared
// memory -> one atomic per (block, channel).
#include <cuda_bf16.h>

#include <algorithm>

#include <cstdlib>

#include "common.h"
#include "det_reduce.cuh"

namespace petsyn {

struct F8 {
  float v[8];
};

__device__ __forceinline__ F8 load8(const __nv_bfloat16* p) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
  F8 r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(b2[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const F8& r) {
  uint4 o;
  __nv_bfloat162 b0 = __floats2bfloat162_rn(r.v[0], r.v[1]), b1 = __floats2bfloat162_rn(r.v[2], r.v[3]);
  __nv_bfloat162 b2 = __floats2bfloat162_rn(r.v[4], r.v[5]), b3 = __floats2bfloat162_rn(r.v[6], r.v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
  o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
  *reinterpret_cast<uint4*>(p) = o;
}
__device__ __forceinline__ F8 load8f(const float* p) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

__device__ __forceinline__ float act_fwd(float b, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return fmaxf(b, 0.f);
    case PETSYN_ACT_LRELU: return b > 0.f ? b : b * slope;
    case PETSYN_ACT_SILU: return b / (1.f + __expf(-b));
    case PETSYN_ACT_TANH: return tanhf(b);
    default: return b;
  }
}
__device__ __forceinline__ float act_grad(float b, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return b > 0.f ? 1.f : 0.f;
    case PETSYN_ACT_LRELU: return b > 0.f ? 1.f : slope;
    case PETSYN_ACT_SILU: { const float s = 1.f / (1.f + __expf(-b)); return s * (1.f + b * (1.f - s)); }
    case PETSYN_ACT_TANH: { const float t = tanhf(b); return 1.f - t * t; }
    default: return 1.f;
  }
}

constexpr int kRowUnroll = 2;   // independent 16-byte loads in flight per thread

// Thread layout shared by the row-streaming kernels: cpt = C/8 threads span the channels, blockDim/cpt rows in flight.
struct RowIter {
  int cpt, rpp, tx, ty;
  bool active;
  __device__ RowIter(int C) {
    cpt = C / 8;
    rpp = blockDim.x / cpt;
    tx = threadIdx.x % cpt;
    ty = threadIdx.x / cpt;
    active = ty < rpp;
  }
};

// Block-level reduction of per-thread 8-channel partials (NV values each) over the rows dimension, then one 64-bit atomic
// per channel into the DOUBLE accumulator out[v*C + c] (exact, hence order-free, summation of the fp32 partials: see
// det_reduce.cuh).
template <int NV>
__device__ __forceinline__ void block_reduce_channels(const RowIter& it, float (&acc)[NV][8], float* smem, double* out,
                                                      int C) {
  // smem: [rpp][NV][C]
  if (it.active) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) smem[(it.ty * NV + v) * C + it.tx * 8 + i] = acc[v][i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NV * C; e += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < it.rpp; ++r) s += smem[r * NV * C + e];
    atomicAdd(out + e, (double)s);
  }
}

// one scalar per CTA (thread 0 holds it) combined the same way into *out
__device__ __forceinline__ void scalar_reduce(float s, float* out, const DetWs& ws) {
  __shared__ float scratch[256];
  det_cta_reduce(
      ws, 1, scratch, [&](int) { return s; },
      [&](int, float t, bool atomic) {
        if (atomic) atomicAdd(out, t); else *out += t;
      });
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const __nv_bfloat16* __restrict__ z, double* __restrict__ sums,
                                                       int64_t rows, int C) {
  extern __shared__ float smem_f[];
  RowIter it(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (it.active) {
    const int64_t stride = (int64_t)gridDim.x * it.rpp;
    for (int64_t r = (int64_t)blockIdx.x * it.rpp + it.ty; r < rows; r += kRowUnroll * stride) {
      F8 x[kRowUnroll];
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u)
        if (r + u * stride < rows) x[u] = load8(z + (r + u * stride) * C + it.tx * 8);
#pragma unroll
      for (int u = 0; u < kRowUnroll; ++u)
        if (r + u * stride < rows) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            acc[0][i] += x[u].v[i];
            acc[1][i] += x[u].v[i] * x[u].v[i];
          }
        }
    }
  }
  block_reduce_channels<2>(it, acc, smem_f, sums, C);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale,
                                   float* __restrict__ shift, float* __restrict__ save_mean,
                                   float* __restrict__ save_rstd, int64_t rows, int C, float eps, float momentum,
                                   int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double mean, var;
  if (training) {
    mean = (double)sums[c] / (double)rows;
    var = (double)sums[C + c] / (double)rows - mean * mean;
    if (var < 0) var = 0;
    if (running_mean) {
      const double unbiased = rows > 1 ? var * (double)rows / (double)(rows - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * mean);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
  } else {
    mean = running_mean[c];
    var = running_var[c];
  }
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale[c] = (float)(g * rstd);
  shift[c] = (float)(b - mean * g * rstd);
  if (save_mean) save_mean[c] = (float)mean;
  if (save_rstd) save_rstd[c] = (float)rstd;
}

__global__ void __launch_bounds__(256) norm_act_fwd_kernel(const __nv_bfloat16* __restrict__ z,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift,
                                                           __nv_bfloat16* __restrict__ dst1, int cs1, int co1, int act1,
                                                           __nv_bfloat16* __restrict__ dst2, int cs2, int co2, int act2,
                                                           float slope, int64_t rows, int C) {
  RowIter it(C);
  if (!it.active) return;
  F8 sc, sh;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc.v[i] = 1.f; sh.v[i] = 0.f; }
  if (scale) { sc = load8f(scale + it.tx * 8); sh = load8f(shift + it.tx * 8); }
  const int64_t stride = (int64_t)gridDim.x * it.rpp;
  for (int64_t r0 = (int64_t)blockIdx.x * it.rpp + it.ty; r0 < rows; r0 += kRowUnroll * stride) {
    F8 xs[kRowUnroll];
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u)
      if (r0 + u * stride < rows) xs[u] = load8(z + (r0 + u * stride) * C + it.tx * 8);
#pragma unroll
    for (int u = 0; u < kRowUnroll; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      F8 o1, o2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float b = xs[u].v[i] * sc.v[i] + sh.v[i];
        o1.v[i] = act_fwd(b, act1, slope);
        o2.v[i] = act_fwd(b, act2, slope);
      }
      store8(dst1 + r * cs1 + co1 + it.tx * 8, o1);
      if (dst2) store8(dst2 + r * cs2 + co2 + it.tx * 8, o2);
    }
  }
}

// g = g1*act1'(b) + g2*act2'(b);  per-channel sum g and sum g*zhat
__global__ void __launch_bounds__(256) norm_act_bwd_reduce_kernel(
    const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ rstd, const __nv_bfloat16* __restrict__ g1, int cs1,
    int co1, int act1, const __nv_bfloat16* __restrict__ g2, int cs2, int co2, int act2, float slope,
    double* __restrict__ sums, int64_t rows, int C) {
  extern __shared__ float smem_f[];
  RowIter it(C);
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (it.active) {
    F8 sc, sh, mu, rs;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc.v[i] = 1.f; sh.v[i] = 0.f; mu.v[i] = 0.f; rs.v[i] = 1.f; }
    if (scale) { sc = load8f(scale + it.tx * 8); sh = load8f(shift + it.tx * 8); }
    if (mean) { mu = load8f(mean + it.tx * 8); rs = load8f(rstd + it.tx * 8); }
    const int64_t stride = (int64_t)gridDim.x * it.rpp;
    for (int64_t r0 = (int64_t)blockIdx.x * it.rpp + it.ty; r0 < rows; r0 += 2 * stride) {
      F8 xs[2], as[2], bs[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t r = r0 + u * stride;
        if (r < rows) {
          xs[u] = load8(z + r * C + it.tx * 8);
          as[u] = load8(g1 + r * cs1 + co1 + it.tx * 8);
          if (g2) bs[u] = load8(g2 + r * cs2 + co2 + it.tx * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (r0 + u * stride >= rows) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float b = xs[u].v[i] * sc.v[i] + sh.v[i];
          float g = as[u].v[i] * act_grad(b, act1, slope);
          if (g2) g += bs[u].v[i] * act_grad(b, act2, slope);
          acc[0][i] += g;
          acc[1][i] += g * (xs[u].v[i] - mu.v[i]) * rs.v[i];
        }
      }
    }
  }
  block_reduce_channels<2>(it, acc, smem_f, sums, C);
}

__global__ void __launch_bounds__(256) norm_act_bwd_apply_kernel(
    const __nv_bfloat16* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
    const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
    const __nv_bfloat16* __restrict__ g1, int cs1, int co1, int act1, const __nv_bfloat16* __restrict__ g2, int cs2,
    int co2, int act2, float slope, const double* __restrict__ sums, __nv_bfloat16* __restrict__ dz,
    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C) {
  RowIter it(C);
  if (blockIdx.x == 0 && dgamma != nullptr) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      dbeta[c] = (float)sums[c];
      dgamma[c] = (float)sums[C + c];
    }
  }
  if (!it.active) return;
  F8 sc, sh, mu, rs, k0, k1, k2;
#pragma unroll
  for (int i = 0; i < 8; ++i) { sc.v[i] = 1.f; sh.v[i] = 0.f; mu.v[i] = 0.f; rs.v[i] = 1.f; k0.v[i] = 1.f; k1.v[i] = 0.f; k2.v[i] = 0.f; }
  if (scale) { sc = load8f(scale + it.tx * 8); sh = load8f(shift + it.tx * 8); }
  if (mean) {
    mu = load8f(mean + it.tx * 8);
    rs = load8f(rstd + it.tx * 8);
    const F8 ga = gamma ? load8f(gamma + it.tx * 8) : k0;
    F8 s0, s1;
#pragma unroll
    for (int i = 0; i < 8; ++i) { s0.v[i] = (float)sums[it.tx * 8 + i]; s1.v[i] = (float)sums[C + it.tx * 8 + i]; }
    const float inv = 1.f / (float)rows;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      k0.v[i] = ga.v[i] * rs.v[i];          // dz = k0 * (g - k1 - zhat * k2)
      k1.v[i] = s0.v[i] * inv;
      k2.v[i] = s1.v[i] * inv;
    }
  }
  const int64_t stride = (int64_t)gridDim.x * it.rpp;
  for (int64_t r0 = (int64_t)blockIdx.x * it.rpp + it.ty; r0 < rows; r0 += 2 * stride) {
    F8 xs[2], as[2], bs[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
      if (r < rows) {
        xs[u] = load8(z + r * C + it.tx * 8);
        as[u] = load8(g1 + r * cs1 + co1 + it.tx * 8);
        if (g2) bs[u] = load8(g2 + r * cs2 + co2 + it.tx * 8);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int64_t r = r0 + u * stride;
      if (r >= rows) break;
      F8 o;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float b = xs[u].v[i] * sc.v[i] + sh.v[i];
        float g = as[u].v[i] * act_grad(b, act1, slope);
        if (g2) g += bs[u].v[i] * act_grad(b, act2, slope);
        const float zh = (xs[u].v[i] - mu.v[i]) * rs.v[i];
        o.v[i] = k0.v[i] * (g - k1.v[i] - zh * k2.v[i]);
      }
      store8(dz + r * C + it.tx * 8, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------ losses / optimiser
__device__ __forceinline__ float block_sum(float v, float* smem) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0.f;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;   // valid in warp 0
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ y, const float* __restrict__ t,
                                                      float* __restrict__ loss, float* __restrict__ dy, int64_t n,
                                                      float inv_n, float gscale, const DetWs ws) {
  pdl_sync();
  __shared__ float red[8];
  float acc = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(y)[i], b = reinterpret_cast<const float4*>(t)[i];
    const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
    acc += fabsf(d0) + fabsf(d1) + fabsf(d2) + fabsf(d3);
    if (dy) {
      float4 g;
      g.x = (d0 > 0.f) - (d0 < 0.f); g.y = (d1 > 0.f) - (d1 < 0.f);
      g.z = (d2 > 0.f) - (d2 < 0.f); g.w = (d3 > 0.f) - (d3 < 0.f);
      g.x *= gscale; g.y *= gscale; g.z *= gscale; g.w *= gscale;
      reinterpret_cast<float4*>(dy)[i] = g;
    }
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float d = y[i] - t[i];
      acc += fabsf(d);
      if (dy) dy[i] = gscale * ((d > 0.f) - (d < 0.f));
    }
  const float s = block_sum(acc, red);
  scalar_reduce(s * inv_n, loss, ws);
}

__global__ void __launch_bounds__(256) mse_const_kernel(const float* __restrict__ x, float target,
                                                        float* __restrict__ loss, float* __restrict__ dx, int64_t n,
                                                        float inv_n, float gscale, const DetWs ws) {
  pdl_sync();
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = x[i] - target;
    acc += d * d;
    if (dx) dx[i] = 2.f * d * gscale;
  }
  const float s = block_sum(acc, red);
  scalar_reduce(s * inv_n, loss, ws);
}

// step_dev (optional): device-resident 1-based step counter, so that a CUDA graph of the training step stays valid as
// the bias corrections change from step to step.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                                                   float b1, float b2, float eps, int step,
                                                   const int32_t* __restrict__ step_dev) {
  pdl_sync();
  const float t = (float)(step_dev ? *step_dev : step);
  const float bc1 = 1.f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, t));
  const float step_size = lr / bc1;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 gi = reinterpret_cast<const float4*>(g)[i];
    float4 mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i], pi = reinterpret_cast<float4*>(p)[i];
#define PETSYN_ADAM1(c)                                        \
    mi.c = b1 * mi.c + (1.f - b1) * gi.c;                      \
    vi.c = b2 * vi.c + (1.f - b2) * gi.c * gi.c;               \
    pi.c -= step_size * (mi.c / (sqrtf(vi.c) / bc2_sqrt + eps));
    PETSYN_ADAM1(x) PETSYN_ADAM1(y) PETSYN_ADAM1(z) PETSYN_ADAM1(w)
#undef PETSYN_ADAM1
    reinterpret_cast<float4*>(m)[i] = mi;
    reinterpret_cast<float4*>(v)[i] = vi;
    reinterpret_cast<float4*>(p)[i] = pi;
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) {
      const float gi = g[i];
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      p[i] -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    }
}

// one block; n*dim is tiny (batch x 8 latent dims)
__global__ void __launch_bounds__(256) kl_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                                 float* __restrict__ loss, float* __restrict__ dmu,
                                                 float* __restrict__ dlv, int n, int dim, int pitch, float gscale) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n * dim; i += blockDim.x) {
    const int r = i / dim, c = i % dim;
    const float m = mu[r * pitch + c], l = lv[r * pitch + c], e = __expf(l);
    acc += -0.5f * (1.f + l - m * m - e);
    if (dmu) dmu[r * pitch + c] = gscale * m / (float)n;
    if (dlv) dlv[r * pitch + c] = gscale * (-0.5f) * (1.f - e) / (float)n;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, s / (float)n);
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, float* __restrict__ out, int64_t n,
                                                    const DetWs ws) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    acc += g[i] * g[i];
  const float s = block_sum(acc, red);
  scalar_reduce(s, out, ws);
}

static int row_blocks(int64_t rows, int C) {
  const int rpp = 256 / (C / 8);
  static int waves = 0;                       // persistent CTAs per SM (PETSYN_BN_WAVES: tuning experiments)
  if (waves == 0) {
    const char* e = getenv("PETSYN_BN_WAVES");
    waves = e != nullptr ? std::max(1, atoi(e)) : 2;        // measured on configs[0]: 8 -> 2.69 ms per step, 4 -> 2.63, 2 -> 2.60, 1 -> 2.83
  }
  return (int)std::max<int64_t>(1, std::min<int64_t>((rows + rpp - 1) / rpp, 148 * waves));
}

}  // namespace petsyn

using namespace petsyn;

#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" {

int32_t petsyn_bn_stats(const void* z, double* sums, int64_t rows, int32_t c, void* stream) {
  PETSYN_REQUIRE(z && sums, "null argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  const int rpp = 256 / (c / 8);
  const size_t smem = (size_t)rpp * 2 * c * sizeof(float);
  PETSYN_CHECK_CUDA(cudaFuncSetAttribute(bn_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  bn_stats_kernel<<<row_blocks(rows, c), 256, smem, as_stream(stream)>>>(CBF(z), sums, rows, c);
  return check_launch("bn_stats_kernel");
}

int32_t petsyn_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float* scale, float* shift, float* save_mean, float* save_rstd,
                           int64_t rows, int32_t c, float eps, float momentum, int32_t training, void* stream) {
  PETSYN_REQUIRE(scale && shift, "null argument");
  PETSYN_REQUIRE(training ? sums != nullptr : (running_mean && running_var), "missing statistics");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(sums, gamma, beta, running_mean, running_var, scale,
                                                                     shift, save_mean, save_rstd, rows, c, eps,
                                                                     momentum, training);
  return check_launch("bn_finalize_kernel");
}

int32_t petsyn_norm_act_fwd(const void* z, const float* scale, const float* shift, void* dst1, int32_t dst1_cstride,
                            int32_t dst1_coff, int32_t act1, void* dst2, int32_t dst2_cstride, int32_t dst2_coff,
                            int32_t act2, float slope, int64_t rows, int32_t c, void* stream) {
  PETSYN_REQUIRE(z && dst1, "null argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  PETSYN_REQUIRE(dst1_cstride % 8 == 0 && dst1_coff % 8 == 0 && dst2_cstride % 8 == 0 && dst2_coff % 8 == 0,
                 "channel pitches/offsets must be multiples of 8");
  norm_act_fwd_kernel<<<row_blocks(rows, c), 256, 0, as_stream(stream)>>>(CBF(z), scale, shift, BF(dst1), dst1_cstride,
                                                                          dst1_coff, act1, BF(dst2), dst2_cstride,
                                                                          dst2_coff, act2, slope, rows, c);
  return check_launch("norm_act_fwd_kernel");
}

int32_t petsyn_norm_act_bwd_reduce(const void* z, const float* scale, const float* shift, const float* mean,
                                   const float* rstd, const void* g1, int32_t g1_cstride, int32_t g1_coff,
                                   int32_t act1, const void* g2, int32_t g2_cstride, int32_t g2_coff, int32_t act2,
                                   float slope, double* sums, int64_t rows, int32_t c, void* stream) {
  PETSYN_REQUIRE(z && g1 && sums, "null argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  const int rpp = 256 / (c / 8);
  const size_t smem = (size_t)rpp * 2 * c * sizeof(float);
  PETSYN_CHECK_CUDA(
      cudaFuncSetAttribute(norm_act_bwd_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  norm_act_bwd_reduce_kernel<<<row_blocks(rows, c), 256, smem, as_stream(stream)>>>(
      CBF(z), scale, shift, mean, rstd, CBF(g1), g1_cstride, g1_coff, act1, CBF(g2), g2_cstride, g2_coff, act2, slope,
      sums, rows, c);
  return check_launch("norm_act_bwd_reduce_kernel");
}

int32_t petsyn_norm_act_bwd_apply(const void* z, const float* scale, const float* shift, const float* mean,
                                  const float* rstd, const float* gamma, const void* g1, int32_t g1_cstride,
                                  int32_t g1_coff, int32_t act1, const void* g2, int32_t g2_cstride, int32_t g2_coff,
                                  int32_t act2, float slope, const double* sums, void* dz, float* dgamma,
                                  float* dbeta, int64_t rows, int32_t c, void* stream) {
  PETSYN_REQUIRE(z && g1 && dz, "null argument");
  PETSYN_REQUIRE(mean == nullptr || sums != nullptr, "normalised backward needs the reduction sums");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  norm_act_bwd_apply_kernel<<<row_blocks(rows, c), 256, 0, as_stream(stream)>>>(
      CBF(z), scale, shift, mean, rstd, gamma, CBF(g1), g1_cstride, g1_coff, act1, CBF(g2), g2_cstride, g2_coff, act2,
      slope, sums, BF(dz), dgamma, dbeta, rows, c);
  return check_launch("norm_act_bwd_apply_kernel");
}

int32_t petsyn_l1_loss_fwd_bwd(const float* y, const float* t, float* loss, float* dy, int64_t numel, float grad_scale,
                               void* stream) {
  PETSYN_REQUIRE(y && t && loss, "null argument");
  PETSYN_REQUIRE(numel > 0, "empty tensor");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((numel / 4 + 255) / 256, 148 * 8));
  DetWs ws;
  {
    int32_t rcw = det_workspace(&ws);
    if (rcw) return rcw;
  }
  PETSYN_CHECK_CUDA(launch_pdl(l1_loss_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), y, t, loss, dy, numel, 1.f / (float)numel,
                                                        grad_scale / (float)numel, ws));
  return check_launch("l1_loss_kernel");
}

int32_t petsyn_mse_const_fwd_bwd(const float* x, float target, float* loss, float* dx, int64_t numel, float grad_scale,
                                 void* stream) {
  PETSYN_REQUIRE(x && loss, "null argument");
  PETSYN_REQUIRE(numel > 0, "empty tensor");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((numel + 255) / 256, 148 * 8));
  DetWs ws;
  {
    int32_t rcw = det_workspace(&ws);
    if (rcw) return rcw;
  }
  PETSYN_CHECK_CUDA(launch_pdl(mse_const_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), x, target, loss, dx, numel, 1.f / (float)numel,
                                                          grad_scale / (float)numel, ws));
  return check_launch("mse_const_kernel");
}

int32_t petsyn_kl_fwd_bwd(const float* mu, const float* logvar, float* loss, float* dmu, float* dlogvar, int32_t n,
                          int32_t dim, int32_t pitch, float grad_scale, void* stream) {
  PETSYN_REQUIRE(mu && logvar && loss && n > 0 && dim > 0 && pitch >= dim, "bad argument");
  kl_kernel<<<1, 256, 0, as_stream(stream)>>>(mu, logvar, loss, dmu, dlogvar, n, dim, pitch, grad_scale);
  return check_launch("kl_kernel");
}

int32_t petsyn_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                         float beta2, float eps, int32_t step, const int32_t* step_dev, void* stream) {
  PETSYN_REQUIRE(p && g && m && v, "null argument");
  PETSYN_REQUIRE(step_dev != nullptr || step >= 1, "Adam step is 1-based");
  PETSYN_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                  reinterpret_cast<uintptr_t>(v)) % 16 == 0, "Adam arenas must be 16-byte aligned");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((numel / 4 + 255) / 256, 148 * 16));
  PETSYN_CHECK_CUDA(launch_pdl(adam_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, m, v, numel, lr, beta1, beta2, eps, step, step_dev));
  return check_launch("adam_kernel");
}

int32_t petsyn_sumsq(const float* g, float* out, int64_t numel, void* stream) {
  PETSYN_REQUIRE(g && out, "null argument");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((numel + 255) / 256, 148 * 8));
  DetWs ws;
  {
    int32_t rcw = det_workspace(&ws);
    if (rcw) return rcw;
  }
  sumsq_kernel<<<blocks, 256, 0, as_stream(stream)>>>(g, out, numel, ws);
  return check_launch("sumsq_kernel");
}

}  // extern "C"
