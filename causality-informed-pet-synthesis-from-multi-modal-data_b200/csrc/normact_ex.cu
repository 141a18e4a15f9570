// Generalised normalisation + activation kernels (descriptor API): batch / instance / group statistics, residual add
// and gradient fan-out.  Used by the BMGAN blocks (Conv -> InstanceNorm3d -> LeakyReLU, ResidualUnit sums, dense
// concatenation; bmgan_model.py:12-70) and the PatchGAN discriminator (BatchNorm3d).
//
// Same access pattern as elementwise.cu: activations are NDHWC bf16, a thread owns 8 consecutive channels (16 bytes),
// a warp covers consecutive channels of consecutive voxels; blockIdx.y is the sample (statistics group) index.
#include <cuda_bf16.h>

#include <algorithm>

#include "common.h"
#include "det_reduce.cuh"

namespace petsyn {
namespace nx {

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
__device__ __forceinline__ uint4 load_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ F8 unpack8(const uint4& raw) {
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
__device__ __forceinline__ F8 splat(float x) {
  F8 r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = x;
  return r;
}

// Per-thread software pipeline: every thread prefetches ITS OWN 16-byte pieces of the next kPf - 1 row iterations into a
// shared-memory ring with cp.async, so the bytes in flight live in shared memory instead of registers (a streaming
// kernel needs ~64 KB in flight per SM to saturate HBM3e).  A slot is only ever touched by its owner: no barriers.
constexpr int kPf = 4;
struct PfRing {
  uint32_t base;   // shared-memory address of this thread's slot 0
  int nt;
  __device__ PfRing(void* smem, int ntensors) : nt(ntensors) {
    base = (uint32_t)__cvta_generic_to_shared(smem) + threadIdx.x * 16;
  }
  __device__ __forceinline__ void issue(int stage, int t, const void* g) const {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(base + (stage * nt + t) * 4096), "l"(g) : "memory");
  }
  __device__ __forceinline__ uint4 get(int stage, int t) const {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(base + (stage * nt + t) * 4096));
    return v;
  }
  static __device__ __forceinline__ void commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
  static __device__ __forceinline__ void wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPf - 1) : "memory"); }
  static __host__ __device__ size_t bytes(int ntensors) { return (size_t)kPf * ntensors * 4096; }
};

// sigmoid(b) = 0.5 * tanh(0.5 * b) + 0.5 on the hardware tanh (one MUFU op instead of ex2 + rcp; |error| <= 2.5e-4, a
// sixteenth of the bf16 rounding step of the values these kernels store).  PETSYN_EXACT_SIGMOID at compile time: ex2 + rcp.
__device__ __forceinline__ float fast_sigmoid(float b) {
#ifdef PETSYN_EXACT_SIGMOID
  return __fdividef(1.f, 1.f + __expf(-b));
#else
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * b));
  return fmaf(0.5f, t, 0.5f);
#endif
}

__device__ __forceinline__ float act_fwd(float b, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return fmaxf(b, 0.f);
    case PETSYN_ACT_LRELU:
    case PETSYN_ACT_PRELU: return b > 0.f ? b : b * slope;
    case PETSYN_ACT_SILU: return b * fast_sigmoid(b);
    case PETSYN_ACT_TANH: return tanhf(b);
    default: return b;
  }
}
__device__ __forceinline__ float act_grad(float b, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return b > 0.f ? 1.f : 0.f;
    case PETSYN_ACT_LRELU:
    case PETSYN_ACT_PRELU: return b > 0.f ? 1.f : slope;
    case PETSYN_ACT_SILU: { const float s = fast_sigmoid(b); return s * (1.f + b * (1.f - s)); }
    case PETSYN_ACT_TANH: { const float t = tanhf(b); return 1.f - t * t; }
    default: return 1.f;
  }
}

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

// Per-channel sums over the rows this CTA streamed, added to DOUBLE-precision accumulators with one 64-bit atomic per CTA
// and value.  fp32 partials summed in a 53-bit accumulator are exact (so the order the CTAs finish in cannot change the
// result) as long as every partial is at least 2^-28 of the running sum; see det_reduce.cuh.  The consumers read the double
// sums directly (finalize / backward constants are computed in double anyway), so there is no ticket and no serial tail.
// out1[v * pitch1 + off1 + c] += sum, likewise out2 (statistics blocks of WIDER tensors use pitch / off); `extra` (optional)
// is one more scalar per CTA added the same way into *extra_out (the PReLU slope gradient).
template <int NV>
__device__ __forceinline__ void block_reduce_channels_to(const RowIter& it, float (&acc)[NV][8], float* smem, double* out1,
                                                         int pitch1, int off1, double* out2, int pitch2, int off2, int C,
                                                         float extra = 0.f, double* extra_out = nullptr) {
  __shared__ float s_extra[8];
  // Rows of one channel octet sit in lanes tx, tx + cpt, tx + 2 cpt, ... of every warp (cpt a power of two <= 32): a warp
  // butterfly folds them first, so the serial tail below adds 8 per-warp partials instead of `rpp` (up to 128) per-row ones
  const bool fold = (it.cpt & (it.cpt - 1)) == 0 && it.cpt <= 32;
  int nparts = it.rpp;
  if (fold) {
    for (int off = it.cpt; off < 32; off <<= 1) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[v][i] += __shfl_xor_sync(0xffffffffu, acc[v][i], off);
    }
    nparts = (int)(blockDim.x >> 5);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane < it.cpt) {
#pragma unroll
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < 8; ++i) smem[(w * NV + v) * C + lane * 8 + i] = acc[v][i];
    }
  } else if (it.active) {
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int i = 0; i < 8; ++i) smem[(it.ty * NV + v) * C + it.tx * 8 + i] = acc[v][i];
  }
  if (extra_out != nullptr) {
    for (int o = 16; o > 0; o >>= 1) extra += __shfl_xor_sync(0xffffffffu, extra, o);
    if ((threadIdx.x & 31) == 0) s_extra[threadIdx.x >> 5] = extra;
  }
  __syncthreads();
  const int nc = NV * C;
  for (int e = threadIdx.x; e < nc; e += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < nparts; ++r) s += smem[r * nc + e];
    const int v = e / C, c = e - v * C;
    if (out1) atomicAdd(out1 + v * pitch1 + off1 + c, (double)s);
    if (out2) atomicAdd(out2 + v * pitch2 + off2 + c, (double)s);
  }
  if (extra_out != nullptr && threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += s_extra[w];
    if (s != 0.f) atomicAdd(extra_out, (double)s);
  }
}

template <int NV>
__device__ __forceinline__ void block_reduce_channels(const RowIter& it, float (&acc)[NV][8], float* smem, double* out,
                                                      int C, float extra = 0.f, double* extra_out = nullptr) {
  block_reduce_channels_to<NV>(it, acc, smem, out, C, 0, nullptr, 0, 0, C, extra, extra_out);
}

// sums[sample][0:C] += sum z, sums[sample][C:2C] += sum z^2
__global__ void __launch_bounds__(256) stats_kernel(const __nv_bfloat16* __restrict__ z, double* __restrict__ sums,
                                                    int64_t rows, int C, int zcs, int zco) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* smem_f = reinterpret_cast<float*>(smem_raw + PfRing::bytes(1));
  RowIter it(C);
  z += (int64_t)blockIdx.y * rows * zcs + zco;
  sums += (int64_t)blockIdx.y * 2 * C;
  float acc[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (it.active) {
    PfRing pf(smem_raw, 1);
    const int64_t stride = (int64_t)gridDim.x * it.rpp;
    const int64_t r0 = (int64_t)blockIdx.x * it.rpp + it.ty;
    const __nv_bfloat16* src = z + it.tx * 8;
    for (int k = 0; k < kPf - 1; ++k) {
      if (r0 + k * stride < rows) pf.issue(k, 0, src + (rows - 1 - (r0 + k * stride)) * zcs);   // descending, see below
      PfRing::commit();
    }
    int k = 0;
    for (int64_t r = r0; r < rows; r += stride, ++k) {
      const int64_t rn = r + (kPf - 1) * stride;
      if (rn < rows) pf.issue((k + kPf - 1) % kPf, 0, src + (rows - 1 - rn) * zcs);
      PfRing::commit();
      PfRing::wait();
      const F8 x0 = unpack8(pf.get(k % kPf, 0));
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[0][i] += x0.v[i];
        acc[1][i] += x0.v[i] * x0.v[i];
      }
    }
  }
  block_reduce_channels<2>(it, acc, smem_f, sums, C);
}

// One thread per (sample, channel).  group_size channels share their statistics (GroupNorm); group_size == 1 is
// Batch (nsamples == 1) / Instance (nsamples == N) normalisation.
__global__ void finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ running_mean,
                                float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                float* __restrict__ save_mean, float* __restrict__ save_rstd, int64_t rows, int C,
                                int nsamples, int group_size, float eps, float momentum, int training) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nsamples * C) return;
  const int s = i / C, c = i % C;
  double mean, var;
  if (training || running_mean == nullptr) {
    const double* sm = sums + (int64_t)s * 2 * C;
    const int g0 = c / group_size * group_size;
    double a = 0, b = 0;
    for (int j = 0; j < group_size; ++j) { a += sm[g0 + j]; b += sm[C + g0 + j]; }
    const double cnt = (double)rows * group_size;
    mean = a / cnt;
    var = b / cnt - mean * mean;
    if (var < 0) var = 0;
    if (running_mean != nullptr && nsamples == 1 && group_size == 1) {
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
  scale[i] = (float)(g * rstd);
  shift[i] = (float)(b - mean * g * rstd);
  if (save_mean) save_mean[i] = (float)mean;
  if (save_rstd) save_rstd[i] = (float)rstd;
}

struct Dev {   // device copy of petsyn_normact_desc with typed pointers
  const __nv_bfloat16* z;
  int64_t rows;
  int C, per_sample;
  const float *scale, *shift, *mean, *rstd, *gamma;
  __nv_bfloat16* t1; int cs1, co1, act1;
  __nv_bfloat16* t2; int cs2, co2, act2;
  float slope;
  __nv_bfloat16* res; int csr, cor, res_acc;
  double* sums;
  __nv_bfloat16* dz;
  float *dgamma, *dbeta;
  const float* slope_dev;
  double* dslope;
  const float *ka, *kb;   // per (sample, channel) backward constants of the affine/group path (nullptr: plain path)
  int dz_acc;
  int cs_off, cs_n;       // channel sub-range of dz that dz_colsum covers
  double* dz_colsum;      // optional [C] += column sums of dz
  double* st1; int st1_c, st1_off;   // optional statistics of destination 1 for the norm that consumes it
  double* st2; int st2_c, st2_off;
  // fused finalize (forward): statistics sums -> scale / shift / mean / rstd inside the apply kernel
  const double* fin_sums;
  const float *fin_gamma, *fin_beta;
  int fin_gs;
  float fin_eps;
  // fused group combine (backward): ka / kb / dgamma / dbeta inside the apply kernel
  int gc_gs, gc_ns, gc_acc;   // gc_gs > 0: enabled
  int zcs, zco;               // channel pitch / offset of z (z may be a channel slice of a wider buffer)
  int dzcs, dzco;             // ... and of dz
  const __nv_bfloat16* ex;    // bwd, optional: extra addend of dz (gradient arriving through an identity skip connection)
  int excs, exco;
};

// ACT < 0: activation codes read from the descriptor at run time (rare combinations); ACT >= 0: act1 == act2 == ACT
// known at compile time (the common case), so the inner loop carries no switch.
template <int ACT>
__global__ void __launch_bounds__(256, 3) fwd_kernel(const Dev d) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RowIter it(d.C);
  const bool has_res = d.res != nullptr;
  const bool want_stats = d.st1 != nullptr || d.st2 != nullptr;      // uniform over the grid
  const int s = blockIdx.y;
  float st[2][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) st[0][i] = st[1][i] = 0.f;
  const int64_t base = (int64_t)s * d.rows;
  const int64_t stride = (int64_t)gridDim.x * it.rpp;
  PfRing pf(smem_raw, has_res ? 2 : 1);
  // Row order alternates between consecutive passes over the same tensor so that each pass starts with what the
  // previous one touched last (still in the 126 MB L2): producers (convs) write ascending, statistics / backward-reduce
  // read DESCENDING, the apply passes read ascending again
  const int64_t q0 = (int64_t)blockIdx.x * it.rpp + it.ty;
  const __nv_bfloat16* zsrc = d.z + d.zco + it.tx * 8;
  const __nv_bfloat16* rsrc = has_res ? d.res + d.cor + it.tx * 8 : nullptr;
  auto issue = [&](int k) {
    const int64_t q = q0 + k * stride;
    if (q < d.rows) {
      const int64_t r = base + q;
      pf.issue(k % kPf, 0, zsrc + r * d.zcs);
      if (has_res) pf.issue(k % kPf, 1, rsrc + r * d.csr);
    }
    PfRing::commit();
  };
  // the first rows are in flight before the finalize prologue below, so its latency hides behind the HBM round trip
  if (it.active)
    for (int k = 0; k < kPf - 1; ++k) issue(k);
  // Fused finalize (Instance / Group normalisation): every CTA derives its sample's scale / shift from the statistics sums
  // (the arithmetic of finalize_kernel, one thread per channel) into shared memory; the first CTA of each sample also
  // publishes scale / shift / mean / rstd for the backward pass.  Saves a 4 us launch in front of every apply pass.
  float* fin = reinterpret_cast<float*>(smem_raw + PfRing::bytes(has_res ? 2 : 1));
  if (d.fin_sums != nullptr) {
    const double* sm = d.fin_sums + (int64_t)s * 2 * d.C;
    const double cnt = (double)d.rows * d.fin_gs;
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
      const int g0 = c / d.fin_gs * d.fin_gs;
      double a = 0, b = 0;
      for (int j = 0; j < d.fin_gs; ++j) { a += sm[g0 + j]; b += sm[d.C + g0 + j]; }
      const double mean = a / cnt;
      double var = b / cnt - mean * mean;
      if (var < 0) var = 0;
      const double rstd = 1.0 / sqrt(var + (double)d.fin_eps);
      const float g = d.fin_gamma ? d.fin_gamma[c] : 1.f, be = d.fin_beta ? d.fin_beta[c] : 0.f;
      const float scl = (float)(g * rstd), shf = (float)(be - mean * g * rstd);
      fin[c] = scl;
      fin[d.C + c] = shf;
      if (blockIdx.x == 0) {
        const int i = s * d.C + c;
        const_cast<float*>(d.scale)[i] = scl;
        const_cast<float*>(d.shift)[i] = shf;
        const_cast<float*>(d.mean)[i] = (float)mean;
        const_cast<float*>(d.rstd)[i] = (float)rstd;
      }
    }
    __syncthreads();
  }
  if (it.active) {
    const int so = d.per_sample ? s * d.C : 0;
    F8 sc = splat(1.f), sh = splat(0.f);
    if (d.fin_sums != nullptr) { sc = load8f(fin + it.tx * 8); sh = load8f(fin + d.C + it.tx * 8); }
    else if (d.scale) { sc = load8f(d.scale + so + it.tx * 8); sh = load8f(d.shift + so + it.tx * 8); }
    const float slope = d.slope_dev ? __ldg(d.slope_dev) : d.slope;
    const int a1 = ACT >= 0 ? ACT : d.act1, a2 = ACT >= 0 ? ACT : d.act2;
    int k = 0;
    for (int64_t q = q0; q < d.rows; q += stride, ++k) {
      issue(k + kPf - 1);
      PfRing::wait();
      const int64_t r = base + q;
      const F8 x = unpack8(pf.get(k % kPf, 0));
      F8 o1, o2;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float b = x.v[i] * sc.v[i] + sh.v[i];
        o1.v[i] = act_fwd(b, a1, slope);
        o2.v[i] = (ACT >= 0) ? o1.v[i] : act_fwd(b, a2, slope);
      }
      if (has_res) {
        const F8 rs = unpack8(pf.get(k % kPf, 1));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o1.v[i] += rs.v[i];
          o2.v[i] += rs.v[i];
        }
      }
      store8(d.t1 + r * d.cs1 + d.co1 + it.tx * 8, o1);
      if (d.t2) store8(d.t2 + r * d.cs2 + d.co2 + it.tx * 8, o2);
      if (want_stats) {
        // statistics of what was STORED (bf16-rounded), so the consumer sees exactly what a separate pass would compute
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float v = __bfloat162float(__float2bfloat16(o1.v[i]));
          st[0][i] += v;
          st[1][i] += v * v;
        }
      }
    }
  }
  if (want_stats) {
    float* smem_f = fin + (d.fin_sums != nullptr ? 2 * d.C : 0);
    block_reduce_channels_to<2>(it, st, smem_f, d.st1 ? d.st1 + (int64_t)s * 2 * d.st1_c : nullptr, d.st1_c, d.st1_off,
                                d.st2 ? d.st2 + (int64_t)s * 2 * d.st2_c : nullptr, d.st2_c, d.st2_off, d.C);
  }
}

template <int ACT>
__global__ void __launch_bounds__(256, 2) bwd_reduce_kernel(const Dev d) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const bool has_t2 = d.t2 != nullptr;
  const int nt = has_t2 ? 3 : 2;
  float* smem_f = reinterpret_cast<float*>(smem_raw + PfRing::bytes(nt));
  RowIter it(d.C);
  const int s = blockIdx.y;
  const int64_t base = (int64_t)s * d.rows;
  const int so = d.per_sample ? s * d.C : 0;
  float acc[2][8];   // sum g, sum g * x  (turned into sum g * zhat after the loop)
  float dsl = 0.f;
  const float slope = d.slope_dev ? __ldg(d.slope_dev) : d.slope;
  const int a1 = ACT >= 0 ? ACT : d.act1, a2 = ACT >= 0 ? ACT : d.act2;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = acc[1][i] = 0.f;
  if (it.active) {
    F8 sc = splat(1.f), sh = splat(0.f);
    if (d.scale) { sc = load8f(d.scale + so + it.tx * 8); sh = load8f(d.shift + so + it.tx * 8); }
    const int64_t stride = (int64_t)gridDim.x * it.rpp;
    PfRing pf(smem_raw, nt);
    const int64_t r0 = (int64_t)blockIdx.x * it.rpp + it.ty;
    const __nv_bfloat16* zsrc = d.z + base * d.zcs + d.zco + it.tx * 8;
    const __nv_bfloat16* asrc = d.t1 + base * d.cs1 + d.co1 + it.tx * 8;
    const __nv_bfloat16* bsrc = has_t2 ? d.t2 + base * d.cs2 + d.co2 + it.tx * 8 : nullptr;
    auto issue = [&](int k) {
      const int64_t q = r0 + k * stride;
      if (q < d.rows) {
        const int64_t r = d.rows - 1 - q;               // descending (see stats_kernel)
        pf.issue(k % kPf, 0, zsrc + r * d.zcs);
        pf.issue(k % kPf, 1, asrc + r * d.cs1);
        if (has_t2) pf.issue(k % kPf, 2, bsrc + r * d.cs2);
      }
      PfRing::commit();
    };
    for (int k = 0; k < kPf - 1; ++k) issue(k);
    int k = 0;
    for (int64_t r = r0; r < d.rows; r += stride, ++k) {
      issue(k + kPf - 1);
      PfRing::wait();
      const F8 x = unpack8(pf.get(k % kPf, 0));
      F8 av = unpack8(pf.get(k % kPf, 1));
      F8 bv = splat(0.f);
      if (has_t2) {
        bv = unpack8(pf.get(k % kPf, 2));
        if (ACT >= 0) {                    // one activation for both destinations: their gradients simply add
#pragma unroll
          for (int i = 0; i < 8; ++i) av.v[i] += bv.v[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float b = x.v[i] * sc.v[i] + sh.v[i];
        float g;
        if (ACT >= 0) {
          g = av.v[i] * act_grad(b, a1, slope);
        } else {
          g = av.v[i] * act_grad(b, a1, slope);
          if (has_t2) g += bv.v[i] * act_grad(b, a2, slope);
        }
        acc[0][i] += g;
        acc[1][i] += g * x.v[i];
        if (d.dslope != nullptr && b < 0.f)                                     // d PReLU / d slope = min(b, 0)
          dsl += (ACT >= 0 ? av.v[i] : av.v[i] + bv.v[i]) * b;
      }
    }
    const F8 mu = load8f(d.mean + so + it.tx * 8), rs = load8f(d.rstd + so + it.tx * 8);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[1][i] = (acc[1][i] - mu.v[i] * acc[0][i]) * rs.v[i];   // sum g * (x - mu) * rstd
  }
  block_reduce_channels<2>(it, acc, smem_f, d.sums + (d.per_sample ? (int64_t)s * 2 * d.C : 0), d.C, dsl, d.dslope);
}

template <int ACT>
__global__ void __launch_bounds__(256, 2) bwd_apply_kernel(const Dev d) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int nt_ring = (d.t2 != nullptr ? 3 : 2) + (d.ex != nullptr ? 1 : 0);
  float* smem_f = reinterpret_cast<float*>(smem_raw + PfRing::bytes(nt_ring));
  RowIter it(d.C);
  if (blockIdx.x == 0 && blockIdx.y == 0 && d.dgamma != nullptr && d.ka == nullptr && d.gc_gs == 0) {
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
      d.dbeta[c] = (float)d.sums[c];
      d.dgamma[c] = (float)d.sums[d.C + c];
    }
  }
  float csum[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) csum[0][i] = 0.f;
  const int64_t base = (int64_t)blockIdx.y * d.rows;
  const int64_t stride = (int64_t)gridDim.x * it.rpp;
  const bool has_t2 = d.t2 != nullptr;
  const bool has_ex = d.ex != nullptr;
  PfRing pf(smem_raw, nt_ring);
  const int64_t q0 = (int64_t)blockIdx.x * it.rpp + it.ty;              // ascending row order (see fwd_kernel)
  const __nv_bfloat16* zsrc = d.z + d.zco + it.tx * 8;
  const __nv_bfloat16* asrc = d.t1 + d.co1 + it.tx * 8;
  const __nv_bfloat16* bsrc = has_t2 ? d.t2 + d.co2 + it.tx * 8 : nullptr;
  const __nv_bfloat16* esrc = has_ex ? d.ex + d.exco + it.tx * 8 : nullptr;
  auto issue = [&](int k) {
    const int64_t q = q0 + k * stride;
    if (q < d.rows) {
      const int64_t r = base + q;
      pf.issue(k % kPf, 0, zsrc + r * d.zcs);
      pf.issue(k % kPf, 1, asrc + r * d.cs1);
      if (has_t2) pf.issue(k % kPf, 2, bsrc + r * d.cs2);
      if (has_ex) pf.issue(k % kPf, nt_ring - 1, esrc + r * d.excs);
    }
    PfRing::commit();
  };
  // first rows in flight before the constant prologue below (its latency hides behind the HBM round trip)
  if (it.active)
    for (int k = 0; k < kPf - 1; ++k) issue(k);
  // Fused group combine (GroupNorm / per-sample affine): with S0 = sum g, S1 = sum g*zhat per (sample, channel) from the reduce pass,
  //   ka[c] = rstd[c] * sum_{c' in group(c)} gamma[c'] S0[c'] / (rows * group_size),  kb likewise with S1  (this CTA's sample);
  //   dgamma[c] = sum_s S1[s,c], dbeta[c] = sum_s S0[s,c]                                      (CTA (0, 0) only)
  float* comb = smem_f + (d.dz_colsum != nullptr ? it.rpp * d.C : 0);
  if (d.gc_gs > 0) {
    const int s = blockIdx.y;
    const double* sm = d.sums + (int64_t)s * 2 * d.C;
    const float inv = 1.f / ((float)d.rows * (float)d.gc_gs);
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
      const int g0 = c / d.gc_gs * d.gc_gs;
      float a = 0.f, b = 0.f;
      for (int j = 0; j < d.gc_gs; ++j) {
        const float ga = d.gamma ? d.gamma[g0 + j] : 1.f;
        a += ga * (float)sm[g0 + j];
        b += ga * (float)sm[d.C + g0 + j];
      }
      const float r = d.rstd[s * d.C + c];
      comb[c] = r * a * inv;
      comb[d.C + c] = r * b * inv;
    }
    if (blockIdx.x == 0 && blockIdx.y == 0 && d.dgamma != nullptr) {
      for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int q = 0; q < d.gc_ns; ++q) {
          a += (float)d.sums[(int64_t)q * 2 * d.C + d.C + c];
          b += (float)d.sums[(int64_t)q * 2 * d.C + c];
        }
        if (d.gc_acc) { d.dgamma[c] += a; d.dbeta[c] += b; } else { d.dgamma[c] = a; d.dbeta[c] = b; }
      }
    }
    __syncthreads();
  }
  if (it.active) {
  const int s = blockIdx.y;
  const int so = d.per_sample ? s * d.C : 0;
  // dz = k0*g - k1 - zhat*k2 with zhat = (x - mu)*rstd   ==   k0*g - kA - x*kB,  kA = k1 - mu*rstd*k2, kB = rstd*k2
  F8 sc = splat(1.f), sh = splat(0.f), k0 = splat(1.f), kA = splat(0.f), kB = splat(0.f);
  if (d.scale) { sc = load8f(d.scale + so + it.tx * 8); sh = load8f(d.shift + so + it.tx * 8); }
  if (d.mean) {
    const F8 mu = load8f(d.mean + so + it.tx * 8), rs = load8f(d.rstd + so + it.tx * 8);
    const F8 ga = d.gamma ? load8f(d.gamma + it.tx * 8) : splat(1.f);
    F8 k1, k2;
    if (d.gc_gs > 0) {                         // group statistics and/or per-sample affine: constants from the prologue
      k1 = load8f(comb + it.tx * 8);
      k2 = load8f(comb + d.C + it.tx * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) k0.v[i] = ga.v[i] * rs.v[i];
    } else if (d.ka != nullptr) {              // ... or precomputed by group_combine_kernel
      k1 = load8f(d.ka + so + it.tx * 8);
      k2 = load8f(d.kb + so + it.tx * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) k0.v[i] = ga.v[i] * rs.v[i];
    } else {
      const double* sm = d.sums + (d.per_sample ? (int64_t)s * 2 * d.C : 0);
      F8 s0, s1;
#pragma unroll
      for (int i = 0; i < 8; ++i) { s0.v[i] = (float)sm[it.tx * 8 + i]; s1.v[i] = (float)sm[d.C + it.tx * 8 + i]; }
      const float inv = 1.f / (float)d.rows;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        k0.v[i] = ga.v[i] * rs.v[i];
        k1.v[i] = k0.v[i] * s0.v[i] * inv;
        k2.v[i] = k0.v[i] * s1.v[i] * inv;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      kB.v[i] = rs.v[i] * k2.v[i];
      kA.v[i] = k1.v[i] - mu.v[i] * kB.v[i];
    }
  }
  const float slope = d.slope_dev ? __ldg(d.slope_dev) : d.slope;
  const int a1 = ACT >= 0 ? ACT : d.act1, a2 = ACT >= 0 ? ACT : d.act2;
  int k = 0;
  for (int64_t q = q0; q < d.rows; q += stride, ++k) {
    issue(k + kPf - 1);
    PfRing::wait();
    const int64_t r = base + q;
    const F8 x = unpack8(pf.get(k % kPf, 0));
    F8 av = unpack8(pf.get(k % kPf, 1));
    F8 bv = splat(0.f);
    if (has_t2) {
      bv = unpack8(pf.get(k % kPf, 2));
      if (ACT >= 0) {                      // one activation for both destinations: their gradients simply add
#pragma unroll
        for (int i = 0; i < 8; ++i) av.v[i] += bv.v[i];
      }
    }
    F8 o, dr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float b = x.v[i] * sc.v[i] + sh.v[i];
      float g;
      if (ACT >= 0) {
        g = av.v[i] * act_grad(b, a1, slope);
        dr.v[i] = av.v[i];
      } else {
        g = av.v[i] * act_grad(b, a1, slope);
        if (has_t2) g += bv.v[i] * act_grad(b, a2, slope);
        dr.v[i] = av.v[i] + bv.v[i];
      }
      o.v[i] = k0.v[i] * g - kA.v[i] - x.v[i] * kB.v[i];
    }
    if (has_ex) {                       // d(out)/d(x) = identity of a skip connection: its gradient joins dz here
      const F8 e = unpack8(pf.get(k % kPf, nt_ring - 1));
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += e.v[i];
    }
    __nv_bfloat16* dzp = d.dz + r * d.dzcs + d.dzco + it.tx * 8;
    if (d.dz_acc) {
      const F8 old = load8(dzp);
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] += old.v[i];
    }
    if (d.dz_colsum != nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) csum[0][i] += o.v[i];   // column sums of the FINAL gradient (after extra / accumulation)
    }
    store8(dzp, o);
    if (d.res) {
      __nv_bfloat16* p = d.res + r * d.csr + d.cor + it.tx * 8;
      if (d.res_acc) {
        const F8 old = load8(p);
#pragma unroll
        for (int i = 0; i < 8; ++i) dr.v[i] += old.v[i];
      }
      store8(p, dr);
    }
  }
  }   // it.active
  // the bias gradient sums over the samples as well: one reduction over the whole grid
  if (d.dz_colsum != nullptr) {
    // column sums of the final dz over the rows of this CTA -> double accumulator, channels [cs_off, cs_off + cs_n) only
    // (warp butterfly over the rows of a channel octet first, as in block_reduce_channels_to)
    const bool fold = (it.cpt & (it.cpt - 1)) == 0 && it.cpt <= 32;
    int nparts = it.rpp;
    if (fold) {
      for (int off = it.cpt; off < 32; off <<= 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) csum[0][i] += __shfl_xor_sync(0xffffffffu, csum[0][i], off);
      }
      nparts = (int)(blockDim.x >> 5);
      const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
      if (lane < it.cpt) {
#pragma unroll
        for (int i = 0; i < 8; ++i) smem_f[w * d.C + lane * 8 + i] = csum[0][i];
      }
    } else if (it.active) {
#pragma unroll
      for (int i = 0; i < 8; ++i) smem_f[it.ty * d.C + it.tx * 8 + i] = csum[0][i];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < d.cs_n; e += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < nparts; ++r) s += smem_f[r * d.C + d.cs_off + e];
      atomicAdd(d.dz_colsum + e, (double)s);
    }
  }
}

// GroupNorm / per-sample affine backward constants.  With S0 = sum g, S1 = sum g*zhat per (sample, channel):
//   ka[s,c] = rstd[s,c] * sum_{c' in group(c)} gamma[c'] S0[s,c'] / (rows * group_size),  kb likewise with S1;
//   dgamma[c] = sum_s S1[s,c], dbeta[c] = sum_s S0[s,c].
__global__ void group_combine_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                     const float* __restrict__ rstd, float* __restrict__ ka, float* __restrict__ kb,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C,
                                     int nsamples, int group_size, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nsamples * C) {
    const int s = i / C, c = i % C;
    const double* sm = sums + (int64_t)s * 2 * C;
    const int g0 = c / group_size * group_size;
    float a = 0.f, b = 0.f;
    for (int j = 0; j < group_size; ++j) {
      const float ga = gamma ? gamma[g0 + j] : 1.f;
      a += ga * (float)sm[g0 + j];
      b += ga * (float)sm[C + g0 + j];
    }
    const float inv = 1.f / ((float)rows * (float)group_size);
    ka[i] = rstd[i] * a * inv;
    kb[i] = rstd[i] * b * inv;
  }
  if (i < C && dgamma != nullptr) {
    float a = 0.f, b = 0.f;
    for (int s = 0; s < nsamples; ++s) {
      a += (float)sums[(int64_t)s * 2 * C + C + i];
      b += (float)sums[(int64_t)s * 2 * C + i];
    }
    if (accumulate) { dgamma[i] += a; dbeta[i] += b; } else { dgamma[i] = a; dbeta[i] = b; }
  }
}

// dst[r, coff:coff+C] (+)= src[r, soff:soff+C]   (gradient fan-in for tensors with several consumers)
__global__ void __launch_bounds__(256) add_slice_kernel(const __nv_bfloat16* __restrict__ src, int css, int cos,
                                                        __nv_bfloat16* __restrict__ dst, int csd, int cod, int64_t rows,
                                                        int C, int accumulate) {
  RowIter it(C);
  if (!it.active) return;
  const int64_t stride = (int64_t)gridDim.x * it.rpp;
  for (int64_t r = (int64_t)blockIdx.x * it.rpp + it.ty; r < rows; r += stride) {
    F8 x = load8(src + r * css + cos + it.tx * 8);
    __nv_bfloat16* p = dst + r * csd + cod + it.tx * 8;
    if (accumulate) {
      const F8 old = load8(p);
#pragma unroll
      for (int i = 0; i < 8; ++i) x.v[i] += old.v[i];
    }
    store8(p, x);
  }
}

// out[c] = sum_r x[r, coff + c]  (bias gradient), fp32, caller-zeroed
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int cs, int co,
                                                     double* __restrict__ out, int64_t rows, int C) {
  pdl_sync();
  extern __shared__ float smem_f[];
  RowIter it(C);
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  if (it.active) {
    const int64_t stride = (int64_t)gridDim.x * it.rpp;
    for (int64_t r = (int64_t)blockIdx.x * it.rpp + it.ty; r < rows; r += stride) {
      const F8 v = load8(x + r * cs + co + it.tx * 8);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[0][i] += v.v[i];
    }
  }
  block_reduce_channels<1>(it, acc, smem_f, out, C);
}

// grid.x per sample: `waves` = CTAs per SM over all samples.  Two persistent CTAs per SM with equal shares of the rows: one
// wave (no partial last wave, one ramp-up), and room left on every SM for the weight-gradient kernels of the side stream.
// Measured on configs[1] (whole step): 8 CTAs per SM 15.50 ms, 6: 15.32, 4: 15.03, 3: 15.15, 2: 14.74, 1: 16.19
static int row_blocks(int64_t rows, int C, int nsamples, int waves = 2) {
  const int rpp = 256 / (C / 8);
  const int64_t want = (rows + rpp - 1) / rpp;
  if (const char* e = getenv("PETSYN_NORM_WAVES")) waves = std::max(1, atoi(e));
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, std::max(1, 148 * waves / nsamples)));
}

static int32_t to_dev(const petsyn_normact_desc* d, Dev* o) {
  PETSYN_REQUIRE(d != nullptr, "null descriptor");
  PETSYN_REQUIRE(d->z != nullptr && d->rows > 0 && d->nsamples >= 1 && d->nsamples <= kDetRows, "bad tensor");
  PETSYN_REQUIRE(d->c % 8 == 0 && d->c >= 8 && d->c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  PETSYN_REQUIRE((d->t1_cstride | d->t1_coff | d->t2_cstride | d->t2_coff | d->res_cstride | d->res_coff) % 8 == 0,
                 "channel pitches/offsets must be multiples of 8");
  o->z = reinterpret_cast<const __nv_bfloat16*>(d->z);
  o->rows = d->rows; o->C = d->c; o->per_sample = d->per_sample_stats;
  o->scale = d->scale; o->shift = d->shift; o->mean = d->mean; o->rstd = d->rstd; o->gamma = d->gamma;
  o->t1 = reinterpret_cast<__nv_bfloat16*>(d->t1); o->cs1 = d->t1_cstride; o->co1 = d->t1_coff; o->act1 = d->act1;
  o->t2 = reinterpret_cast<__nv_bfloat16*>(d->t2); o->cs2 = d->t2_cstride; o->co2 = d->t2_coff; o->act2 = d->act2;
  o->slope = d->slope;
  o->res = reinterpret_cast<__nv_bfloat16*>(d->res); o->csr = d->res_cstride; o->cor = d->res_coff;
  o->res_acc = d->res_accumulate;
  o->sums = reinterpret_cast<double*>(d->sums);
  o->dz = reinterpret_cast<__nv_bfloat16*>(d->dz);
  o->dgamma = d->dgamma; o->dbeta = d->dbeta;
  o->slope_dev = d->slope_dev; o->dslope = reinterpret_cast<double*>(d->dslope);
  o->ka = o->kb = nullptr;
  o->dz_acc = d->dz_accumulate;
  o->dz_colsum = reinterpret_cast<double*>(d->dz_colsum);
  o->cs_off = d->dz_colsum_c > 0 ? d->dz_colsum_coff : 0;
  o->cs_n = d->dz_colsum_c > 0 ? d->dz_colsum_c : d->c;
  PETSYN_REQUIRE(o->cs_off >= 0 && o->cs_off + o->cs_n <= d->c, "dz_colsum channel range outside dz");
  o->st1 = reinterpret_cast<double*>(d->t1_stats); o->st1_c = d->t1_stats_c; o->st1_off = d->t1_stats_coff;
  o->st2 = reinterpret_cast<double*>(d->t2_stats); o->st2_c = d->t2_stats_c; o->st2_off = d->t2_stats_coff;
  o->fin_sums = reinterpret_cast<const double*>(d->fin_sums); o->fin_gamma = d->fin_gamma; o->fin_beta = d->fin_beta;
  o->fin_gs = d->fin_group_size > 1 ? d->fin_group_size : 1; o->fin_eps = d->fin_eps;
  o->gc_gs = o->gc_ns = o->gc_acc = 0;
  o->zcs = d->z_cstride > 0 ? d->z_cstride : d->c; o->zco = d->z_coff;
  o->dzcs = d->dz_cstride > 0 ? d->dz_cstride : d->c; o->dzco = d->dz_coff;
  o->ex = reinterpret_cast<const __nv_bfloat16*>(d->extra); o->excs = d->extra_cstride; o->exco = d->extra_coff;
  PETSYN_REQUIRE((d->z_cstride | d->z_coff | d->dz_cstride | d->dz_coff | d->extra_cstride | d->extra_coff) % 8 == 0,
                 "channel pitches/offsets must be multiples of 8");
  PETSYN_REQUIRE(d->extra == nullptr || d->extra_cstride >= d->c, "extra addend needs its channel pitch");
  PETSYN_REQUIRE(d->fin_sums == nullptr || (d->scale && d->shift && d->mean && d->rstd && d->per_sample_stats &&
                                            d->c % o->fin_gs == 0),
                 "fused finalize needs per-sample statistics and the scale / shift / mean / rstd outputs");
  return PETSYN_OK;
}

}  // namespace nx
}  // namespace petsyn

using namespace petsyn;
using namespace petsyn::nx;

// compile-time activation when both destinations share it (always true for the module mirrors), run-time otherwise
#define PETSYN_NX_CASE(KERNEL, A, GRID, SMEM, ST, D)                                                         \
  {                                                                                                          \
    static bool attr_ = false;                                                                               \
    if (!attr_) {                                                                                            \
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024)); \
      attr_ = true;                                                                                          \
    }                                                                                                        \
    PETSYN_CHECK_CUDA(launch_pdl(KERNEL<A>, GRID, dim3(256), SMEM, ST, D));                                  \
  }                                                                                                          \
  break;
#define PETSYN_NX_DISPATCH(KERNEL, GRID, SMEM, ST, D)                                         \
  do {                                                                                        \
    const int _a = ((D).t2 == nullptr || (D).act1 == (D).act2) ? (D).act1 : -1;               \
    switch (_a) {                                                                             \
      case PETSYN_ACT_NONE: PETSYN_NX_CASE(KERNEL, PETSYN_ACT_NONE, GRID, SMEM, ST, D)        \
      case PETSYN_ACT_RELU: PETSYN_NX_CASE(KERNEL, PETSYN_ACT_RELU, GRID, SMEM, ST, D)        \
      case PETSYN_ACT_LRELU: PETSYN_NX_CASE(KERNEL, PETSYN_ACT_LRELU, GRID, SMEM, ST, D)      \
      case PETSYN_ACT_SILU: PETSYN_NX_CASE(KERNEL, PETSYN_ACT_SILU, GRID, SMEM, ST, D)        \
      default: PETSYN_NX_CASE(KERNEL, -1, GRID, SMEM, ST, D)                                  \
    }                                                                                         \
  } while (0)

extern "C" {

int32_t petsyn_norm_stats(const void* z, double* sums, int64_t rows, int32_t c, int32_t nsamples, void* stream) {
  return petsyn_norm_stats_slice(z, c, 0, sums, rows, c, nsamples, stream);
}

int32_t petsyn_norm_stats_slice(const void* z, int32_t cstride, int32_t coff, double* sums, int64_t rows, int32_t c,
                                int32_t nsamples, void* stream) {
  PETSYN_REQUIRE(z && sums && rows > 0 && nsamples >= 1 && nsamples <= kDetRows, "bad argument");
  PETSYN_REQUIRE(cstride >= c && cstride % 8 == 0 && coff % 8 == 0 && coff + c <= cstride, "bad channel slice");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048, "channels must be a multiple of 8 in [8, 2048]");
  const int rpp = 256 / (c / 8);
  const size_t smem = PfRing::bytes(1) + (size_t)rpp * 2 * c * sizeof(float);
  dim3 grid((unsigned)row_blocks(rows, c, nsamples), (unsigned)nsamples);
  PETSYN_CHECK_CUDA(launch_pdl(stats_kernel, grid, dim3(256), smem, as_stream(stream),
                               reinterpret_cast<const __nv_bfloat16*>(z), sums, rows, c, cstride, coff));
  return check_launch("stats_kernel");
}

int32_t petsyn_norm_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float* scale, float* shift, float* save_mean, float* save_rstd,
                             int64_t rows, int32_t c, int32_t nsamples, int32_t group_size, float eps, float momentum,
                             int32_t training, void* stream) {
  PETSYN_REQUIRE(scale && shift && nsamples >= 1 && group_size >= 1 && c % group_size == 0, "bad argument");
  PETSYN_REQUIRE(training || running_mean != nullptr || sums != nullptr, "missing statistics");
  const int total = nsamples * c;
  finalize_kernel<<<(total + 127) / 128, 128, 0, as_stream(stream)>>>(sums, gamma, beta, running_mean, running_var, scale,
                                                                      shift, save_mean, save_rstd, rows, c, nsamples,
                                                                      group_size, eps, momentum, training);
  return check_launch("finalize_kernel");
}

int32_t petsyn_normact_fwd(const petsyn_normact_desc* desc, void* stream) {
  Dev d;
  int32_t rc = to_dev(desc, &d);
  if (rc) return rc;
  PETSYN_REQUIRE(d.t1 != nullptr, "missing destination");
  dim3 grid((unsigned)row_blocks(d.rows, d.C, desc->nsamples), (unsigned)desc->nsamples);
  PETSYN_REQUIRE(!(d.st1 || d.st2) || desc->act1 == desc->act2 || d.t2 == nullptr,
                 "destination statistics need one activation for both destinations");
  const size_t fwd_smem = PfRing::bytes(d.res ? 2 : 1) + (d.fin_sums ? (size_t)2 * d.C * sizeof(float) : 0) +
                          ((d.st1 || d.st2) ? (size_t)(256 / (d.C / 8)) * 2 * d.C * sizeof(float) : 0);
  PETSYN_NX_DISPATCH(fwd_kernel, grid, fwd_smem, as_stream(stream), d);
  return check_launch("normact fwd_kernel");
}

int32_t petsyn_normact_bwd(const petsyn_normact_desc* desc, void* stream) {
  Dev d;
  int32_t rc = to_dev(desc, &d);
  if (rc) return rc;
  PETSYN_REQUIRE(d.t1 != nullptr && d.dz != nullptr, "missing gradient source / destination");
  PETSYN_REQUIRE(d.mean == nullptr || d.sums != nullptr, "normalised backward needs the sums workspace");
  cudaStream_t st = as_stream(stream);
  dim3 grid((unsigned)row_blocks(d.rows, d.C, desc->nsamples), (unsigned)desc->nsamples);
  if (d.mean != nullptr) {
    const int nst = d.per_sample ? desc->nsamples : 1;
    PETSYN_REQUIRE(!desc->sums_precomputed || (desc->sums_prezeroed && d.per_sample && d.t2 == nullptr && d.dslope == nullptr),
                   "precomputed backward sums need a pre-zeroed per-sample workspace and one gradient source");
    if (!desc->sums_prezeroed) PETSYN_CHECK_CUDA(cudaMemsetAsync(d.sums, 0, (size_t)nst * 2 * d.C * sizeof(double), st));
    const int rpp = 256 / (d.C / 8);
    const size_t smem = PfRing::bytes(d.t2 ? 3 : 2) + (size_t)rpp * 2 * d.C * sizeof(float);
    if (!desc->sums_precomputed) {     // else: filled by the producing data-gradient convolution's epilogue
      PETSYN_NX_DISPATCH(bwd_reduce_kernel, grid, smem, st, d);
      rc = check_launch("normact bwd_reduce_kernel");
      if (rc) return rc;
    }
    const int gs = desc->group_size > 1 ? desc->group_size : 1;
    if (d.per_sample && (gs > 1 || d.gamma != nullptr) && !desc->separate_group_combine) {
      d.gc_gs = gs; d.gc_ns = nst; d.gc_acc = desc->affine_accumulate;      // combined inside bwd_apply_kernel's prologue
    } else if (d.per_sample && (gs > 1 || d.gamma != nullptr)) {
      // the sums workspace holds [S0|S1] for every sample followed by ka and kb: 4 * nsamples * C floats
      float* ka = reinterpret_cast<float*>(d.sums + (size_t)nst * 2 * d.C);
      float* kb = ka + (size_t)nst * d.C;
      const int total = nst * d.C;
      group_combine_kernel<<<(total + 127) / 128, 128, 0, st>>>(d.sums, d.gamma, d.rstd, ka, kb, d.dgamma, d.dbeta, d.rows,
                                                                d.C, nst, gs, desc->affine_accumulate);
      rc = check_launch("group_combine_kernel");
      if (rc) return rc;
      d.ka = ka;
      d.kb = kb;
    }
  }
  {
    const size_t apply_smem = PfRing::bytes((d.t2 ? 3 : 2) + (d.ex ? 1 : 0)) +
                              (d.dz_colsum ? (size_t)(256 / (d.C / 8)) * d.C * sizeof(float) : 0) +
                              (d.gc_gs > 0 ? (size_t)2 * d.C * sizeof(float) : 0);
    PETSYN_NX_DISPATCH(bwd_apply_kernel, grid, apply_smem, st, d);
  }
  return check_launch("normact bwd_apply_kernel");
}

int32_t petsyn_add_slice(const void* src, int32_t src_cstride, int32_t src_coff, void* dst, int32_t dst_cstride,
                         int32_t dst_coff, int64_t rows, int32_t c, int32_t accumulate, void* stream) {
  PETSYN_REQUIRE(src && dst && rows > 0, "bad argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048 && (src_cstride | src_coff | dst_cstride | dst_coff) % 8 == 0,
                 "channels, pitches and offsets must be multiples of 8");
  add_slice_kernel<<<row_blocks(rows, c, 1), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src), src_cstride, src_coff, reinterpret_cast<__nv_bfloat16*>(dst),
      dst_cstride, dst_coff, rows, c, accumulate);
  return check_launch("add_slice_kernel");
}

int32_t petsyn_colsum(const void* x, int32_t cstride, int32_t coff, double* out, int64_t rows, int32_t c, void* stream) {
  PETSYN_REQUIRE(x && out && rows > 0, "bad argument");
  PETSYN_REQUIRE(c % 8 == 0 && c >= 8 && c <= 2048 && (cstride | coff) % 8 == 0, "channels must be a multiple of 8");
  cudaStream_t st = as_stream(stream);
  PETSYN_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)c * sizeof(double), st));
  const int rpp = 256 / (c / 8);
  PETSYN_CHECK_CUDA(launch_pdl(colsum_kernel, dim3(row_blocks(rows, c, 1)), dim3(256), (size_t)rpp * c * sizeof(float), st,
                               reinterpret_cast<const __nv_bfloat16*>(x), cstride, coff, out, rows, c));
  return check_launch("colsum_kernel");
}

}  // extern "C"
