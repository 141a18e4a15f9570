// Host side of the convolution family: turns a petsyn_conv_desc into "tap programs" (which input view / voxel offset /
// merged kernel taps every GEMM K-block uses), packs weights accordingly, encodes TMA descriptors and launches the
// tcgen05 kernels of igemm_kernels.cuh.
//
// One formalism covers every operator on the hot path.  Per spatial axis an operator is a small table of
//   (output phase r, input phase b, voxel offset, {original kernel indices merged into this tap})
// and the 3-D program is the outer product of the three axis tables:
//   Conv s1            y[o]     = sum_k W[k] x[o + k - p]                      1 output view, 1 input view
//   Conv s2            y[o]     = sum_k W[k] x[2o + k - p]                     input read through 8 stride-2 phase views
//   Upsample x2 + Conv y[2q+r]  = sum_k W[k] x[q + floor((r + k - p)/2)]       8 output phase views; taps that hit the same
//                                                                              source voxel are merged (27 -> 8 per phase)
//   ConvTranspose s2   y[2q+b]  = sum_{k: b+p-k even} W[k] x[q + (b+p-k)/2]    8 output phase views
// and the backward-data operators are the same four shapes with x and y exchanged.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "igemm_kernels.cuh"
#include "slab_kernels.cuh"
#include "slab_epi_kernel.cuh"

namespace petsyn {

// ------------------------------------------------------------------------------------------------ axis programs
struct AxisTap {
  int a_phase;           // 0/1 phase of the (stride-2) input view, 0 when the input is not phased
  int off;               // offset in the input view's grid
  std::vector<int> ks;   // original kernel indices along this axis that land on this tap
};
struct AxisProg {
  bool a_phased = false;
  bool out_phased = false;
  std::vector<std::vector<AxisTap>> subs;   // [1] or [2 output phases]
};

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// gather with unit stride: in = out + sign*(k - p)
static AxisProg axis_s1(int k, int p, int sign) {
  AxisProg a;
  a.subs.resize(1);
  for (int i = 0; i < k; ++i) a.subs[0].push_back({0, sign * (i - p), {i}});
  return a;
}
// gather through stride-2 phase views: in = 2*out + k - p
static AxisProg axis_s2_gather(int k, int p) {
  AxisProg a;
  a.a_phased = true;
  a.subs.resize(1);
  for (int i = 0; i < k; ++i) {
    int e = i - p;
    int b = ((e % 2) + 2) % 2;
    a.subs[0].push_back({b, (e - b) / 2, {i}});
  }
  return a;
}
// scatter form (transposed / backward of stride 2): out = 2*q + b gets in[q + (b + p - k)/2] for matching parity
static AxisProg axis_s2_scatter(int k, int p) {
  AxisProg a;
  a.out_phased = true;
  a.subs.resize(2);
  for (int b = 0; b < 2; ++b)
    for (int i = 0; i < k; ++i)
      if (((b + p - i) % 2) == 0) a.subs[b].push_back({0, floordiv(b + p - i, 2), {i}});
  return a;
}
// nearest x2 upsample followed by a k-tap conv: out = 2q + r reads src[q + floor((r + k - p)/2)]; equal offsets merge
static AxisProg axis_up_fwd(int k, int p) {
  AxisProg a;
  a.out_phased = true;
  a.subs.resize(2);
  for (int r = 0; r < 2; ++r)
    for (int i = 0; i < k; ++i) {
      int off = floordiv(r + i - p, 2);
      bool merged = false;
      for (auto& t : a.subs[r])
        if (t.off == off) { t.ks.push_back(i); merged = true; }
      if (!merged) a.subs[r].push_back({0, off, {i}});
    }
  return a;
}
// backward-data of the above: dsrc[s] = sum_r sum_taps dy_r[s - off]
static AxisProg axis_up_bwd(int k, int p) {
  AxisProg f = axis_up_fwd(k, p);
  AxisProg a;
  a.a_phased = true;
  a.subs.resize(1);
  for (int r = 0; r < 2; ++r)
    for (auto& t : f.subs[r]) a.subs[0].push_back({r, -t.off, t.ks});
  return a;
}

// ------------------------------------------------------------------------------------------------ 3-D programs
struct TapDef {
  int a_view, dw, dh, dd;
  std::vector<int> src;   // flat indices (kd*k + kh)*k + kw of the original kernel taps summed into this tap
};
struct Program {
  bool a_phased = false, out_phased = false;
  std::vector<std::vector<TapDef>> subs;
  int max_taps = 0;
};

static Program make_program(const AxisProg& ax, int k) {
  Program P;
  P.a_phased = ax.a_phased;
  P.out_phased = ax.out_phased;
  const int np = (int)ax.subs.size();
  for (int rd = 0; rd < np; ++rd)
    for (int rh = 0; rh < np; ++rh)
      for (int rw = 0; rw < np; ++rw) {
        std::vector<TapDef> taps;
        for (auto& td : ax.subs[rd])
          for (auto& th : ax.subs[rh])
            for (auto& tw : ax.subs[rw]) {
              TapDef t;
              t.a_view = ax.a_phased ? (td.a_phase * 4 + th.a_phase * 2 + tw.a_phase) : 0;
              t.dd = td.off; t.dh = th.off; t.dw = tw.off;
              for (int kd : td.ks)
                for (int kh : th.ks)
                  for (int kw : tw.ks) t.src.push_back((kd * k + kh) * k + kw);
              taps.push_back(t);
            }
        P.max_taps = std::max(P.max_taps, (int)taps.size());
        P.subs.push_back(std::move(taps));
      }
  return P;
}

// ------------------------------------------------------------------------------------------------ device tables
struct TapSrcDev {
  int32_t nsrc;
  int32_t src[8];
};
struct InvEntryDev {   // for one original tap k: where its gradient contributions live in the packed scratch
  int32_t n;
  int32_t sub[8];
  int32_t tap[8];
};

// Weight packing / gradient unpacking move W between PyTorch layout W[a][b][k] (k fastest; (a,b) = (Cout,Cin) for
// Conv3d, (Cin,Cout) for ConvTranspose3d) and the GEMM operand layout B[(sub*rows + row)][t*kc_pad + col].
// Both go tile-wise through shared memory -- TA x TB (a,b) pairs x all k^3 taps, 256 pairs per block -- shaped so that
// the image's column index is the long tile edge (64): global reads AND writes are then 128-byte runs.  The tap count
// is a template parameter so that all index arithmetic is constant-folded.
struct PackImage {
  __nv_bfloat16* out;
  const TapSrcDev* tbl;     // [nsubs * max_taps]
  int32_t nsubs, max_taps;
  int32_t rows, kc_pad;     // rows per sub-problem, padded columns per tap
};

// TRANSPOSED = false: (row, col) = (a, b), tile 4 x 64;  true: (row, col) = (b, a), tile 64 x 4.
// NS = merged source taps per slot (1: plain convs, 8: Upsample+Conv); short slots are padded with a pointer to a
// zero element so that the inner sum is a fixed, fully unrolled chain of independent shared-memory loads.
template <int K3, bool TRANSPOSED, int NS>
__device__ __forceinline__ void pack_tile(const float* __restrict__ w, int A, int B, const PackImage& im, int bx, int by) {
  constexpr int TA = TRANSPOSED ? 64 : 4, TB = TRANSPOSED ? 4 : 64;
  constexpr int K3P = (K3 & 1) ? K3 + 2 : K3 + 1;    // odd pitch along b with at least one spare (zero) element
  constexpr int ROWP = (TB * K3P) | 1;               // odd pitch along a
  extern __shared__ float tile[];                    // [TA][ROWP]
  __shared__ int16_t s_src[kMaxTaps][NS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = bx * TA, b0 = by * TB;
  const int nslots = im.nsubs * im.max_taps;
  for (int i = tid; i < nslots * NS; i += 256) {
    const TapSrcDev& e = im.tbl[i / NS];
    const int q = i % NS;
    s_src[i / NS][q] = (int16_t)(q < e.nsrc ? e.src[q] : K3);   // K3 = index of the spare zero element
  }
  for (int i = tid; i < TA * ROWP; i += 256) tile[i] = 0.f;
  __syncthreads();
  // all 256 threads stream the tile: for every a the TB * K3 weights of b0 .. b0 + TB - 1 are one contiguous run
  for (int e = tid; e < TA * TB * K3; e += 256) {
    const int al = e / (TB * K3), idx = e - al * (TB * K3);
    const int bl = idx / K3, k = idx - bl * K3;
    const int a = a0 + al;
    if (a < A && b0 + bl < B) tile[al * ROWP + bl * K3P + k] = __ldg(w + ((int64_t)a * B + b0) * K3 + idx);
  }
  __syncthreads();
  // thread <-> one (row, col) of the tile; col is the fast index (64 consecutive columns per row)
  const int ci = tid & 63, ri = tid >> 6;
  const int row = TRANSPOSED ? b0 + ri : a0 + ri;
  const int col = TRANSPOSED ? a0 + ci : b0 + ci;
  if (row >= im.rows || col >= im.kc_pad) return;
  const float* mine = TRANSPOSED ? tile + ci * ROWP + ri * K3P : tile + ri * ROWP + ci * K3P;
  const int64_t ldb = (int64_t)im.max_taps * im.kc_pad;
  for (int sub = 0; sub < im.nsubs; ++sub) {
    __nv_bfloat16* dst = im.out + ((int64_t)sub * im.rows + row) * ldb + col;
    const int16_t(*ss)[NS] = s_src + sub * im.max_taps;
#pragma unroll 4
    for (int t = 0; t < im.max_taps; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < NS; ++q) acc += mine[ss[t][q]];
      dst[(int64_t)t * im.kc_pad] = __float2bfloat16(acc);
    }
  }
}

template <int K3, bool TRANSPOSED, int NS>
__global__ void __launch_bounds__(256) pack_weights_kernel(const float* __restrict__ w, int A, int B, PackImage im) {
  pack_tile<K3, TRANSPOSED, NS>(w, A, B, im, blockIdx.x, blockIdx.y);
}

// Batched form: one launch packs every weight tensor of a network that shares (K3, TRANSPOSED, NS); a block finds its
// job by binary search over the jobs' first-tile indices.
struct PackJob {
  const float* w;
  int32_t A, B;
  PackImage im;
  int32_t tile_begin, tiles_x;
};

template <int K3, bool TRANSPOSED, int NS>
__global__ void __launch_bounds__(256) pack_weights_multi_kernel(const PackJob* __restrict__ jobs, int njobs) {
  pdl_sync();
  int lo = 0, hi = njobs - 1;
  const int b = blockIdx.x;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].tile_begin <= b) lo = mid; else hi = mid - 1;
  }
  const PackJob& j = jobs[lo];
  const int local = b - j.tile_begin;
  pack_tile<K3, TRANSPOSED, NS>(j.w, j.A, j.B, j.im, local % j.tiles_x, local / j.tiles_x);
}

// scratch [nsubs*rows, max_taps*kc_pad] fp32 -> dw in PyTorch layout (sums the slots every original tap was merged into)
template <int K3, bool TRANSPOSED, int NS>
__global__ void __launch_bounds__(256) unpack_wgrad_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                           const InvEntryDev* __restrict__ inv, int A, int B, int nsubs,
                                                           int max_taps, int rows, int kc_pad, int accumulate,
                                                           int nsplit, const double* __restrict__ dbias_acc,
                                                           float* __restrict__ dbias, int nbias) {
  pdl_sync();
  if (dbias != nullptr && blockIdx.x == 0 && blockIdx.y == 0)       // bias gradient: double accumulator -> fp32 slot
    for (int i = threadIdx.x; i < nbias; i += 256) dbias[i] = accumulate ? dbias[i] + (float)dbias_acc[i] : (float)dbias_acc[i];
  constexpr int TA = TRANSPOSED ? 64 : 4, TB = TRANSPOSED ? 4 : 64;
  constexpr int kAP = TB + 1;                        // pitch along a inside a slot
  constexpr int kSlotPitch = (TA * kAP) | 1;
  extern __shared__ float tile[];                    // [nslots + 1][kSlotPitch]; the extra slot is all zeros
  __shared__ int16_t s_slot[K3][NS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int a0 = blockIdx.x * TA, b0 = blockIdx.y * TB;
  const int nslots = nsubs * max_taps;
  for (int i = tid; i < K3 * NS; i += 256) {
    const InvEntryDev& e = inv[i / NS];
    const int q = i % NS;
    s_slot[i / NS][q] = (int16_t)(q < e.n ? e.sub[q] * max_taps + e.tap[q] : nslots);
  }
  for (int i = tid; i < kSlotPitch; i += 256) tile[nslots * kSlotPitch + i] = 0.f;
  const int64_t ldb = (int64_t)max_taps * kc_pad;
  {
    const int ci = tid & 63, ri = tid >> 6;
    const int row = TRANSPOSED ? b0 + ri : a0 + ri, col = TRANSPOSED ? a0 + ci : b0 + ci;
    const int a_off = TRANSPOSED ? ci : ri, b_off = TRANSPOSED ? ri : ci;
    const bool ok = (a0 + a_off) < A && (b0 + b_off) < B;
    float* dst = tile + a_off * kAP + b_off;
    // the voxel reduction was split over `nsplit` CTAs, each adding into its own partial image: summed here in split
    // order, so the weight gradient does not depend on the order the CTAs finished in
    const int64_t image = (int64_t)nsubs * rows * ldb;
    for (int sub = 0; sub < nsubs; ++sub) {
      const float* src = scratch + ((int64_t)sub * rows + row) * ldb + col;
      if (nsplit == 1) {
#pragma unroll 8
        for (int t = 0; t < max_taps; ++t)
          dst[(sub * max_taps + t) * kSlotPitch] = ok ? __ldg(src + (int64_t)t * kc_pad) : 0.f;
      } else {
        for (int t = 0; t < max_taps; ++t) {
          float v = 0.f;
          if (ok) {
#pragma unroll 8
            for (int k = 0; k < nsplit; ++k) v += __ldg(src + k * image + (int64_t)t * kc_pad);   // split order: reproducible
          }
          dst[(sub * max_taps + t) * kSlotPitch] = v;
        }
      }
    }
  }
  __syncthreads();
  for (int al = warp; al < TA; al += 8) {
    const int a = a0 + al;
    if (a >= A) continue;
    float* dst = dw + ((int64_t)a * B + b0) * K3;
    const float* trow = tile + al * kAP;
#pragma unroll 2
    for (int idx = lane; idx < TB * K3; idx += 32) {
      const int bl = idx / K3, k = idx - bl * K3;
      if (b0 + bl >= B) continue;
      float acc = 0.f;
#pragma unroll
      for (int q = 0; q < NS; ++q) acc += trow[s_slot[k][q] * kSlotPitch + bl];
      if (accumulate) dst[idx] += acc; else dst[idx] = acc;
    }
  }
}

// small-channel wgrad scratch T[(sub*m_tiles + mt)*128 + tap_local*Cin + ci][co] -> dw[co][ci][k]
__global__ void __launch_bounds__(256) unpack_wgrad_small_kernel(const float* __restrict__ scratch,
                                                                 float* __restrict__ dw,
                                                                 const InvEntryDev* __restrict__ inv, int Cout, int Cin,
                                                                 int k3, int npad, int tpm, int m_tiles, int accumulate,
                                                                 int nsplit, int64_t image,
                                                                 const double* __restrict__ dbias_acc,
                                                                 float* __restrict__ dbias, int nbias) {
  pdl_sync();
  if (dbias != nullptr && blockIdx.x == 0)
    for (int i = threadIdx.x; i < nbias; i += 256) dbias[i] = accumulate ? dbias[i] + (float)dbias_acc[i] : (float)dbias_acc[i];
  const int total = Cout * Cin * k3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % k3;
    const int ci = (i / k3) % Cin;
    const int co = i / (k3 * Cin);
    const InvEntryDev e = inv[k];
    float acc = 0.f;
    for (int q = 0; q < e.n; ++q) {
      const int mt = e.tap[q] / tpm, tl = e.tap[q] - mt * tpm;
      const float* src = scratch + ((int64_t)(e.sub[q] * m_tiles + mt) * 128 + tl * Cin + ci) * npad + co;
#pragma unroll 8
      for (int ks = 0; ks < nsplit; ++ks) acc += __ldg(src + ks * image);   // partial images in split order (reproducible)
    }
    if (accumulate) dw[i] += acc; else dw[i] = acc;
  }
}

// split-K finish: `nslots` fp32 partial images [rows, C] (one per K split, zero where a split had no work), added in slot
// order (reproducible: no atomics, no reduction whose order depends on the run) -> act(x + bias) as bf16 into a channel slice
__global__ void __launch_bounds__(256) splitk_finish_kernel(const float* __restrict__ src, int nslots,
                                                            __nv_bfloat16* __restrict__ dst, int64_t rows, int C,
                                                            int cstride, int coff, const float* __restrict__ bias, int act,
                                                            float slope, int accumulate) {
  pdl_sync();
  const int cpt = C / 8;
  const int64_t total = rows * cpt;
  const int64_t slot = rows * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cpt;
    const int c = (int)(i - r * cpt) * 8;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nslots; ++k) {
      const float* sp = src + k * slot + r * C + c;
      const float4 v0 = *reinterpret_cast<const float4*>(sp), v1 = *reinterpret_cast<const float4*>(sp + 4);
      f[0] += v0.x; f[1] += v0.y; f[2] += v0.z; f[3] += v0.w; f[4] += v1.x; f[5] += v1.y; f[6] += v1.z; f[7] += v1.w;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (bias) f[q] += __ldg(bias + c + q);
      f[q] = apply_act(f[q], act, slope);
    }
    if (accumulate) {
      const uint4 old = *reinterpret_cast<const uint4*>(dst + r * cstride + coff + c);
      const __nv_bfloat162* ob = reinterpret_cast<const __nv_bfloat162*>(&old);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 of = __bfloat1622float2(ob[q]);
        f[2 * q] += of.x;
        f[2 * q + 1] += of.y;
      }
    }
    uint4 o;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
    o.x = *reinterpret_cast<uint32_t*>(&b0); o.y = *reinterpret_cast<uint32_t*>(&b1);
    o.z = *reinterpret_cast<uint32_t*>(&b2); o.w = *reinterpret_cast<uint32_t*>(&b3);
    *reinterpret_cast<uint4*>(dst + r * cstride + coff + c) = o;
  }
}

// The same finish pass, also summing what it stores: per (sample, channel) sum and sum of squares of the bf16 output into up
// to two [sample][2][st_c] double accumulators (the statistics pass of the normalisation that consumes a split-K convolution).
// grid (x, sample); thread = (8-channel group, row of the pass); C / 8 divides 256.
__global__ void __launch_bounds__(256) splitk_finish_stats_kernel(const float* __restrict__ src, int nslots,
                                                                  __nv_bfloat16* __restrict__ dst, int64_t rows_per_sample,
                                                                  int C, int cstride, int coff, const float* __restrict__ bias,
                                                                  int act, float slope, double* st1, int st1_c, int st1_off,
                                                                  double* st2, int st2_c, int st2_off) {
  pdl_sync();
  __shared__ float s_red[4096];                      // [rows of a pass][2][C]: (256 / cpt) * 2 * 8 * cpt floats
  const int cpt = C / 8, rpp = 256 / cpt;
  const int tx = threadIdx.x % cpt, ty = threadIdx.x / cpt;
  const int c = tx * 8;
  const int sample = blockIdx.y;
  const int64_t slot = rows_per_sample * gridDim.y * C;
  float a0[8], a1[8], bs[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) { a0[q] = a1[q] = 0.f; bs[q] = bias ? __ldg(bias + c + q) : 0.f; }
  for (int64_t rr = (int64_t)blockIdx.x * rpp + ty; rr < rows_per_sample; rr += (int64_t)gridDim.x * rpp) {
    const int64_t r = (int64_t)sample * rows_per_sample + rr;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nslots; ++k) {
      const float* sp = src + k * slot + r * C + c;
      const float4 v0 = *reinterpret_cast<const float4*>(sp), v1 = *reinterpret_cast<const float4*>(sp + 4);
      f[0] += v0.x; f[1] += v0.y; f[2] += v0.z; f[3] += v0.w; f[4] += v1.x; f[5] += v1.y; f[6] += v1.z; f[7] += v1.w;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) f[q] = apply_act(f[q] + bs[q], act, slope);
    uint4 o;
    __nv_bfloat162 b[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      b[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
      const float2 back = __bfloat1622float2(b[q]);            // statistics of the values AS STORED
      a0[2 * q] += back.x; a1[2 * q] += back.x * back.x;
      a0[2 * q + 1] += back.y; a1[2 * q + 1] += back.y * back.y;
    }
    o.x = *reinterpret_cast<uint32_t*>(&b[0]); o.y = *reinterpret_cast<uint32_t*>(&b[1]);
    o.z = *reinterpret_cast<uint32_t*>(&b[2]); o.w = *reinterpret_cast<uint32_t*>(&b[3]);
    *reinterpret_cast<uint4*>(dst + r * cstride + coff + c) = o;
  }
  float* mine = s_red + (size_t)ty * 2 * C;
#pragma unroll
  for (int q = 0; q < 8; ++q) { mine[c + q] = a0[q]; mine[C + c + q] = a1[q]; }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * C; e += 256) {
    float t = 0.f;
    for (int y = 0; y < rpp; ++y) t += s_red[(size_t)y * 2 * C + e];      // fixed order
    const int which = e / C, ch = e - which * C;
    if (st1 != nullptr) atomicAdd(st1 + ((int64_t)sample * 2 + which) * st1_c + st1_off + ch, (double)t);
    if (st2 != nullptr) atomicAdd(st2 + ((int64_t)sample * 2 + which) * st2_c + st2_off + ch, (double)t);
  }
}

// Partial images of a weight gradient (one per K split / per persistent CTA) -> image 0, added in image order: the result
// does not depend on the order the producing CTAs finished in (reproducible weight gradients without atomics).
// thread = (float4 group g, lane l): lane l adds images l, l + L, ... in ascending order, then the L lane sums are added in
// lane order by lane 0, which overwrites image 0.
__global__ void __launch_bounds__(256) reduce_images_kernel(float* __restrict__ scratch, int nimages, int64_t image_floats,
                                                            int lanes) {
  pdl_sync();
  __shared__ float4 part[256];
  const int gpb = 256 / lanes;
  const int64_t ng = image_floats >> 2;
  const int64_t g = (int64_t)blockIdx.x * gpb + threadIdx.x % gpb;
  const int l = threadIdx.x / gpb;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g < ng) {
    const float4* src = reinterpret_cast<const float4*>(scratch) + g;
#pragma unroll 4
    for (int k = l; k < nimages; k += lanes) {
      const float4 v = __ldcg(src + (int64_t)k * ng);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  part[threadIdx.x] = a;
  __syncthreads();
  if (l == 0 && g < ng) {
    float4 t = part[threadIdx.x];
    for (int q = 1; q < lanes; ++q) {
      const float4 v = part[q * gpb + threadIdx.x];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(scratch)[g] = t;
  }
}

static int32_t reduce_images(float* scratch, int nimages, int64_t image_floats, cudaStream_t st) {
  if (nimages <= 1) return PETSYN_OK;
  int lanes = 1;
  while (lanes < 32 && lanes * 4 <= nimages) lanes *= 2;
  const int gpb = 256 / lanes;
  const int64_t ng = image_floats >> 2;
  PETSYN_CHECK_CUDA(launch_pdl(reduce_images_kernel, dim3((unsigned)((ng + gpb - 1) / gpb)), dim3(256), 0, st, scratch, nimages, image_floats, lanes));
  return check_launch("reduce_images_kernel");
}

// ------------------------------------------------------------------------------------------------ plan
struct GemmSide {           // one gather-form GEMM (fprop or dgrad)
  Program prog;
  int R = 0, Kc = 0;        // output channels / reduction channels of this GEMM
  int kc_pad = 0, kch = 64; // K-chunk (channels per pipeline stage)
  bool swap = false;        // packed B reads W[c][r][k] instead of W[r][c][k]
  int out_d = 0, out_h = 0, out_w = 0;   // grid of ONE output view (phase grid if out_phased)
  int box_w = 0, box_h = 0, box_d = 0;
  int block_n = 128;
  bool out_fp32 = false;
  bool accumulate = false;    // add into the destination (bf16 TMA reduction) instead of overwriting
  int ksplit = 1;             // >1: fp32 split-K through `workspace`, finished by splitk_finish_kernel
  void* workspace = nullptr;
  size_t workspace_bytes = 0;
  int64_t out_rows_full = 0;  // voxels of the full (un-phased) output tensor
  IgemmTap* d_taps = nullptr;
  TapSrcDev* d_src = nullptr;
  std::vector<IgemmSub> subs;
  // cached TMA descriptors, keyed by the base pointers they were built for
  const void* key_a = nullptr; const void* key_b = nullptr; const void* key_c = nullptr; const void* key_w = nullptr;
  IgemmParams params;
  // small-channel slab path (slab_kernels.cuh): k3 s1 p1, 16..64 channels, many voxels
  bool slab = false;
  int slab_halo = 1;        // 1: 3x3x3, 0: 1x1x1
  bool slab3 = false;       // depth-folded kernel (slab_conv3_kernel): three depth taps per MMA
  int slab_grid = 0, slab_smem = 0;
  SlabParams sparams;
  // fused-epilogue variant of the depth-folded slab kernel (slab_epi_kernel.cuh); geometry re-planned for its smem footprint
  SlabParams eparams;
  SlabEpi epi;
  int epi_grid = 0, epi_smem = 0;
  const void* ekey_a = nullptr; const void* ekey_b = nullptr; const void* ekey_c = nullptr; const void* ekey_side = nullptr;
  int ekey_cs = -1, ekey_co = -1;
};

struct ViewSpec {           // an NDHWC tensor (channel slice) and how to derive its tensor maps
  int C, cstride, coff;
  int W, H, D, N;           // full (un-phased) spatial dims
};

}  // namespace petsyn

struct petsyn_pack_batch {
  petsyn::PackJob* d_jobs[16] = {nullptr};
  int njobs[16] = {0};
  int tiles[16] = {0};
};

struct petsyn_conv_plan {
  petsyn_conv_desc desc;
  int k3 = 0;
  int od = 0, oh = 0, ow = 0;   // forward output dims
  petsyn::ViewSpec vx, vy, vdx, vdy;
  petsyn::GemmSide fprop, dgrad;
  // wgrad (uses the fprop program)
  petsyn::InvEntryDev* d_inv = nullptr;
  int inv_max = 1;              // max number of packed slots one original tap contributes to
  int wg_box_w = 0, wg_box_h = 0, wg_box_d = 0;
  int wg_block_n = 128, wg_ksplit = 1;
  bool wg_small = false;        // small-channel path: taps folded into the MMA M dimension (wgrad_small_kernel)
  int wg_tpm = 1, wg_mtiles = 1, wg_npad = 16;
  petsyn::WgradSmallParams wgs_params;
  const void* wg_key_x = nullptr; const void* wg_key_g = nullptr; const void* wg_key_s = nullptr;
  petsyn::WgradParams wg_params;
  bool wg_slab = false;         // slab path (slab_wgrad_kernel): k3 s1 p1 or k1, Cin <= 96, many voxels
  int wg_slab_halo = 1;
  int wg_slab_grid = 0, wg_slab_smem = 0;
  petsyn::SlabWgradParams wgl_params;
};

namespace petsyn {

static void choose_box(int W, int H, int D, int max_rows, bool exact, int* bw, int* bh, int* bd) {
  // maximise useful rows per tile; prefer long W runs (contiguous memory).  When `exact` the box must hold exactly
  // max_rows voxels (wgrad: stale smem rows would pollute the reduction) and may overhang the tensor.
  double best = -1;
  *bw = *bh = *bd = 1;
  for (int w = 1; w <= max_rows; ++w) {
    if (!exact && w > W) break;
    for (int h = 1; w * h <= max_rows; ++h) {
      if (!exact && h > H) break;
      for (int d = 1; w * h * d <= max_rows; ++d) {
        if (!exact && d > D) break;
        if (exact && w * h * d != max_rows) continue;
        const int64_t tiles = (int64_t)((W + w - 1) / w) * ((H + h - 1) / h) * ((D + d - 1) / d);
        const double eff = (double)W * H * D / ((double)tiles * max_rows);
        const double score = eff + 1e-4 * w - 1e-6 * d;
        if (score > best) { best = score; *bw = w; *bh = h; *bd = d; }
      }
    }
  }
}

static int32_t upload(const void* src, size_t bytes, void** dst) {
  PETSYN_CHECK_CUDA(cudaMalloc(dst, bytes));
  PETSYN_CHECK_CUDA(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
  return PETSYN_OK;
}

// persistent-CTA work split of the slab kernels: columns of (tile_w x tile_h) tiles cut into depth chunks
static void slab_split(int W, int H, int D, int N, int tile_w, int tile_h, int ctas, int* dchunk, int* nchunks, int* items) {
  const int cols = ((W + tile_w - 1) / tile_w) * ((H + tile_h - 1) / tile_h) * N;
  double best = -1;
  *dchunk = D; *nchunks = 1;
  for (int dc = std::min(D, 4); dc <= D; ++dc) {
    const int nc = (D + dc - 1) / dc;
    const int it = cols * nc;
    const int waves = (it + ctas - 1) / ctas;
    // longest CTA processes `waves` items of (dc + 2) slab loads and dc tiles each
    const double cost = (double)waves * (dc + 0.35 * 2.0);
    const double score = (double)cols * D / ((double)ctas * cost);
    if (score > best) { best = score; *dchunk = dc; *nchunks = nc; }
  }
  *items = cols * *nchunks;
}

static int32_t finish_side(GemmSide& g, int N, int allow_slab = 0 /* 0 no, 1 = k3 s1 p1, 2 = k1 s1 p0 */) {
  g.kch = 64;
  g.kc_pad = (g.Kc + g.kch - 1) / g.kch * g.kch;
  if (g.R >= 128) g.block_n = 128;
  else g.block_n = (g.R + 15) / 16 * 16;
  // 256 output channels per CTA: the 128 x 64 activation tile of a K step is read from shared memory once for twice the MMA
  // work (SS-mode tcgen05.mma is bound by its operand reads at N = 128); needs enough tiles to fill the machine without it
  if (g.R % 256 == 0 && getenv("PETSYN_BN256") != nullptr) g.block_n = 256;
  if (g.block_n > 64 && g.block_n < 128) g.block_n = 128;
  choose_box(g.out_w, g.out_h, g.out_d, 128, false, &g.box_w, &g.box_h, &g.box_d);
  {
    // split-K when the output tiles alone cannot fill the machine (deep, small-M layers whose cost is streaming weights)
    const int64_t tiles = (int64_t)((g.out_w + g.box_w - 1) / g.box_w) * ((g.out_h + g.box_h - 1) / g.box_h) *
                          ((g.out_d + g.box_d - 1) / g.box_d) * N * ((g.R + g.block_n - 1) / g.block_n) *
                          (int64_t)g.prog.subs.size();
    int min_steps = 1 << 30;
    for (auto& sp : g.prog.subs) min_steps = std::min<int>(min_steps, (int)sp.size() * (g.kc_pad / g.kch));
    int ks = (int)(296 / std::max<int64_t>(tiles, 1));
    ks = std::min(ks, std::max(1, min_steps / 8));        // (<= min_steps: no split of any sub-problem is ever empty)
    g.ksplit = (ks >= 2 && !g.out_fp32) ? ks : 1;
  }
  std::vector<IgemmTap> taps;
  std::vector<TapSrcDev> srcs(g.prog.subs.size() * g.prog.max_taps);
  memset(srcs.data(), 0, srcs.size() * sizeof(TapSrcDev));
  g.subs.clear();
  for (size_t s = 0; s < g.prog.subs.size(); ++s) {
    IgemmSub sub;
    sub.c_view = g.prog.out_phased ? (int)s : 0;
    sub.b_row = (int)s * g.R;
    sub.tap_begin = (int)taps.size();
    sub.tap_count = (int)g.prog.subs[s].size();
    if (sub.tap_count > kMaxTaps) return fail(PETSYN_EINVAL, "too many taps per sub-problem (%d)", sub.tap_count);
    for (size_t t = 0; t < g.prog.subs[s].size(); ++t) {
      const TapDef& td = g.prog.subs[s][t];
      taps.push_back({td.a_view, td.dw, td.dh, td.dd});
      TapSrcDev& e = srcs[s * g.prog.max_taps + t];
      if (td.src.size() > 8) return fail(PETSYN_EINVAL, "more than 8 merged kernel taps");
      e.nsrc = (int)td.src.size();
      for (size_t j = 0; j < td.src.size(); ++j) e.src[j] = td.src[j];
    }
    g.subs.push_back(sub);
  }
  if (allow_slab && (!g.out_fp32 || g.R <= 32) && g.Kc % 16 == 0 && g.Kc <= 64 && g.R % 16 == 0 && g.R <= 64 &&
      slab_smem_bytes(allow_slab == 1 ? 27 : 1, g.Kc / 16, g.R, ((g.Kc / 16) * kSlabWp * kSlabHp * 32 + 1023) / 1024 * 1024, 4,
                      g.out_fp32 ? 4 : 2) <= 224 * 1024 &&
      (int64_t)g.out_w * g.out_h * g.out_d * N >= 128 * 148) {
    g.slab = true;
    g.slab_halo = allow_slab == 1 ? 1 : 0;
    g.ksplit = 1;
    g.block_n = g.R;
  }
  int32_t rc = upload(taps.data(), taps.size() * sizeof(IgemmTap), (void**)&g.d_taps);
  if (rc) return rc;
  return upload(srcs.data(), srcs.size() * sizeof(TapSrcDev), (void**)&g.d_src);
}

static int32_t view_map(CUtensorMap* out, const void* base, const ViewSpec& v, bool phased, int phase, int esz,
                        CUtensorMapDataType dt, int box_c, int bw, int bh, int bd, int swizzle);

// tensor map of an NDHWC bf16 view for the slab kernels: dims (16 ch, W, atoms, H, D*N) so that one box load lands as
// [h][atom][w][16 ch] in shared memory (32B swizzle)
static int32_t slab_view_map(CUtensorMap* out, const void* base, const ViewSpec& v, int atoms, int box_w, int box_h,
                             int box_atoms = 0) {
  const uint64_t cs = (uint64_t)v.cstride * 2;
  const uint8_t* b = reinterpret_cast<const uint8_t*>(base) + (uint64_t)v.coff * 2;
  uint64_t dims[5] = {16, (uint64_t)v.W, (uint64_t)atoms, (uint64_t)v.H, (uint64_t)v.D * v.N};
  uint64_t strides[4] = {cs, 32, cs * v.W, cs * v.W * v.H};
  uint32_t box[5] = {16u, (uint32_t)box_w, (uint32_t)(box_atoms > 0 ? box_atoms : atoms), (uint32_t)box_h, 1u};
  return encode_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, b, dims, strides, box, 32);
}

static int32_t bind_slab(GemmSide& g, const ViewSpec& va, const ViewSpec& vc, const void* a, const void* b, const void* c,
                         const float* bias, int act, float slope) {
  SlabParams& p = g.sparams;
  p.bias = bias; p.epi_act = act; p.epi_slope = slope;
  p.reduce = g.accumulate ? 1 : 0;
  if (g.key_a == a && g.key_b == b && g.key_c == c) return PETSYN_OK;
  const int atoms = g.Kc / 16;
  const int halo = g.slab_halo;
  int32_t rc = slab_view_map(&p.a_map, a, va, atoms, kSlabW + 2 * halo, kSlabH + 2 * halo);
  if (rc) return rc;
  {
    uint64_t dims[2] = {(uint64_t)g.prog.max_taps * g.kc_pad, (uint64_t)g.subs.size() * g.R};
    uint64_t strides[1] = {dims[0] * 2};
    uint32_t box[2] = {16u, (uint32_t)g.R};
    rc = encode_tmap(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, strides, box, 32);
    if (rc) return rc;
  }
  rc = view_map(&p.c_map, c, vc, false, 0, g.out_fp32 ? 4 : 2,
                g.out_fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g.R, kSlabW, kSlabH, 1, 0);
  p.out_f32 = g.out_fp32 ? 1 : 0;
  if (rc) return rc;
  if (g.prog.subs.size() != 1 || (int)g.prog.subs[0].size() != (halo ? 27 : 1))
    return fail(PETSYN_EINVAL, "slab path needs one 27-tap (k3) or 1-tap (k1) program");
  for (int t = 0; t < (halo ? 27 : 0); ++t) {
    const TapDef& td = g.prog.subs[0][t];
    if (td.dd != g.prog.subs[0][t / 9 * 9].dd) return fail(PETSYN_EINVAL, "slab path: taps are not grouped by depth offset");
    p.tap_off[t] = (((td.dh + 1) * atoms) * (kSlabWp * 32) + (td.dw + 1) * 32) >> 4;
    p.tap_slab[t / 9] = td.dd + 1;
  }
  g.slab3 = false;
  if (halo) {
    // depth-folded variant: pair j = (dh, dw); B block b of pair j is the tap (dd5[b], dh, dw)
    static const int dd5[5] = {+1, 0, -1, +1, 0};
    bool ok = true;
    for (int j = 0; j < 9 && ok; ++j) {
      const int dh = j / 3 - 1, dw = j % 3 - 1;
      p.pair_off[j] = (((dh + 1) * atoms) * (kSlabWp * 32) + (dw + 1) * 32) >> 4;
      for (int b = 0; b < 5; ++b) {
        int found = -1;
        for (int t = 0; t < 27; ++t) {
          const TapDef& td = g.prog.subs[0][t];
          if (td.dd == dd5[b] && td.dh == dh && td.dw == dw) found = t;
        }
        if (found < 0) ok = false;
        p.wtap[j][b] = found;
      }
    }
    const int slab_b = (atoms * kSlabWp * kSlabHp * 32 + 1023) / 1024 * 1024;
    g.slab3 = ok && getenv("PETSYN_NO_SLAB3") == nullptr &&
              slab3_smem_bytes(atoms, g.R, slab_b, 4, g.out_fp32 ? 4 : 2) <= 200 * 1024;
  }
  p.ntaps = g.subs[0].tap_count; p.atoms = atoms; p.kc_pad = g.kc_pad; p.b_row = g.subs[0].b_row;
  p.block_n = g.R; p.rows = g.R;
  p.W = va.W; p.H = va.H; p.D = va.D; p.batch = va.N;
  p.tiles_w = (va.W + kSlabW - 1) / kSlabW;
  p.tiles_h = (va.H + kSlabH - 1) / kSlabH;
  p.halo = halo;
  p.slab_tx = atoms * (kSlabW + 2 * halo) * (kSlabH + 2 * halo) * 32;
  p.slab_bytes = (p.slab_tx + 1023) / 1024 * 1024;
  // ring depth and CTAs per SM from the shared-memory budget
  const int fixed = g.slab3 ? slab3_smem_bytes(atoms, g.R, p.slab_bytes, 0, g.out_fp32 ? 4 : 2)
                            : slab_smem_bytes(p.ntaps, atoms, g.R, p.slab_bytes, 0, g.out_fp32 ? 4 : 2);
  int ring = (fixed + 8 * p.slab_bytes <= 110 * 1024) ? 8 : 4;
  if (g.slab3) ring = 4;                                  // the depth-folded kernel keeps one slab live + prefetch
  if (g.slab3 && getenv("PETSYN_SLAB3_RING")) ring = atoi(getenv("PETSYN_SLAB3_RING")) == 2 ? 2 : 4;   // tuning experiments only
  if (const char* e = getenv("PETSYN_SLAB_RING")) ring = atoi(e) == 4 ? 4 : 8;       // tuning experiments only
  p.ring = ring;
  g.slab_smem = fixed + ring * p.slab_bytes;
  int occ = std::max(1, std::min(g.slab3 ? 4 : 3, (227 * 1024) / (g.slab_smem + 1024)));
  if (const char* e = getenv("PETSYN_SLAB_OCC")) occ = std::max(1, std::min(atoi(e), (227 * 1024) / (g.slab_smem + 1024)));
  const int ctas = 148 * occ;
  slab_split(va.W, va.H, va.D, va.N, kSlabW, kSlabH, ctas, &p.dchunk, &p.nchunks, &p.items);
  g.slab_grid = std::min(ctas, p.items);
  g.key_a = a; g.key_b = b; g.key_c = c;
  return PETSYN_OK;
}

template <int ATOMS>
static int32_t launch_slab3(const GemmSide& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(slab_conv3_kernel<ATOMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  PETSYN_CHECK_CUDA(launch_pdl(slab_conv3_kernel<ATOMS>, dim3(g.slab_grid), dim3(192), g.slab_smem, st, g.sparams));
  return check_launch("slab_conv3_kernel");
}

template <int ATOMS>
static int32_t launch_slab(const GemmSide& g, cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(slab_conv_kernel<ATOMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  PETSYN_CHECK_CUDA(launch_pdl(slab_conv_kernel<ATOMS>, dim3(g.slab_grid), dim3(192), g.slab_smem, st, g.sparams));
  return check_launch("slab_conv_kernel");
}

static int32_t run_slab(GemmSide& g, cudaStream_t st) {
  if (g.slab3) {
    switch (g.Kc / 16) {
      case 1: return launch_slab3<1>(g, st);
      case 2: return launch_slab3<2>(g, st);
      case 3: return launch_slab3<3>(g, st);
      case 4: return launch_slab3<4>(g, st);
      default: return fail(PETSYN_EINVAL, "slab path: unsupported channel count %d", g.Kc);
    }
  }
  switch (g.Kc / 16) {
    case 1: return launch_slab<1>(g, st);
    case 2: return launch_slab<2>(g, st);
    case 3: return launch_slab<3>(g, st);
    case 4: return launch_slab<4>(g, st);
    default: return fail(PETSYN_EINVAL, "slab path: unsupported channel count %d", g.Kc);
  }
}

// ---------------------------------------------------------------------------------------------- fused epilogue
static bool epi_supported(const GemmSide& g, int batch) {
  if (!g.slab || g.slab_halo != 1 || g.out_fp32 || getenv("PETSYN_NO_EPI") != nullptr || getenv("PETSYN_NO_SLAB3") != nullptr)
    return false;
  if (!(g.R == 16 || g.R == 32) || g.Kc % 16 != 0 || g.Kc / 16 < 1 || g.Kc / 16 > 3 || batch > kEpiMaxBatch) return false;
  const int atoms = g.Kc / 16;
  const int slab_b = (atoms * kSlabWp * kSlabHp * 32 + 1023) / 1024 * 1024;
  if (slab3_smem_bytes(atoms, g.R, slab_b, 4, 2) > 200 * 1024) return false;          // same bound as bind_slab's slab3 choice
  return slab3_smem_bytes(atoms, g.R, slab_b, 4, 2) + slab3_epi_extra_bytes(g.R, 2, batch) <= 224 * 1024;
}

// `g` has been bound by bind_slab for (a, b, c); plan the fused variant and fill its epilogue block
static int32_t run_slab_epi(GemmSide& g, const ViewSpec& va, const ViewSpec& vc, const void* a, const void* b, const void* c,
                            const petsyn_conv_epilogue* ep, bool is_dgrad, cudaStream_t st) {
  if (!g.slab3) return fail(PETSYN_EINVAL, "fused epilogue needs the depth-folded slab kernel");
  static const petsyn_conv_epilogue none = {};
  const bool plain = ep == nullptr;       // the same kernel template without epilogue work (channel counts known at compile time)
  if (plain) ep = &none;
  if (g.accumulate && !plain) return fail(PETSYN_EINVAL, "fused epilogue cannot add into the destination");
  const int N = g.R, atoms = g.Kc / 16;
  const bool side = ep->side != nullptr;
  PETSYN_REQUIRE(!(ep->add_side && !side), "add_side without a side tensor");
  PETSYN_REQUIRE(plain || !is_dgrad || (side && ep->norm_scale && ep->norm_shift && ep->norm_mean && ep->norm_rstd && ep->bsums),
                 "dgrad epilogue needs z and the normalisation constants");
  PETSYN_REQUIRE(plain || is_dgrad || ep->add_side || ep->stats1 || ep->stats2, "empty epilogue");
  SlabParams& p = g.eparams;
  const int flags = plain ? 0 : ((ep->add_side ? kEpiSide : 0) | ((!is_dgrad && (ep->stats1 || ep->stats2)) ? kEpiStats : 0) |
                                 (is_dgrad ? kEpiNormReduce : 0));
  const bool rekey = !(g.ekey_a == a && g.ekey_b == b && g.ekey_c == c && g.ekey_side == ep->side &&
                       g.ekey_cs == ep->side_cstride && g.ekey_co == ep->side_coff && g.epi.flags == flags);
  if (rekey) {
    p = g.sparams;
    // shared-memory plan: a side ring of 4 tiles unless that costs a resident CTA (occupancy from the runtime: registers count)
    const int fixed = slab3_smem_bytes(atoms, N, p.slab_bytes, 0, 2) + p.ring * p.slab_bytes;
    const int smem4 = fixed + slab3_epi_extra_bytes(N, 4, va.N), smem2 = fixed + slab3_epi_extra_bytes(N, side ? 2 : 0, va.N);
    const int occ4 = (side && smem4 <= 224 * 1024) ? slab3_epi_occupancy(atoms, N / 16, flags, smem4) : 0;
    const int occ2 = slab3_epi_occupancy(atoms, N / 16, flags, smem2);
    const int ring = (occ4 >= occ2 && occ4 > 0) ? 4 : 2;          // (without a side tensor the ring is never touched)
    int occ = ring == 4 ? occ4 : occ2;
    if (occ < 1) return fail(PETSYN_ECUDA, "fused epilogue kernel does not fit an SM (%d bytes of shared memory)", smem2);
    if (getenv("PETSYN_DEBUG_PLAN"))
      fprintf(stderr, "[petsyn] epi plan: atoms %d N %d flags %d side ring %d smem %d occ %d\n", atoms, N, flags, ring,
              ring == 4 ? smem4 : smem2, occ);
    g.epi.side_ring = side ? ring : 0;
    g.epi_smem = ring == 4 ? smem4 : smem2;
    if (const char* e = getenv("PETSYN_SLAB_OCC")) occ = std::max(1, std::min(atoi(e), occ));
    const int ctas = 148 * occ;
    slab_split(va.W, va.H, va.D, va.N, kSlabW, kSlabH, ctas, &p.dchunk, &p.nchunks, &p.items);
    g.epi_grid = std::min(ctas, p.items);
    if (side) {
      ViewSpec vs = vc;
      vs.cstride = ep->side_cstride > 0 ? ep->side_cstride : N;
      vs.coff = ep->side_coff;
      PETSYN_REQUIRE(vs.cstride % 8 == 0 && vs.coff % 8 == 0 && vs.coff + N <= vs.cstride, "bad side channel slice");
      int32_t rc = view_map(&g.epi.e_map, ep->side, vs, false, 0, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, N, kSlabW, kSlabH, 1, 0);
      if (rc) return rc;
    }
    g.ekey_a = a; g.ekey_b = b; g.ekey_c = c; g.ekey_side = ep->side; g.ekey_cs = ep->side_cstride; g.ekey_co = ep->side_coff;
  }
  // per-call scalars (bind_slab refreshed them in sparams)
  p.bias = g.sparams.bias; p.epi_act = g.sparams.epi_act; p.epi_slope = g.sparams.epi_slope; p.reduce = g.sparams.reduce;
  SlabEpi& e = g.epi;
  e.flags = flags;
  e.nact = ep->norm_act; e.nslope = ep->norm_slope;
  e.st1 = ep->stats1; e.st1_c = ep->stats1_c; e.st1_off = ep->stats1_coff;
  e.st2 = ep->stats2; e.st2_c = ep->stats2_c; e.st2_off = ep->stats2_coff;
  PETSYN_REQUIRE(!e.st1 || (e.st1_off >= 0 && e.st1_off + N <= e.st1_c), "statistics channel range outside the target");
  PETSYN_REQUIRE(!e.st2 || (e.st2_off >= 0 && e.st2_off + N <= e.st2_c), "statistics channel range outside the target");
  e.nscale = ep->norm_scale; e.nshift = ep->norm_shift; e.nmean = ep->norm_mean; e.nrstd = ep->norm_rstd;
  e.bsums = ep->bsums;
  return slab3_epi_launch(atoms, N / 16, flags, p, e, g.epi_grid, g.epi_smem, st);
}

// tensor map of (a phase of) an NDHWC bf16/fp32 view; box = (box_c, bw, bh, bd, 1)
static int32_t view_map(CUtensorMap* out, const void* base, const ViewSpec& v, bool phased, int phase, int esz,
                        CUtensorMapDataType dt, int box_c, int bw, int bh, int bd, int swizzle) {
  const int step = phased ? 2 : 1;
  const int pd = phased ? ((phase >> 2) & 1) : 0, ph = phased ? ((phase >> 1) & 1) : 0, pw = phased ? (phase & 1) : 0;
  const uint64_t cs = (uint64_t)v.cstride * esz;
  const uint8_t* b = reinterpret_cast<const uint8_t*>(base) + (uint64_t)v.coff * esz +
                     (((uint64_t)pd * v.H + ph) * v.W + pw) * cs;
  uint64_t dims[5] = {(uint64_t)v.C, (uint64_t)((v.W - pw + step - 1) / step), (uint64_t)((v.H - ph + step - 1) / step),
                      (uint64_t)((v.D - pd + step - 1) / step), (uint64_t)v.N};
  uint64_t strides[4] = {cs * step, cs * v.W * step, cs * v.W * v.H * step, cs * v.W * v.H * v.D};
  uint32_t box[5] = {(uint32_t)box_c, (uint32_t)bw, (uint32_t)bh, (uint32_t)bd, 1u};
  return encode_tmap(out, dt, 5, b, dims, strides, box, swizzle);
}

template <int BN, int OUT_MODE, bool STATS = false>
static int32_t launch_igemm_bn(const GemmSide& g, dim3 grid, cudaStream_t st) {
  constexpr int STAGES = (BN <= 128) ? 3 : 4;
  using Cfg = IgemmCfg<BN, 64, STAGES, OUT_MODE>;
  auto kern = igemm_kernel<BN, 64, STAGES, OUT_MODE, STATS>;
  static bool attr_set = false;   // per instantiation
  if (!attr_set) {
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  PETSYN_CHECK_CUDA(launch_pdl(kern, grid, dim3(128), Cfg::kSmemBytes, st, g.params));
  return check_launch("igemm_kernel");
}

static int32_t launch_igemm(const GemmSide& g, int batch, cudaStream_t st) {
  dim3 grid;
  grid.x = (unsigned)(g.params.tiles_w * g.params.tiles_h * g.params.tiles_d * batch);
  grid.y = (unsigned)((g.R + g.block_n - 1) / g.block_n);
  grid.z = (unsigned)(g.subs.size() * g.ksplit);
  if (g.accumulate && g.ksplit == 1) {
    switch (g.block_n) {
      case 16: return launch_igemm_bn<16, OUT_BF16_REDUCE>(g, grid, st);
      case 32: return launch_igemm_bn<32, OUT_BF16_REDUCE>(g, grid, st);
      case 48: return launch_igemm_bn<48, OUT_BF16_REDUCE>(g, grid, st);
      case 64: return launch_igemm_bn<64, OUT_BF16_REDUCE>(g, grid, st);
      case 128: return launch_igemm_bn<128, OUT_BF16_REDUCE>(g, grid, st);
    case 256: return launch_igemm_bn<256, OUT_BF16_REDUCE>(g, grid, st);
      default: return fail(PETSYN_EINVAL, "unsupported BLOCK_N %d for accumulating output", g.block_n);
    }
  }
  if (g.ksplit > 1) {
    switch (g.block_n) {
      case 16: return launch_igemm_bn<16, OUT_F32_REDUCE>(g, grid, st);
      case 32: return launch_igemm_bn<32, OUT_F32_REDUCE>(g, grid, st);
      case 48: return launch_igemm_bn<48, OUT_F32_REDUCE>(g, grid, st);
      case 64: return launch_igemm_bn<64, OUT_F32_REDUCE>(g, grid, st);
      case 128: return launch_igemm_bn<128, OUT_F32_REDUCE>(g, grid, st);
    case 256: return launch_igemm_bn<256, OUT_F32_REDUCE>(g, grid, st);
      default: return fail(PETSYN_EINVAL, "unsupported BLOCK_N %d for split-K", g.block_n);
    }
  }
  if (g.out_fp32) {
    switch (g.block_n) {
      case 16: return launch_igemm_bn<16, OUT_F32>(g, grid, st);
      case 32: return launch_igemm_bn<32, OUT_F32>(g, grid, st);
      case 48: return launch_igemm_bn<48, OUT_F32>(g, grid, st);
      case 64: return launch_igemm_bn<64, OUT_F32>(g, grid, st);
      case 128: return launch_igemm_bn<128, OUT_F32>(g, grid, st);
    case 256: return launch_igemm_bn<256, OUT_F32>(g, grid, st);
      default: return fail(PETSYN_EINVAL, "unsupported BLOCK_N %d for fp32 output", g.block_n);
    }
  }
  switch (g.block_n) {
    case 16: return launch_igemm_bn<16, OUT_BF16>(g, grid, st);
    case 32: return launch_igemm_bn<32, OUT_BF16>(g, grid, st);
    case 48: return launch_igemm_bn<48, OUT_BF16>(g, grid, st);
    case 64: return launch_igemm_bn<64, OUT_BF16>(g, grid, st);
    case 128: return launch_igemm_bn<128, OUT_BF16>(g, grid, st);
    case 256: return launch_igemm_bn<256, OUT_BF16>(g, grid, st);
    default: return fail(PETSYN_EINVAL, "unsupported BLOCK_N %d", g.block_n);
  }
}

// Output statistics in the gather-form kernel's epilogue (petsyn_conv_fprop_epi with statistics targets only): a plain
// bf16 store of one un-phased output view, no split-K, 32 / 64 / 128 channels per tile
static bool finish_stats_ok(const GemmSide& g, int batch) {
  const int cpt = g.R / 8;
  return g.R % 8 == 0 && cpt >= 1 && cpt <= 256 && 256 % cpt == 0 && !g.accumulate && g.out_rows_full % batch == 0;
}

static bool igemm_stats_supported(const GemmSide& g, int batch) {
  if (g.slab || g.out_fp32 || g.accumulate || getenv("PETSYN_NO_EPI") != nullptr || getenv("PETSYN_NO_IGEMM_STATS") != nullptr)
    return false;
  if (g.ksplit > 1) return finish_stats_ok(g, batch) && getenv("PETSYN_NO_FINISH_STATS") == nullptr;   // in the finish pass
  return g.subs.size() == 1 && !g.prog.out_phased && (g.block_n == 32 || g.block_n == 64 || g.block_n == 128);
}

static int32_t launch_igemm_stats(const GemmSide& g, int batch, cudaStream_t st) {
  dim3 grid;
  grid.x = (unsigned)(g.params.tiles_w * g.params.tiles_h * g.params.tiles_d * batch);
  grid.y = (unsigned)((g.R + g.block_n - 1) / g.block_n);
  grid.z = 1;
  switch (g.block_n) {
    case 32: return launch_igemm_bn<32, OUT_BF16, true>(g, grid, st);
    case 64: return launch_igemm_bn<64, OUT_BF16, true>(g, grid, st);
    case 128: return launch_igemm_bn<128, OUT_BF16, true>(g, grid, st);
    default: return fail(PETSYN_EINVAL, "unsupported BLOCK_N %d for output statistics", g.block_n);
  }
}

// run one gather-form GEMM: (split-K: zero the fp32 workspace, reduce into it, convert) or a direct launch
static int32_t run_side(GemmSide& g, const ViewSpec& vc, void* c, const float* bias, int act, float slope, int batch,
                        cudaStream_t st, const petsyn_conv_epilogue* ep = nullptr) {
  if (g.ksplit > 1) {
    const size_t bytes = (size_t)g.out_rows_full * g.R * sizeof(float) * g.ksplit;     // one partial image per K split
    if (g.workspace == nullptr || g.workspace_bytes < bytes)
      return fail(PETSYN_ENOMEM, "split-K needs a %zu-byte workspace (petsyn_conv_set_workspace)", bytes);
    // (no clearing: every split of every sub-problem owns at least one K step, so each partial image is stored in full)
    int32_t rc = launch_igemm(g, batch, st);
    if (rc) return rc;
    if (ep != nullptr) {             // the finish pass also sums what it stores (statistics for the consuming normalisation)
      const int64_t rps = g.out_rows_full / batch;
      const int rpp = 256 / (g.R / 8);
      // two CTAs per SM: every CTA ends with 2 C double atomics, so more (smaller) CTAs cost more than they spread (512 channels,
      // 2304 rows: +10 us over the plain finish pass with 592 CTAs, +4 us with 296)
      static const int waves = getenv("PETSYN_FINISH_WAVES") ? std::max(1, atoi(getenv("PETSYN_FINISH_WAVES"))) : 2;
      const int bx = (int)std::max<int64_t>(1, std::min<int64_t>((rps + rpp - 1) / rpp, std::max(1, 148 * waves / batch)));
      PETSYN_CHECK_CUDA(launch_pdl(splitk_finish_stats_kernel, dim3(bx, batch), dim3(256), 0, st,
                                   reinterpret_cast<const float*>(g.workspace), g.ksplit, reinterpret_cast<__nv_bfloat16*>(c), rps,
                                   g.R, vc.cstride, vc.coff, bias, act, slope, ep->stats1, ep->stats1_c, ep->stats1_coff,
                                   ep->stats2, ep->stats2_c, ep->stats2_coff));
      return check_launch("splitk_finish_stats_kernel");
    }
    const int64_t total = g.out_rows_full * (g.R / 8);
    const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148 * 8));
    PETSYN_CHECK_CUDA(launch_pdl(splitk_finish_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const float*>(g.workspace), g.ksplit,
                                                 reinterpret_cast<__nv_bfloat16*>(c), g.out_rows_full, g.R, vc.cstride,
                                                 vc.coff, bias, act, slope, g.accumulate ? 1 : 0));
    return check_launch("splitk_finish_kernel");
  }
  return launch_igemm(g, batch, st);
}

// (re)build the TMA descriptors of a gather-form GEMM for the given base pointers
static int32_t bind_side(GemmSide& g, const ViewSpec& va, const ViewSpec& vc, const void* a, const void* b,
                         const void* c, const float* bias, int act, float slope) {
  const bool split = g.ksplit > 1;
  if (g.key_a == a && g.key_b == b && g.key_c == c && g.key_w == g.workspace) {
    g.params.bias = split ? nullptr : bias;
    return PETSYN_OK;
  }
  IgemmParams& p = g.params;
  memset(&p, 0, sizeof(p));
  const int n_a = g.prog.a_phased ? 8 : 1;
  const int n_c = g.prog.out_phased ? 8 : 1;
  for (int i = 0; i < n_a; ++i) {
    int32_t rc = view_map(&p.a_maps[i], a, va, g.prog.a_phased, i, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, g.kch, g.box_w,
                          g.box_h, g.box_d, 128);
    if (rc) return rc;
  }
  const bool c_f32 = g.out_fp32 || split;
  const int c_esz = c_f32 ? 4 : 2;
  const bool c_swizzled = (g.block_n * c_esz) % 128 == 0;
  ViewSpec vws = vc;               // split-K: every split stores its own dense fp32 partial image of the output tensor
  vws.cstride = g.R; vws.coff = 0; // (slot s of sample n = "sample" s * N + n of the workspace view)
  if (split) vws.N = vc.N * g.ksplit;
  if (split && g.workspace == nullptr) return fail(PETSYN_ENOMEM, "split-K workspace not set (petsyn_conv_set_workspace)");
  const int chunk_c = c_swizzled ? 128 / c_esz : g.block_n;
  const int c_swz = c_swizzled ? 128 : 0;
  for (int i = 0; i < n_c; ++i) {
    int32_t rc = view_map(&p.c_maps[i], split ? g.workspace : c, split ? vws : vc, g.prog.out_phased, i, c_esz,
                          c_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, chunk_c,
                          g.box_w, g.box_h, g.box_d, c_swz);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.prog.max_taps * g.kc_pad, (uint64_t)g.subs.size() * g.R};
    uint64_t strides[1] = {dims[0] * 2};
    uint32_t box[2] = {(uint32_t)g.kch, (uint32_t)g.block_n};
    int32_t rc = encode_tmap(&p.b_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, b, dims, strides, box, 128);
    if (rc) return rc;
  }
  for (size_t i = 0; i < g.subs.size(); ++i) p.subs[i] = g.subs[i];
  p.taps = g.d_taps;
  p.bias = split ? nullptr : bias;
  p.ksplit = g.ksplit;
  p.tiles_w = (g.out_w + g.box_w - 1) / g.box_w;
  p.tiles_h = (g.out_h + g.box_h - 1) / g.box_h;
  p.tiles_d = (g.out_d + g.box_d - 1) / g.box_d;
  p.batch = va.N;
  p.box_w = g.box_w; p.box_h = g.box_h; p.box_d = g.box_d;
  p.a_stage_bytes = g.box_w * g.box_h * g.box_d * g.kch * 2;
  p.kc_chunks = g.kc_pad / g.kch;
  p.kc_pad = g.kc_pad;
  p.rows = g.R;
  p.epi_act = split ? PETSYN_ACT_NONE : act;
  p.epi_slope = slope;
  g.key_a = a; g.key_b = b; g.key_c = c; g.key_w = g.workspace;
  return PETSYN_OK;
}

static size_t packed_bytes(const GemmSide& g) {
  return (size_t)g.subs.size() * g.R * g.prog.max_taps * g.kc_pad * 2;
}

static PackImage make_image(const GemmSide& g, void* out) {
  PackImage im;
  im.out = reinterpret_cast<__nv_bfloat16*>(out);
  im.tbl = g.d_src;
  im.nsubs = (int)g.subs.size();
  im.max_taps = g.prog.max_taps;
  im.rows = g.R;
  im.kc_pad = g.kc_pad;
  return im;
}

static int max_nsrc(const GemmSide& g) {
  size_t m = 1;
  for (auto& sp : g.prog.subs)
    for (auto& t : sp) m = std::max(m, t.src.size());
  return (int)m;
}

template <int K3, bool T, int NS>
static int32_t launch_pack_t(const float* w, int A, int B, const PackImage& im, cudaStream_t st) {
  constexpr int TA = T ? 64 : 4, TB = T ? 4 : 64;
  constexpr int K3P = (K3 & 1) ? K3 + 2 : K3 + 1;
  constexpr size_t smem = (size_t)TA * ((TB * K3P) | 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    PETSYN_CHECK_CUDA(
        cudaFuncSetAttribute(pack_weights_kernel<K3, T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  // the tile grid also covers the zero padding of the image's column dimension
  const int amax = T ? std::max(A, im.kc_pad) : A, bmax = T ? B : std::max(B, im.kc_pad);
  dim3 grid((unsigned)((amax + TA - 1) / TA), (unsigned)((bmax + TB - 1) / TB));
  pack_weights_kernel<K3, T, NS><<<grid, 256, smem, st>>>(w, A, B, im);
  return check_launch("pack_weights_kernel");
}

static int32_t launch_pack(int k3, bool transposed, int ns, const float* w, int A, int B, const PackImage& im,
                           cudaStream_t st) {
#define PETSYN_PACK_CASE(K, NS)                                                           \
  if (k3 == K && ns <= NS)                                                                \
    return transposed ? launch_pack_t<K, true, NS>(w, A, B, im, st) : launch_pack_t<K, false, NS>(w, A, B, im, st);
  PETSYN_PACK_CASE(1, 1)
  PETSYN_PACK_CASE(8, 1)
  PETSYN_PACK_CASE(27, 1)
  PETSYN_PACK_CASE(27, 8)
  PETSYN_PACK_CASE(64, 1)
#undef PETSYN_PACK_CASE
  return fail(PETSYN_EINVAL, "unsupported kernel volume %d / merge factor %d", k3, ns);
}

// ---- batched packing: jobs grouped by kernel instantiation
struct PackClass { int k3; bool transposed; int ns; };
static const PackClass kPackClasses[] = {{1, false, 1}, {1, true, 1}, {8, false, 1}, {8, true, 1}, {27, false, 1},
                                          {27, true, 1}, {27, false, 8}, {27, true, 8}, {64, false, 1}, {64, true, 1}};
constexpr int kNumPackClasses = sizeof(kPackClasses) / sizeof(kPackClasses[0]);

static int pack_class_of(int k3, bool transposed, int ns) {
  for (int i = 0; i < kNumPackClasses; ++i)
    if (kPackClasses[i].k3 == k3 && kPackClasses[i].transposed == transposed && ns <= kPackClasses[i].ns) return i;
  return -1;
}

template <int K3, bool T, int NS>
static int32_t launch_pack_multi_t(const PackJob* jobs, int njobs, int tiles, cudaStream_t st) {
  constexpr int TA = T ? 64 : 4, TB = T ? 4 : 64;
  constexpr int K3P = (K3 & 1) ? K3 + 2 : K3 + 1;
  constexpr size_t smem = (size_t)TA * ((TB * K3P) | 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(pack_weights_multi_kernel<K3, T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    attr = true;
  }
  PETSYN_CHECK_CUDA(launch_pdl(pack_weights_multi_kernel<K3, T, NS>, dim3(tiles), dim3(256), smem, st, jobs, njobs));
  return check_launch("pack_weights_multi_kernel");
}

static int32_t launch_pack_multi(int cls, const PackJob* jobs, int njobs, int tiles, cudaStream_t st) {
  switch (cls) {
    case 0: return launch_pack_multi_t<1, false, 1>(jobs, njobs, tiles, st);
    case 1: return launch_pack_multi_t<1, true, 1>(jobs, njobs, tiles, st);
    case 2: return launch_pack_multi_t<8, false, 1>(jobs, njobs, tiles, st);
    case 3: return launch_pack_multi_t<8, true, 1>(jobs, njobs, tiles, st);
    case 4: return launch_pack_multi_t<27, false, 1>(jobs, njobs, tiles, st);
    case 5: return launch_pack_multi_t<27, true, 1>(jobs, njobs, tiles, st);
    case 6: return launch_pack_multi_t<27, false, 8>(jobs, njobs, tiles, st);
    case 7: return launch_pack_multi_t<27, true, 8>(jobs, njobs, tiles, st);
    case 8: return launch_pack_multi_t<64, false, 1>(jobs, njobs, tiles, st);
    case 9: return launch_pack_multi_t<64, true, 1>(jobs, njobs, tiles, st);
    default: return fail(PETSYN_EINVAL, "bad pack class %d", cls);
  }
}

struct BiasCast { const double* acc = nullptr; float* out = nullptr; int n = 0; };

template <int K3, bool T, int NS>
static int32_t launch_unpack_t(const float* scratch, float* dw, const InvEntryDev* inv, int A, int B, const GemmSide& f,
                               int accumulate, int nsplit, const BiasCast& bc, cudaStream_t st) {
  constexpr int TA = T ? 64 : 4, TB = T ? 4 : 64;
  const size_t smem = ((size_t)f.subs.size() * f.prog.max_taps + 1) * ((TA * (TB + 1)) | 1) * sizeof(float);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    PETSYN_CHECK_CUDA(
        cudaFuncSetAttribute(unpack_wgrad_kernel<K3, T, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((unsigned)((A + TA - 1) / TA), (unsigned)((B + TB - 1) / TB));
  PETSYN_CHECK_CUDA(launch_pdl(unpack_wgrad_kernel<K3, T, NS>, dim3(grid), dim3(256), smem, st, scratch, dw, inv, A, B, (int)f.subs.size(), f.prog.max_taps,
                                                          f.R, f.kc_pad, accumulate, nsplit, bc.acc, bc.out, bc.n));
  return check_launch("unpack_wgrad_kernel");
}

static int32_t launch_unpack(int k3, bool transposed, int ns, const float* scratch, float* dw, const InvEntryDev* inv,
                             int A, int B, const GemmSide& f, int accumulate, int nsplit, const BiasCast& bc,
                             cudaStream_t st) {
#define PETSYN_UNPACK_CASE(K, NS)                                                                          \
  if (k3 == K && ns <= NS)                                                                                 \
    return transposed ? launch_unpack_t<K, true, NS>(scratch, dw, inv, A, B, f, accumulate, nsplit, bc, st) \
                      : launch_unpack_t<K, false, NS>(scratch, dw, inv, A, B, f, accumulate, nsplit, bc, st);
  PETSYN_UNPACK_CASE(1, 1)
  PETSYN_UNPACK_CASE(8, 1)
  PETSYN_UNPACK_CASE(27, 1)
  PETSYN_UNPACK_CASE(27, 8)
  PETSYN_UNPACK_CASE(64, 1)
#undef PETSYN_UNPACK_CASE
  return fail(PETSYN_EINVAL, "unsupported kernel volume %d / merge factor %d", k3, ns);
}

}  // namespace petsyn

using namespace petsyn;

extern "C" {

int32_t petsyn_conv_plan_create(const petsyn_conv_desc* d, petsyn_conv_plan** out) {
  PETSYN_REQUIRE(d != nullptr && out != nullptr, "null argument");
  PETSYN_REQUIRE(d->n > 0 && d->d > 0 && d->h > 0 && d->w > 0, "non-positive tensor dims");
  PETSYN_REQUIRE(d->cin > 0 && d->cout > 0 && d->cin % 8 == 0 && d->cout % 8 == 0,
                 "cin (%d) and cout (%d) must be positive multiples of 8", d->cin, d->cout);
  PETSYN_REQUIRE(d->x_cstride % 8 == 0 && d->y_cstride % 8 == 0 && d->dx_cstride % 8 == 0 && d->dy_cstride % 8 == 0 &&
                     d->x_coff % 8 == 0 && d->y_coff % 8 == 0 && d->dx_coff % 8 == 0 && d->dy_coff % 8 == 0,
                 "channel pitches/offsets must be multiples of 8 (16-byte TMA alignment)");
  PETSYN_REQUIRE(d->x_coff + d->cin <= d->x_cstride && d->dx_coff + d->cin <= d->dx_cstride &&
                     d->y_coff + d->cout <= d->y_cstride && d->dy_coff + d->cout <= d->dy_cstride,
                 "channel slice exceeds the buffer's channel pitch");
  auto* pl = new petsyn_conv_plan();
  pl->desc = *d;
  const int k = d->ksize, s = d->stride, p = d->pad;
  AxisProg fwd, bwd;
  int od, oh, ow;
  if (d->op == PETSYN_OP_CONV) {
    if (!(s == 1 || s == 2) || k < 1 || k > 4 || p < 0 || p >= k) {
      delete pl;
      return fail(PETSYN_EINVAL, "Conv3d: unsupported kernel/stride/pad %d/%d/%d", k, s, p);
    }
    od = (d->d + 2 * p - k) / s + 1; oh = (d->h + 2 * p - k) / s + 1; ow = (d->w + 2 * p - k) / s + 1;
    if (od < 1 || oh < 1 || ow < 1) {
      delete pl;
      return fail(PETSYN_EINVAL, "Conv3d: empty output for input %dx%dx%d (k=%d s=%d p=%d)", d->d, d->h, d->w, k, s, p);
    }
    if (s == 2 && (od != (d->d + 1) / 2 || oh != (d->h + 1) / 2 || ow != (d->w + 1) / 2)) {
      delete pl;
      return fail(PETSYN_EINVAL, "stride-2 Conv3d must map n -> ceil(n/2) (k=%d p=%d on %dx%dx%d)", k, p, d->d, d->h,
                  d->w);
    }
    fwd = (s == 1) ? axis_s1(k, p, +1) : axis_s2_gather(k, p);
    bwd = (s == 1) ? axis_s1(k, p, -1) : axis_s2_scatter(k, p);
  } else if (d->op == PETSYN_OP_UPCONV) {
    if (!(k == 3 && s == 1 && p == 1)) {
      delete pl;
      return fail(PETSYN_EINVAL, "Upsample+Conv3d supports k3 s1 p1 only");
    }
    od = 2 * d->d; oh = 2 * d->h; ow = 2 * d->w;
    fwd = axis_up_fwd(k, p);
    bwd = axis_up_bwd(k, p);
  } else if (d->op == PETSYN_OP_CONVT) {
    if (!(k == 4 && s == 2 && p == 1)) {
      delete pl;
      return fail(PETSYN_EINVAL, "ConvTranspose3d supports k4 s2 p1 only");
    }
    od = 2 * d->d; oh = 2 * d->h; ow = 2 * d->w;
    fwd = axis_s2_scatter(k, p);
    bwd = axis_s2_gather(k, p);
  } else {
    delete pl;
    return fail(PETSYN_EINVAL, "unknown op %d", d->op);
  }
  pl->k3 = k * k * k;
  pl->od = od; pl->oh = oh; pl->ow = ow;
  pl->vx = {d->cin, d->x_cstride, d->x_coff, d->w, d->h, d->d, d->n};
  pl->vdx = {d->cin, d->dx_cstride, d->dx_coff, d->w, d->h, d->d, d->n};
  pl->vy = {d->cout, d->y_cstride, d->y_coff, ow, oh, od, d->n};
  pl->vdy = {d->cout, d->dy_cstride, d->dy_coff, ow, oh, od, d->n};

  GemmSide& f = pl->fprop;
  f.prog = make_program(fwd, k);
  f.R = d->cout; f.Kc = d->cin;
  f.swap = (d->op == PETSYN_OP_CONVT);
  f.out_fp32 = d->y_fp32 != 0;
  f.out_w = f.prog.out_phased ? ow / 2 : ow;
  f.out_h = f.prog.out_phased ? oh / 2 : oh;
  f.out_d = f.prog.out_phased ? od / 2 : od;
  f.out_rows_full = (int64_t)d->n * od * oh * ow;
  const int slab_ok = (d->op == PETSYN_OP_CONV && k == 3 && s == 1 && p == 1) ? 1
                      : (d->op == PETSYN_OP_CONV && k == 1 && s == 1 && p == 0) ? 2 : 0;
  int32_t rc = finish_side(f, d->n, slab_ok);
  GemmSide& g = pl->dgrad;
  if (!rc) {
    g.prog = make_program(bwd, k);
    g.R = d->cin; g.Kc = d->cout;
    g.swap = (d->op != PETSYN_OP_CONVT);
    g.out_w = g.prog.out_phased ? (d->w + 1) / 2 : d->w;      // phase 0 of an odd extent has one more element
    g.out_h = g.prog.out_phased ? (d->h + 1) / 2 : d->h;
    g.out_d = g.prog.out_phased ? (d->d + 1) / 2 : d->d;
    g.out_rows_full = (int64_t)d->n * d->d * d->h * d->w;
    rc = finish_side(g, d->n, slab_ok);
  }
  if (!rc) {
    // wgrad: inverse table (original tap -> packed (sub, tap) slots) from the fprop program
    std::vector<InvEntryDev> inv(pl->k3);
    memset(inv.data(), 0, inv.size() * sizeof(InvEntryDev));
    for (size_t s2 = 0; s2 < f.prog.subs.size() && !rc; ++s2)
      for (size_t t = 0; t < f.prog.subs[s2].size(); ++t)
        for (int src : f.prog.subs[s2][t].src) {
          InvEntryDev& e = inv[src];
          if (e.n >= 8) { rc = fail(PETSYN_EINVAL, "tap appears in more than 8 merged slots"); break; }
          e.sub[e.n] = (int)s2; e.tap[e.n] = (int)t; ++e.n;
          pl->inv_max = std::max(pl->inv_max, e.n);
        }
    if (!rc) rc = upload(inv.data(), inv.size() * sizeof(InvEntryDev), (void**)&pl->d_inv);
    choose_box(f.out_w, f.out_h, f.out_d, 64, true, &pl->wg_box_w, &pl->wg_box_h, &pl->wg_box_d);
    pl->wg_block_n = (d->cin % 128 == 0 || d->cin > 64) ? 128 : 64;
    // split the voxel reduction so that the grid fills the machine (~2 waves of 148 SMs x 2 CTAs)
    int64_t total_taps = 0;
    for (auto& s2 : f.prog.subs) total_taps += (int64_t)s2.size();
    const int64_t base_ctas = total_taps * ((d->cout + 127) / 128) * ((d->cin + pl->wg_block_n - 1) / pl->wg_block_n);
    const int64_t nboxes = (int64_t)((f.out_w + pl->wg_box_w - 1) / pl->wg_box_w) *
                           ((f.out_h + pl->wg_box_h - 1) / pl->wg_box_h) *
                           ((f.out_d + pl->wg_box_d - 1) / pl->wg_box_d) * d->n;
    int64_t ks = (148 * 4 + base_ctas - 1) / base_ctas;
    ks = std::max<int64_t>(1, std::min<int64_t>(ks, std::max<int64_t>(1, nboxes / 8)));
    pl->wg_ksplit = (int)ks;
    if (d->op != PETSYN_OP_CONVT && d->cin % 16 == 0 && d->cin <= 64 && d->cout <= 64) {
      pl->wg_small = true;
      pl->wg_tpm = 128 / d->cin;
      pl->wg_mtiles = (f.prog.max_taps + pl->wg_tpm - 1) / pl->wg_tpm;
      pl->wg_npad = (d->cout + 15) / 16 * 16;
      const int64_t ctas = (int64_t)pl->wg_mtiles * (int64_t)f.prog.subs.size();
      int64_t ks2 = (148 * 3 + ctas - 1) / ctas;
      ks2 = std::max<int64_t>(1, std::min<int64_t>(ks2, std::max<int64_t>(1, nboxes / 8)));
      pl->wg_ksplit = (int)ks2;
    }
    pl->wg_slab_halo = slab_ok == 1 ? 1 : 0;
    if (slab_ok && d->cin % 16 == 0 && d->cin <= 96 && d->cout % 16 == 0 && d->cout <= 64 &&
        (int64_t)d->w * d->h * d->d * d->n >= 32768) {
      pl->wg_slab = true;
      pl->wg_small = false;
    }
  }
  if (rc) {
    petsyn_conv_plan_destroy(pl);
    return rc;
  }
  *out = pl;
  return PETSYN_OK;
}

void petsyn_conv_plan_destroy(petsyn_conv_plan* pl) {
  if (!pl) return;
  cudaFree(pl->fprop.d_taps); cudaFree(pl->fprop.d_src);
  cudaFree(pl->dgrad.d_taps); cudaFree(pl->dgrad.d_src);
  cudaFree(pl->d_inv);
  delete pl;
}

int32_t petsyn_conv_out_dims(const petsyn_conv_plan* pl, int32_t* od, int32_t* oh, int32_t* ow) {
  PETSYN_REQUIRE(pl != nullptr, "null plan");
  *od = pl->od; *oh = pl->oh; *ow = pl->ow;
  return PETSYN_OK;
}

int32_t petsyn_conv_flops(const petsyn_conv_plan* pl, double* algorithmic, double* executed) {
  PETSYN_REQUIRE(pl != nullptr, "null plan");
  const petsyn_conv_desc& d = pl->desc;
  const double mout = (double)d.n * pl->od * pl->oh * pl->ow;
  const double min_ = (double)d.n * d.d * d.h * d.w;
  if (algorithmic)
    *algorithmic = (d.op == PETSYN_OP_CONVT) ? 2.0 * min_ * d.cin * d.cout * pl->k3 : 2.0 * mout * d.cin * d.cout * pl->k3;
  if (executed) {
    double taps = 0;
    for (auto& s : pl->fprop.prog.subs) taps += (double)s.size();
    const double vox_per_sub = (double)d.n * pl->fprop.out_d * pl->fprop.out_h * pl->fprop.out_w;
    *executed = 2.0 * vox_per_sub * taps * d.cin * d.cout;
  }
  return PETSYN_OK;
}

int32_t petsyn_conv_kernel_path(const petsyn_conv_plan* pl, int32_t pass) {
  if (!pl) return -1;
  switch (pass) {
    case 0: return pl->fprop.slab ? 1 : 0;
    case 1: return pl->dgrad.slab ? 1 : 0;
    case 2: return pl->wg_slab ? 1 : (pl->wg_small ? 2 : 0);
    default: return -1;
  }
}

size_t petsyn_conv_packed_fprop_bytes(const petsyn_conv_plan* pl) { return pl ? packed_bytes(pl->fprop) : 0; }
size_t petsyn_conv_packed_dgrad_bytes(const petsyn_conv_plan* pl) { return pl ? packed_bytes(pl->dgrad) : 0; }
size_t petsyn_conv_wgrad_scratch_bytes(const petsyn_conv_plan* pl) {
  if (!pl) return 0;
  if (pl->wg_slab) {
    const int atoms = pl->desc.cin / 16, groups = (atoms + 2) / 3, apg = (atoms + groups - 1) / groups;
    const int nacc = pl->wg_slab_halo ? 3 : 1;
    // one image per persistent CTA (at most 2 CTAs per SM share the (co atom, channel group) grid); the unpack kernel adds
    // them in CTA order
    const int images = std::max(1, 296 / ((pl->desc.cout / 16) * groups));
    return (size_t)groups * (pl->desc.cout / 16) * nacc * 48 * (nacc * apg * 16) * sizeof(float) * images;
  }
  // gather-form kernels: one fp32 partial image per K split (summed in split order by the unpack kernel)
  if (pl->wg_small) return (size_t)pl->fprop.subs.size() * pl->wg_mtiles * 128 * pl->wg_npad * sizeof(float) * pl->wg_ksplit;
  return packed_bytes(pl->fprop) * 2 * pl->wg_ksplit;
}

size_t petsyn_conv_workspace_bytes(const petsyn_conv_plan* pl) {
  if (!pl) return 0;
  size_t b = 0;
  for (const GemmSide* g : {&pl->fprop, &pl->dgrad})
    if (g->ksplit > 1) b = std::max(b, (size_t)g->out_rows_full * g->R * sizeof(float) * g->ksplit);
  return b;
}

int32_t petsyn_conv_set_workspace(petsyn_conv_plan* pl, void* workspace, size_t bytes) {
  PETSYN_REQUIRE(pl != nullptr, "null plan");
  PETSYN_REQUIRE(bytes >= petsyn_conv_workspace_bytes(pl), "workspace too small (%zu < %zu)", bytes,
                 petsyn_conv_workspace_bytes(pl));
  for (GemmSide* g : {&pl->fprop, &pl->dgrad}) {
    g->workspace = workspace;
    g->workspace_bytes = bytes;
  }
  return PETSYN_OK;
}

int32_t petsyn_conv_pack_weights(petsyn_conv_plan* pl, const float* w, void* packed_fprop, void* packed_dgrad,
                                 void* stream) {
  PETSYN_REQUIRE(pl != nullptr && w != nullptr, "null argument");
  // W[a][b][k]: (a, b) = (cout, cin) for Conv/UpConv, (cin, cout) for ConvTranspose.  An image whose rows are W's dim 1
  // is "transposed".
  const bool convt = pl->desc.op == PETSYN_OP_CONVT;
  const int A = convt ? pl->desc.cin : pl->desc.cout, B = convt ? pl->desc.cout : pl->desc.cin;
  cudaStream_t st = as_stream(stream);
  int32_t rc = PETSYN_OK;
  if (packed_fprop)
    rc = launch_pack(pl->k3, /*transposed=*/convt, max_nsrc(pl->fprop), w, A, B, make_image(pl->fprop, packed_fprop), st);
  if (!rc && packed_dgrad)
    rc = launch_pack(pl->k3, /*transposed=*/!convt, max_nsrc(pl->dgrad), w, A, B, make_image(pl->dgrad, packed_dgrad), st);
  return rc;
}

int32_t petsyn_pack_batch_create(int32_t n, petsyn_conv_plan* const* plans, const float* const* weights,
                                 void* const* packed_fprop, void* const* packed_dgrad, petsyn_pack_batch** out) {
  PETSYN_REQUIRE(n > 0 && plans && weights && packed_fprop && packed_dgrad && out, "bad argument");
  std::vector<PackJob> jobs[kNumPackClasses];
  int tiles[kNumPackClasses] = {0};
  auto add = [&](const petsyn_conv_plan* pl, const GemmSide& g, bool transposed, const float* w, void* dst) -> int32_t {
    const bool convt = pl->desc.op == PETSYN_OP_CONVT;
    const int A = convt ? pl->desc.cin : pl->desc.cout, B = convt ? pl->desc.cout : pl->desc.cin;
    const int cls = pack_class_of(pl->k3, transposed, max_nsrc(g));
    if (cls < 0) return fail(PETSYN_EINVAL, "unsupported kernel volume %d / merge factor %d", pl->k3, max_nsrc(g));
    const int TA = transposed ? 64 : 4, TB = transposed ? 4 : 64;
    PackJob j;
    j.w = w; j.A = A; j.B = B;
    j.im = make_image(g, dst);
    const int amax = transposed ? std::max(A, j.im.kc_pad) : A, bmax = transposed ? B : std::max(B, j.im.kc_pad);
    j.tiles_x = (amax + TA - 1) / TA;
    const int ty = (bmax + TB - 1) / TB;
    j.tile_begin = tiles[cls];
    tiles[cls] += j.tiles_x * ty;
    jobs[cls].push_back(j);
    return PETSYN_OK;
  };
  for (int i = 0; i < n; ++i) {
    const petsyn_conv_plan* pl = plans[i];
    PETSYN_REQUIRE(pl && weights[i], "null plan / weight in batch entry %d", i);
    const bool convt = pl->desc.op == PETSYN_OP_CONVT;
    int32_t rc = PETSYN_OK;
    if (packed_fprop[i]) rc = add(pl, pl->fprop, convt, weights[i], packed_fprop[i]);
    if (!rc && packed_dgrad[i]) rc = add(pl, pl->dgrad, !convt, weights[i], packed_dgrad[i]);
    if (rc) return rc;
  }
  auto* b = new petsyn_pack_batch();
  for (int c = 0; c < kNumPackClasses; ++c) {
    if (jobs[c].empty()) continue;
    b->njobs[c] = (int)jobs[c].size();
    b->tiles[c] = tiles[c];
    int32_t rc = upload(jobs[c].data(), jobs[c].size() * sizeof(PackJob), (void**)&b->d_jobs[c]);
    if (rc) { petsyn_pack_batch_destroy(b); return rc; }
  }
  *out = b;
  return PETSYN_OK;
}

int32_t petsyn_pack_batch_run(petsyn_pack_batch* b, void* stream) {
  PETSYN_REQUIRE(b != nullptr, "null batch");
  for (int c = 0; c < kNumPackClasses; ++c) {
    if (b->njobs[c] == 0) continue;
    int32_t rc = launch_pack_multi(c, b->d_jobs[c], b->njobs[c], b->tiles[c], as_stream(stream));
    if (rc) return rc;
  }
  return PETSYN_OK;
}

void petsyn_pack_batch_destroy(petsyn_pack_batch* b) {
  if (!b) return;
  for (int c = 0; c < 16; ++c) cudaFree(b->d_jobs[c]);
  delete b;
}

int32_t petsyn_conv_fprop(petsyn_conv_plan* pl, const void* x, const void* packed, const float* bias, void* y,
                          void* stream) {
  PETSYN_REQUIRE(pl && x && packed && y, "null argument");
  if (pl->fprop.slab) {
    int32_t rs = bind_slab(pl->fprop, pl->vx, pl->vy, x, packed, y, bias, pl->desc.epi_act, pl->desc.epi_slope);
    if (!rs && epi_supported(pl->fprop, pl->desc.n) && pl->fprop.slab3)
      return run_slab_epi(pl->fprop, pl->vx, pl->vy, x, packed, y, nullptr, false, as_stream(stream));
    return rs ? rs : run_slab(pl->fprop, as_stream(stream));
  }
  int32_t rc = bind_side(pl->fprop, pl->vx, pl->vy, x, packed, y, bias, pl->desc.epi_act, pl->desc.epi_slope);
  if (rc) return rc;
  return run_side(pl->fprop, pl->vy, y, bias, pl->desc.epi_act, pl->desc.epi_slope, pl->desc.n, as_stream(stream));
}

int32_t petsyn_conv_dgrad(petsyn_conv_plan* pl, const void* dy, const void* packed, void* dx, void* stream) {
  PETSYN_REQUIRE(pl && dy && packed && dx, "null argument");
  if (pl->dgrad.slab) {
    int32_t rs = bind_slab(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, PETSYN_ACT_NONE, 0.f);
    if (!rs && epi_supported(pl->dgrad, pl->desc.n) && pl->dgrad.slab3)
      return run_slab_epi(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, true, as_stream(stream));
    return rs ? rs : run_slab(pl->dgrad, as_stream(stream));
  }
  int32_t rc = bind_side(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, PETSYN_ACT_NONE, 0.f);
  if (rc) return rc;
  return run_side(pl->dgrad, pl->vdx, dx, nullptr, PETSYN_ACT_NONE, 0.f, pl->desc.n, as_stream(stream));
}

int32_t petsyn_conv_epilogue_supported(const petsyn_conv_plan* pl, int32_t pass) {
  if (!pl || pass < 0 || pass > 1) return 0;
  if (epi_supported(pass == 0 ? pl->fprop : pl->dgrad, pl->desc.n)) return 1;
  return (pass == 0 && igemm_stats_supported(pl->fprop, pl->desc.n)) ? 2 : 0;   // 2: statistics targets only
}

int32_t petsyn_conv_fprop_epi(petsyn_conv_plan* pl, const void* x, const void* packed, const float* bias, void* y,
                              const petsyn_conv_epilogue* epi, void* stream) {
  PETSYN_REQUIRE(pl && x && packed && y && epi, "null argument");
  if (!pl->fprop.slab && igemm_stats_supported(pl->fprop, pl->desc.n)) {
    PETSYN_REQUIRE(epi->side == nullptr && epi->bsums == nullptr && (epi->stats1 != nullptr || epi->stats2 != nullptr),
                   "the gather-form kernel's epilogue takes statistics targets only");
    int32_t rc = bind_side(pl->fprop, pl->vx, pl->vy, x, packed, y, bias, pl->desc.epi_act, pl->desc.epi_slope);
    if (rc) return rc;
    PETSYN_REQUIRE((epi->stats1 == nullptr || epi->stats1_coff + pl->fprop.R <= epi->stats1_c) &&
                   (epi->stats2 == nullptr || epi->stats2_coff + pl->fprop.R <= epi->stats2_c),
                   "statistics target narrower than the output channels");
    if (pl->fprop.ksplit > 1)
      return run_side(pl->fprop, pl->vy, y, bias, pl->desc.epi_act, pl->desc.epi_slope, pl->desc.n, as_stream(stream), epi);
    IgemmParams& p = pl->fprop.params;
    p.st1 = epi->stats1; p.st1_c = epi->stats1_c; p.st1_off = epi->stats1_coff;
    p.st2 = epi->stats2; p.st2_c = epi->stats2_c; p.st2_off = epi->stats2_coff;
    p.ext_w = pl->vy.W; p.ext_h = pl->vy.H; p.ext_d = pl->vy.D;
    PETSYN_REQUIRE((p.st1 == nullptr || p.st1_off + pl->fprop.R <= p.st1_c) && (p.st2 == nullptr || p.st2_off + pl->fprop.R <= p.st2_c),
                   "statistics target narrower than the output channels");
    rc = launch_igemm_stats(pl->fprop, pl->desc.n, as_stream(stream));
    p.st1 = p.st2 = nullptr;
    return rc;
  }
  PETSYN_REQUIRE(epi_supported(pl->fprop, pl->desc.n), "this plan's forward pass has no fused epilogue");
  int32_t rs = bind_slab(pl->fprop, pl->vx, pl->vy, x, packed, y, bias, pl->desc.epi_act, pl->desc.epi_slope);
  return rs ? rs : run_slab_epi(pl->fprop, pl->vx, pl->vy, x, packed, y, epi, false, as_stream(stream));
}

int32_t petsyn_conv_dgrad_epi(petsyn_conv_plan* pl, const void* dy, const void* packed, void* dx,
                              const petsyn_conv_epilogue* epi, void* stream) {
  PETSYN_REQUIRE(pl && dy && packed && dx && epi, "null argument");
  PETSYN_REQUIRE(epi_supported(pl->dgrad, pl->desc.n), "this plan's data-gradient pass has no fused epilogue");
  int32_t rs = bind_slab(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, PETSYN_ACT_NONE, 0.f);
  return rs ? rs : run_slab_epi(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, epi, true, as_stream(stream));
}

int32_t petsyn_conv_dgrad_accumulate(petsyn_conv_plan* pl, const void* dy, const void* packed, void* dx, void* stream) {
  PETSYN_REQUIRE(pl && dy && packed && dx, "null argument");
  pl->dgrad.accumulate = true;
  if (pl->dgrad.slab) {
    int32_t rs = bind_slab(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, PETSYN_ACT_NONE, 0.f);
    if (!rs) rs = (epi_supported(pl->dgrad, pl->desc.n) && pl->dgrad.slab3)
                      ? run_slab_epi(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, true, as_stream(stream))
                      : run_slab(pl->dgrad, as_stream(stream));
    pl->dgrad.accumulate = false;
    return rs;
  }
  int32_t rc = bind_side(pl->dgrad, pl->vdy, pl->vdx, dy, packed, dx, nullptr, PETSYN_ACT_NONE, 0.f);
  if (!rc) rc = run_side(pl->dgrad, pl->vdx, dx, nullptr, PETSYN_ACT_NONE, 0.f, pl->desc.n, as_stream(stream));
  pl->dgrad.accumulate = false;
  return rc;
}

int32_t petsyn_conv_wgrad(petsyn_conv_plan* pl, const void* x, const void* dy, void* scratch, float* dw,
                          int32_t accumulate, void* stream) {
  return petsyn_conv_wgrad_bias(pl, x, dy, scratch, dw, accumulate, nullptr, nullptr, 0, stream);
}

int32_t petsyn_conv_wgrad_bias(petsyn_conv_plan* pl, const void* x, const void* dy, void* scratch, float* dw,
                               int32_t accumulate, const double* dbias_acc, float* dbias, int32_t nbias, void* stream) {
  PETSYN_REQUIRE(pl && x && dy && scratch && dw, "null argument");
  PETSYN_REQUIRE(dbias == nullptr || (dbias_acc != nullptr && nbias > 0), "bias gradient needs its accumulator and length");
  BiasCast bc;
  bc.acc = dbias_acc; bc.out = dbias; bc.n = nbias;
  GemmSide& f = pl->fprop;
  cudaStream_t st = as_stream(stream);
  if (pl->wg_slab) {
    SlabWgradParams& q = pl->wgl_params;
    const int atoms_total = pl->desc.cin / 16, co_atoms = pl->desc.cout / 16;
    const int groups = (atoms_total + 2) / 3;              // <= 3 atoms (48 channels) per CTA: 3 accumulators x 144 TMEM columns
    const int atoms = (atoms_total + groups - 1) / groups; // atoms per group; a short last group reads zero-filled atoms
    const int hh = pl->wg_slab_halo, nacc = hh ? 3 : 1;
    const int ncols = nacc * atoms * 16;
    if (!(pl->wg_key_x == x && pl->wg_key_g == dy && pl->wg_key_s == scratch)) {
      memset(&q, 0, sizeof(q));
      int32_t rc = slab_view_map(&q.x_map, x, pl->vx, atoms_total, kWgW, kWgH + 2 * hh, atoms);
      if (rc) return rc;
      {
        const ViewSpec& v = pl->vdy;
        const uint64_t cs = (uint64_t)v.cstride * 2;
        const uint8_t* b = reinterpret_cast<const uint8_t*>(dy) + (uint64_t)v.coff * 2;
        uint64_t dims[4] = {(uint64_t)v.C, (uint64_t)v.W, (uint64_t)v.H, (uint64_t)v.D * v.N};
        uint64_t strides[3] = {cs, cs * v.W, cs * v.W * v.H};
        uint32_t box[4] = {16u, (uint32_t)(kWgW + 2), (uint32_t)kWgH, 1u};
        rc = encode_tmap(&q.g_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, b, dims, strides, box, 32);
        if (rc) return rc;
      }
      q.scratch = reinterpret_cast<float*>(scratch);
      q.atoms = atoms; q.co_atoms = co_atoms;
      q.W = pl->desc.w; q.H = pl->desc.h; q.D = pl->desc.d; q.batch = pl->desc.n;
      q.tiles_w = (q.W + kWgW - 1) / kWgW;
      q.tiles_h = (q.H + kWgH - 1) / kWgH;
      q.halo = hh;
      q.m64 = getenv("PETSYN_WGRAD_M64") != nullptr ? atoi(getenv("PETSYN_WGRAD_M64")) : 1;   // M = 64 unless switched off
      q.xslab_tx = atoms * kWgW * (kWgH + 2 * hh) * 32;
      q.xslab_bytes = (q.xslab_tx + 1023) / 1024 * 1024;
      q.gslab_tx = (kWgW + 2) * kWgH * 32;
      q.gslab_bytes = (q.gslab_tx + 7 * 32 + 1023) / 1024 * 1024;   // M atoms 3..7 read past the last row: keep it in bounds
      q.acc_stride = ncols;
      q.tmem_cols = nacc * ncols <= 64 ? 64 : (nacc * ncols <= 128 ? 128 : (nacc * ncols <= 256 ? 256 : 512));
      q.gring = 4;
      const int cap = (q.tmem_cols <= 256 ? 110 : 200) * 1024;
      q.xring = slab_wgrad_smem_bytes(q.xslab_bytes, q.gslab_bytes, ncols, 8, q.gring) <= cap ? 8 : 4;
      pl->wg_slab_smem = slab_wgrad_smem_bytes(q.xslab_bytes, q.gslab_bytes, ncols, q.xring, q.gring);
      int occ = q.tmem_cols <= 256 ? std::max(1, std::min(2, (227 * 1024) / (pl->wg_slab_smem + 1024))) : 1;
      if (const char* e = getenv("PETSYN_WGRAD_OCC")) occ = std::max(1, std::min(occ, atoi(e)));       // tuning experiments only
      const int ctas = std::max(1, 148 * occ / (co_atoms * groups));
      slab_split(q.W, q.H, q.D, q.batch, kWgW, kWgH, ctas, &q.dchunk, &q.nchunks, &q.items);
      pl->wg_slab_grid = std::min(std::min(ctas, q.items), std::max(1, 296 / (co_atoms * groups)));
      q.image_floats = (int64_t)groups * co_atoms * nacc * 48 * ncols;
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(slab_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
      pl->wg_key_x = x; pl->wg_key_g = dy; pl->wg_key_s = scratch;
    }
    dim3 grid((unsigned)pl->wg_slab_grid, (unsigned)co_atoms, (unsigned)groups);
    PETSYN_CHECK_CUDA(launch_pdl(slab_wgrad_kernel, grid, dim3(192), pl->wg_slab_smem, st, q));
    int32_t rc = check_launch("slab_wgrad_kernel");
    if (rc) return rc;
    rc = reduce_images(reinterpret_cast<float*>(scratch), pl->wg_slab_grid, q.image_floats, st);
    if (rc) return rc;
    PETSYN_CHECK_CUDA(launch_pdl(slab_wgrad_unpack_kernel, dim3((unsigned)std::min<int64_t>((q.image_floats + 255) / 256, 148 * 8)), dim3(256), 0, st,
                                 reinterpret_cast<const float*>(scratch), dw, pl->desc.cout, pl->desc.cin, atoms, pl->k3, accumulate, 1,
        q.image_floats, bc.acc, bc.out, bc.n));
    return check_launch("slab_wgrad_unpack_kernel");
  }
  if (pl->wg_small) {
    WgradSmallParams& q = pl->wgs_params;
    if (!(pl->wg_key_x == x && pl->wg_key_g == dy && pl->wg_key_s == scratch)) {
      memset(&q, 0, sizeof(q));
      const int n_x = f.prog.a_phased ? 8 : 1;
      const int n_g = f.prog.out_phased ? 8 : 1;
      for (int i = 0; i < n_x; ++i) {
        int32_t rc = view_map(&q.x_maps[i], x, pl->vx, f.prog.a_phased, i, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 16,
                              pl->wg_box_w, pl->wg_box_h, pl->wg_box_d, 32);
        if (rc) return rc;
      }
      for (int i = 0; i < n_g; ++i) {
        int32_t rc = view_map(&q.g_maps[i], dy, pl->vdy, f.prog.out_phased, i, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 16,
                              pl->wg_box_w, pl->wg_box_h, pl->wg_box_d, 32);
        if (rc) return rc;
      }
      uint64_t dims[2] = {(uint64_t)pl->wg_npad, (uint64_t)f.subs.size() * pl->wg_mtiles * 128 * pl->wg_ksplit};
      uint64_t strides[1] = {dims[0] * 4};
      uint32_t box[2] = {(uint32_t)pl->wg_npad, 128u};
      int32_t rc = encode_tmap(&q.d_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, scratch, dims, strides, box, 0);
      if (rc) return rc;
      for (size_t i = 0; i < f.subs.size(); ++i) q.subs[i] = f.subs[i];
      q.taps = f.d_taps;
      q.tiles_w = (f.out_w + pl->wg_box_w - 1) / pl->wg_box_w;
      q.tiles_h = (f.out_h + pl->wg_box_h - 1) / pl->wg_box_h;
      q.tiles_d = (f.out_d + pl->wg_box_d - 1) / pl->wg_box_d;
      q.batch = pl->desc.n;
      q.box_w = pl->wg_box_w; q.box_h = pl->wg_box_h; q.box_d = pl->wg_box_d;
      q.cin = pl->desc.cin; q.cin_atoms = pl->desc.cin / 16;
      q.n_atoms = pl->wg_npad / 16;
      q.tpm = pl->wg_tpm; q.m_tiles = pl->wg_mtiles; q.ksplit = pl->wg_ksplit;
      q.image_rows = (int)f.subs.size() * pl->wg_mtiles * 128;
      pl->wg_key_x = x; pl->wg_key_g = dy; pl->wg_key_s = scratch;
    }
    PETSYN_CHECK_CUDA(cudaMemsetAsync(scratch, 0, petsyn_conv_wgrad_scratch_bytes(pl), st));
    constexpr int STAGES = 4;
    constexpr int smem = STAGES * (12 * 2048) + 1024 + 256;
    auto kern = wgrad_small_kernel<STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    dim3 grid((unsigned)pl->wg_mtiles, (unsigned)pl->wg_ksplit, (unsigned)f.subs.size());
    PETSYN_CHECK_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, q));
    int32_t rc = check_launch("wgrad_small_kernel");
    if (rc) return rc;
    const int total = pl->desc.cout * pl->desc.cin * pl->k3;
    const int64_t image = (int64_t)f.subs.size() * pl->wg_mtiles * 128 * pl->wg_npad;
    // few splits: added in split order by the unpack kernel itself; many: one coalesced pass first
    const int inline_split = pl->wg_ksplit <= 4 ? pl->wg_ksplit : 1;
    if (inline_split == 1) {
      rc = reduce_images(reinterpret_cast<float*>(scratch), pl->wg_ksplit, image, st);
      if (rc) return rc;
    }
    PETSYN_CHECK_CUDA(launch_pdl(unpack_wgrad_small_kernel, dim3(std::min((total + 255) / 256, 148 * 8)), dim3(256), 0, st, 
        reinterpret_cast<const float*>(scratch), dw, pl->d_inv, pl->desc.cout, pl->desc.cin, pl->k3, pl->wg_npad, pl->wg_tpm,
        pl->wg_mtiles, accumulate, inline_split, image, bc.acc, bc.out, bc.n));
    return check_launch("unpack_wgrad_small_kernel");
  }
  WgradParams& p = pl->wg_params;
  if (!(pl->wg_key_x == x && pl->wg_key_g == dy && pl->wg_key_s == scratch)) {
    memset(&p, 0, sizeof(p));
    const int n_x = f.prog.a_phased ? 8 : 1;
    const int n_g = f.prog.out_phased ? 8 : 1;
    for (int i = 0; i < n_x; ++i) {
      int32_t rc = view_map(&p.x_maps[i], x, pl->vx, f.prog.a_phased, i, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64,
                            pl->wg_box_w, pl->wg_box_h, pl->wg_box_d, 128);
      if (rc) return rc;
    }
    for (int i = 0; i < n_g; ++i) {
      int32_t rc = view_map(&p.g_maps[i], dy, pl->vdy, f.prog.out_phased, i, 2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 64,
                            pl->wg_box_w, pl->wg_box_h, pl->wg_box_d, 128);
      if (rc) return rc;
    }
    uint64_t dims[2] = {(uint64_t)f.prog.max_taps * f.kc_pad, (uint64_t)f.subs.size() * f.R * pl->wg_ksplit};
    uint64_t strides[1] = {dims[0] * 4};
    uint32_t box[2] = {32u, 128u};
    int32_t rc = encode_tmap(&p.d_map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, scratch, dims, strides, box, 128);
    p.image_rows = (int)f.subs.size() * f.R;
    if (rc) return rc;
    for (size_t i = 0; i < f.subs.size(); ++i) p.subs[i] = f.subs[i];
    p.taps = f.d_taps;
    p.tiles_w = (f.out_w + pl->wg_box_w - 1) / pl->wg_box_w;
    p.tiles_h = (f.out_h + pl->wg_box_h - 1) / pl->wg_box_h;
    p.tiles_d = (f.out_d + pl->wg_box_d - 1) / pl->wg_box_d;
    p.batch = pl->desc.n;
    p.box_w = pl->wg_box_w; p.box_h = pl->wg_box_h; p.box_d = pl->wg_box_d;
    p.box_bytes = 64 * 64 * 2;
    p.kc_pad = f.kc_pad;
    p.ksplit = pl->wg_ksplit;
    p.r_tiles = (f.R + 127) / 128;
    p.c_tiles = (f.Kc + pl->wg_block_n - 1) / pl->wg_block_n;
    pl->wg_key_x = x; pl->wg_key_g = dy; pl->wg_key_s = scratch;
  }
  PETSYN_CHECK_CUDA(cudaMemsetAsync(scratch, 0, petsyn_conv_wgrad_scratch_bytes(pl), st));
  dim3 grid((unsigned)(f.prog.max_taps * p.r_tiles * p.c_tiles), (unsigned)p.ksplit, (unsigned)f.subs.size());
  if (pl->wg_block_n == 128) {
    constexpr int STAGES = 3;
    constexpr int smem = STAGES * (16384 + 16384) + 1024 + 256;
    auto kern = wgrad_kernel<128, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    PETSYN_CHECK_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, p));
  } else {
    constexpr int STAGES = 4;
    constexpr int smem = STAGES * (16384 + 8192) + 1024 + 256;
    auto kern = wgrad_kernel<64, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      attr_set = true;
    }
    PETSYN_CHECK_CUDA(launch_pdl(kern, grid, dim3(128), smem, st, p));
  }
  int32_t rc = check_launch("wgrad_kernel");
  if (rc) return rc;
  {
    const bool convt = pl->desc.op == PETSYN_OP_CONVT;
    const int A = convt ? pl->desc.cin : pl->desc.cout, B = convt ? pl->desc.cout : pl->desc.cin;
    // few loads per thread: the splits are added (in split order) by the unpack kernel itself; else one coalesced pass first
    const int inline_split = pl->wg_ksplit * f.prog.max_taps <= 16 ? pl->wg_ksplit : 1;
    if (inline_split == 1) {
      rc = reduce_images(reinterpret_cast<float*>(scratch), pl->wg_ksplit,
                         (int64_t)f.subs.size() * f.R * f.prog.max_taps * f.kc_pad, st);
      if (rc) return rc;
    }
    return launch_unpack(pl->k3, convt, pl->inv_max, reinterpret_cast<const float*>(scratch), dw, pl->d_inv, A, B, f,
                         accumulate, inline_split, bc, st);
  }
}

}  // extern "C"
