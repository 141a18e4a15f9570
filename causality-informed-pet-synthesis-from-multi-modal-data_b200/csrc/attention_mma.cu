// Self-attention of the level-3 SpatialTransformer (atten_unet_model.py:137-154: baddbmm * scale -> softmax -> bmm),
// head dim 32, as flash-style kernels on the warp-level tensor-core path (mma.sync.m16n8k16 bf16 -> fp32): the L x L
// score tensor never exists.  tcgen05 is the wrong tool here: with d_head = 32 a 128-row UMMA tile would spend its
// time in the TMEM <-> register round trips of the online softmax, while the whole layer is ~25 GFLOP.
//
//   qkv   bf16 [N*L, 3*H*32] = (q | k | v), heads contiguous inside each third
//   out   bf16 [N*L, H*32]
//   lse   fp32 [N, H, L]      natural-log sum of exp(scale * q.k) per query (saved for backward)
//
// One CTA = 64 queries (fwd, dq) or 64 keys (dk/dv) of one (sample, head); 4 warps x 16 rows.  The streamed operand
// tiles (64 rows x 32 dims) live in shared memory with an 80-byte row pitch (conflict-free ldmatrix) and are double
// buffered with cp.async.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>

#include "common.h"

namespace petsyn {
namespace fa {

constexpr int kD = 32;          // head dim
constexpr int kT = 64;          // rows per tile
constexpr int kPitch = 40;      // bf16 elements per smem row (80 B)
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Stage a [64 rows x 32 dims] tile (rows r0 .. r0+63 of a [L, ld] matrix starting at `src`) into smem; rows >= L are zero.
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int64_t ld, int r0, int L) {
  for (int e = threadIdx.x & 127; e < kT * 4; e += 128) {
    const int r = e >> 2, c = e & 3;
    const bool ok = r0 + r < L;
    cp_async16(dst + r * kPitch + c * 8, src + (int64_t)(ok ? r0 + r : 0) * ld + c * 8, ok);
  }
}

// A fragments (16 rows x 32 dims = 2 k-steps) of rows row0 + {g, g+8} straight from global memory; rows >= L are zero.
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[2][4], const __nv_bfloat16* src, int64_t ld, int row0, int L) {
  const int g = (threadIdx.x & 31) >> 2, t = threadIdx.x & 3;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + g + (i & 1) * 8;
      const int c = ks * 16 + 2 * t + (i >> 1) * 8;
      a[ks][i] = r < L ? *reinterpret_cast<const uint32_t*>(src + (int64_t)r * ld + c) : 0u;
    }
}

// acc[j] (16 x 8 tile j of a 16 x 64 product) = A (16 x 32, registers) * T^T, T = smem tile [64][32] (row = output column)
__device__ __forceinline__ void gemm_a_tileT(float (&acc)[8][4], const uint32_t (&a)[2][4], const __nv_bfloat16* tile) {
  const int lane = threadIdx.x & 31;
  const uint32_t base = smem_addr(tile + (lane & 7) * kPitch + (lane >> 3) * 8);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t b[4];
    ldsm_x4(b, base + j * 8 * kPitch * 2);
    mma16816(acc[j], a[0], b[0], b[1]);
    mma16816(acc[j], a[1], b[2], b[3]);
  }
}

// o[nt] (16 x 8 tile nt of a 16 x 32 product) += P (16 x 64, as 4 A fragments) * T, T = smem tile [64][32] (row = k)
__device__ __forceinline__ void gemm_p_tile(float (&o)[4][4], const uint32_t (&pa)[4][4], const __nv_bfloat16* tile) {
  const int lane = threadIdx.x & 31;
  const uint32_t base = smem_addr(tile + ((lane & 7) + ((lane >> 3) & 1) * 8) * kPitch + (lane >> 4) * 8);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t b[4];
    ldsm_x4_t(b, base + kk * 16 * kPitch * 2);
    mma16816(o[0], pa[kk], b[0], b[1]);
    mma16816(o[1], pa[kk], b[2], b[3]);
    ldsm_x4_t(b, base + kk * 16 * kPitch * 2 + 32);
    mma16816(o[2], pa[kk], b[0], b[1]);
    mma16816(o[3], pa[kk], b[2], b[3]);
  }
}

// A CTA is G independent groups of four warps that share the same 64 rows and split the streamed tiles between them
// (group g takes tiles g, g + G, ...): with one group a CTA has four warps and an SM two CTAs' worth of work -- every
// mma / exp2 / shuffle of the online softmax waits for the one before it; G groups put G times as many warps on the SM.
// Each group has its own double-buffered tiles and its own named barrier; partial results are merged through shared memory.
constexpr int kTileBytes = kT * kPitch * 2;                       // one staged [64 x 32] tile
constexpr int kGroupBytes = 2 * 2 * kTileBytes + 2 * 2 * kT * 4;   // two buffers x two tiles (+ two buffers x two fp32 rows)
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// ------------------------------------------------------------------------------------------------ forward
template <int G>
__global__ void __launch_bounds__(128 * G) attn_fwd_mma_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                           float* __restrict__ lse, int L, int H, float scale) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t fa_smem[];
  const int grp = threadIdx.x >> 7, ltid = threadIdx.x & 127;
  __nv_bfloat16(*sK)[kT * kPitch] = reinterpret_cast<__nv_bfloat16(*)[kT * kPitch]>(fa_smem + grp * kGroupBytes);
  __nv_bfloat16(*sV)[kT * kPitch] = sK + 2;
  const int n = blockIdx.z, h = blockIdx.y;
  const int warp = ltid >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t ld = 3 * H * kD;
  const __nv_bfloat16* q = qkv + (int64_t)n * L * ld + h * kD;
  const __nv_bfloat16* k = q + H * kD;
  const __nv_bfloat16* v = q + 2 * H * kD;
  const int row0 = blockIdx.x * kT + warp * 16;
  uint32_t aq[2][4];
  load_a_frags(aq, q, ld, row0, L);
  const float c = scale * kLog2e;
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  float o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;

  const int ntiles = (L + kT - 1) / kT;
  if (grp < ntiles) {
    load_tile(sK[0], k, ld, grp * kT, L);
    load_tile(sV[0], v, ld, grp * kT, L);
  }
  cp_async_commit();
  for (int it = grp, li = 0; it < ntiles; it += G, ++li) {
    const int buf = li & 1;
    if (it + G < ntiles) {
      load_tile(sK[buf ^ 1], k, ld, (it + G) * kT, L);
      load_tile(sV[buf ^ 1], v, ld, (it + G) * kT, L);
    }
    cp_async_commit();
    cp_async_wait<1>();
    group_sync(grp);
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
    gemm_a_tileT(s, aq, sK[buf]);
    const int kn = L - it * kT;   // valid keys in this tile
    if (kn < kT) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (j * 8 + 2 * t + (i & 1) >= kn) s[j][i] = -INFINITY;
    }
    float mx[2] = {m[0], m[1]};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    const float corr[2] = {exp2f((m[0] - mx[0]) * c), exp2f((m[1] - mx[1]) * c)};
    m[0] = mx[0]; m[1] = mx[1];
    l[0] *= corr[0]; l[1] *= corr[1];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
      o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    uint32_t pa[4][4];
    const float mc[2] = {mx[0] * c, mx[1] * c};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] * c - mc[0]), p1 = exp2f(s[j][1] * c - mc[0]);
      const float p2 = exp2f(s[j][2] * c - mc[1]), p3 = exp2f(s[j][3] * c - mc[1]);
      l[0] += p0 + p1;
      l[1] += p2 + p3;
      pa[j >> 1][(j & 1) * 2] = pack2(p0, p1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
    }
    gemm_p_tile(o, pa, sV[buf]);
    group_sync(grp);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 1);
    l[r] += __shfl_xor_sync(0xffffffffu, l[r], 2);
  }
  if (G > 1) {
    // merge the groups' online-softmax states (m, l, o) in group order: groups 1 .. G-1 leave theirs in their own tile
    // buffers (laid out [value][thread]), group 0 rescales to the common maximum and adds
    float* mine = reinterpret_cast<float*>(fa_smem + grp * kGroupBytes);
    if (grp > 0) {
      mine[0 * 128 + ltid] = m[0]; mine[1 * 128 + ltid] = m[1];
      mine[2 * 128 + ltid] = l[0]; mine[3 * 128 + ltid] = l[1];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) mine[(4 + nt * 4 + i) * 128 + ltid] = o[nt][i];
    }
    __syncthreads();
    if (grp > 0) return;
    for (int og = 1; og < G; ++og) {
      const float* oth = reinterpret_cast<const float*>(fa_smem + og * kGroupBytes);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float m2 = oth[r * 128 + ltid], l2 = oth[(2 + r) * 128 + ltid];
        const float mn = fmaxf(m[r], m2);
        const float a1 = m[r] == -INFINITY ? 0.f : exp2f((m[r] - mn) * c);
        const float a2 = m2 == -INFINITY ? 0.f : exp2f((m2 - mn) * c);
        m[r] = mn;
        l[r] = l[r] * a1 + l2 * a2;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          o[nt][r * 2] = o[nt][r * 2] * a1 + oth[(4 + nt * 4 + r * 2) * 128 + ltid] * a2;
          o[nt][r * 2 + 1] = o[nt][r * 2 + 1] * a1 + oth[(4 + nt * 4 + r * 2 + 1) * 128 + ltid] * a2;
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + r * 8;
    if (row >= L) continue;
    const float inv = 1.f / l[r];
    __nv_bfloat16* dst = out + ((int64_t)n * L + row) * (H * kD) + h * kD;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8 + 2 * t) = pack2(o[nt][r * 2] * inv, o[nt][r * 2 + 1] * inv);
    if (t == 0) lse[((int64_t)n * H + h) * L + row] = m[r] * scale + __logf(l[r]);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
// dS = P * (dP - delta), dP = dO V^T, dQ = scale * dS K
template <int G>
__global__ void __launch_bounds__(128 * G) attn_bwd_dq_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                              const __nv_bfloat16* __restrict__ dout,
                                                              const float* __restrict__ lse, const float* __restrict__ delta,
                                                              __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t fa_smem[];
  const int grp = threadIdx.x >> 7, ltid = threadIdx.x & 127;
  __nv_bfloat16(*sK)[kT * kPitch] = reinterpret_cast<__nv_bfloat16(*)[kT * kPitch]>(fa_smem + grp * kGroupBytes);
  __nv_bfloat16(*sV)[kT * kPitch] = sK + 2;
  const int n = blockIdx.z, h = blockIdx.y;
  const int warp = ltid >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t ld = 3 * H * kD;
  const __nv_bfloat16* q = qkv + (int64_t)n * L * ld + h * kD;
  const __nv_bfloat16* k = q + H * kD;
  const __nv_bfloat16* v = q + 2 * H * kD;
  const __nv_bfloat16* go = dout + (int64_t)n * L * (H * kD) + h * kD;
  const int row0 = blockIdx.x * kT + warp * 16;
  uint32_t aq[2][4], ag[2][4];
  load_a_frags(aq, q, ld, row0, L);
  load_a_frags(ag, go, H * kD, row0, L);
  const float c = scale * kLog2e;
  float ls[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + r * 8;
    ls[r] = row < L ? lse[((int64_t)n * H + h) * L + row] * kLog2e : 0.f;
    dl[r] = row < L ? delta[((int64_t)n * H + h) * L + row] : 0.f;
  }
  float dq[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dq[i][j] = 0.f;

  const int ntiles = (L + kT - 1) / kT;
  if (grp < ntiles) {
    load_tile(sK[0], k, ld, grp * kT, L);
    load_tile(sV[0], v, ld, grp * kT, L);
  }
  cp_async_commit();
  for (int it = grp, li = 0; it < ntiles; it += G, ++li) {
    const int buf = li & 1;
    if (it + G < ntiles) {
      load_tile(sK[buf ^ 1], k, ld, (it + G) * kT, L);
      load_tile(sV[buf ^ 1], v, ld, (it + G) * kT, L);
    }
    cp_async_commit();
    cp_async_wait<1>();
    group_sync(grp);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = dp[j][i] = 0.f;
    gemm_a_tileT(s, aq, sK[buf]);
    gemm_a_tileT(dp, ag, sV[buf]);
    const int kn = L - it * kT;
    uint32_t da[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float ds[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1;
        const float p = (j * 8 + 2 * t + (i & 1) < kn) ? exp2f(s[j][i] * c - ls[r]) : 0.f;
        ds[i] = p * (dp[j][i] - dl[r]);
      }
      da[j >> 1][(j & 1) * 2] = pack2(ds[0], ds[1]);
      da[j >> 1][(j & 1) * 2 + 1] = pack2(ds[2], ds[3]);
    }
    gemm_p_tile(dq, da, sK[buf]);
    group_sync(grp);
  }
  if (G > 1) {                         // add the groups' partial dQ in group order
    float* mine = reinterpret_cast<float*>(fa_smem + grp * kGroupBytes);
    if (grp > 0) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) mine[(nt * 4 + i) * 128 + ltid] = dq[nt][i];
    }
    __syncthreads();
    if (grp > 0) return;
    for (int og = 1; og < G; ++og) {
      const float* oth = reinterpret_cast<const float*>(fa_smem + og * kGroupBytes);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) dq[nt][i] += oth[(nt * 4 + i) * 128 + ltid];
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + r * 8;
    if (row >= L) continue;
    __nv_bfloat16* dst = dqkv + ((int64_t)n * L + row) * ld + h * kD;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      *reinterpret_cast<uint32_t*>(dst + nt * 8 + 2 * t) = pack2(dq[nt][r * 2] * scale, dq[nt][r * 2 + 1] * scale);
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
// One CTA = 64 keys.  S^T = K Q^T, P^T = exp(scale S^T - lse[q]), dV = P^T dO, dP^T = V dO^T,
// dS^T = P^T * (dP^T - delta[q]), dK = scale * dS^T Q
template <int G>
__global__ void __launch_bounds__(128 * G) attn_bwd_dkv_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                               const __nv_bfloat16* __restrict__ dout,
                                                               const float* __restrict__ lse, const float* __restrict__ delta,
                                                               __nv_bfloat16* __restrict__ dqkv, int L, int H, float scale) {
  pdl_sync();
  extern __shared__ __align__(16) uint8_t fa_smem[];
  const int grp = threadIdx.x >> 7, ltid = threadIdx.x & 127;
  __nv_bfloat16(*sQ)[kT * kPitch] = reinterpret_cast<__nv_bfloat16(*)[kT * kPitch]>(fa_smem + grp * kGroupBytes);
  __nv_bfloat16(*sG)[kT * kPitch] = sQ + 2;
  float(*sL)[kT] = reinterpret_cast<float(*)[kT]>(fa_smem + grp * kGroupBytes + 4 * kTileBytes);
  float(*sD)[kT] = sL + 2;
  const int n = blockIdx.z, h = blockIdx.y;
  const int warp = ltid >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t ld = 3 * H * kD;
  const __nv_bfloat16* q = qkv + (int64_t)n * L * ld + h * kD;
  const __nv_bfloat16* k = q + H * kD;
  const __nv_bfloat16* v = q + 2 * H * kD;
  const __nv_bfloat16* go = dout + (int64_t)n * L * (H * kD) + h * kD;
  const float* lrow = lse + ((int64_t)n * H + h) * L;
  const float* drow = delta + ((int64_t)n * H + h) * L;
  const int row0 = blockIdx.x * kT + warp * 16;   // key rows of this warp
  uint32_t ak[2][4], av[2][4];
  load_a_frags(ak, k, ld, row0, L);
  load_a_frags(av, v, ld, row0, L);
  const float c = scale * kLog2e;
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dk[i][j] = dv[i][j] = 0.f;

  const int ntiles = (L + kT - 1) / kT;
  auto stage = [&](int buf, int q0) {
    load_tile(sQ[buf], q, ld, q0, L);
    load_tile(sG[buf], go, H * kD, q0, L);
    if (ltid < kT) {
      const bool ok = q0 + ltid < L;
      sL[buf][ltid] = ok ? lrow[q0 + ltid] * kLog2e : INFINITY;   // +inf -> P = 0 for padded queries
      sD[buf][ltid] = ok ? drow[q0 + ltid] : 0.f;
    }
  };
  if (grp < ntiles) stage(0, grp * kT);
  cp_async_commit();
  for (int it = grp, li = 0; it < ntiles; it += G, ++li) {
    const int buf = li & 1;
    if (it + G < ntiles) stage(buf ^ 1, (it + G) * kT);
    cp_async_commit();
    cp_async_wait<1>();
    group_sync(grp);
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = dp[j][i] = 0.f;
    gemm_a_tileT(s, ak, sQ[buf]);     // [16 keys x 64 queries]
    gemm_a_tileT(dp, av, sG[buf]);
    uint32_t pa[4][4], da[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[4], ds[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int col = j * 8 + 2 * t + (i & 1);
        p[i] = exp2f(s[j][i] * c - sL[buf][col]);
        ds[i] = p[i] * (dp[j][i] - sD[buf][col]);
      }
      pa[j >> 1][(j & 1) * 2] = pack2(p[0], p[1]);
      pa[j >> 1][(j & 1) * 2 + 1] = pack2(p[2], p[3]);
      da[j >> 1][(j & 1) * 2] = pack2(ds[0], ds[1]);
      da[j >> 1][(j & 1) * 2 + 1] = pack2(ds[2], ds[3]);
    }
    gemm_p_tile(dv, pa, sG[buf]);
    gemm_p_tile(dk, da, sQ[buf]);
    group_sync(grp);
  }
  if (G > 1) {                         // add the groups' partial dK / dV in group order
    float* mine = reinterpret_cast<float*>(fa_smem + grp * kGroupBytes);
    if (grp > 0) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          mine[(nt * 4 + i) * 128 + ltid] = dk[nt][i];
          mine[(16 + nt * 4 + i) * 128 + ltid] = dv[nt][i];
        }
    }
    __syncthreads();
    if (grp > 0) return;
    for (int og = 1; og < G; ++og) {
      const float* oth = reinterpret_cast<const float*>(fa_smem + og * kGroupBytes);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          dk[nt][i] += oth[(nt * 4 + i) * 128 + ltid];
          dv[nt][i] += oth[(16 + nt * 4 + i) * 128 + ltid];
        }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = row0 + g + r * 8;
    if (row >= L) continue;
    __nv_bfloat16* dkp = dqkv + ((int64_t)n * L + row) * ld + H * kD + h * kD;
    __nv_bfloat16* dvp = dkp + H * kD;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(dkp + nt * 8 + 2 * t) = pack2(dk[nt][r * 2] * scale, dk[nt][r * 2 + 1] * scale);
      *reinterpret_cast<uint32_t*>(dvp + nt * 8 + 2 * t) = pack2(dv[nt][r * 2], dv[nt][r * 2 + 1]);
    }
  }
}

// delta[n,h,t] = sum_d dO * O   (one thread per (token, head): 2 x 64-byte rows)
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o,
                                                         const __nv_bfloat16* __restrict__ dout,
                                                         float* __restrict__ delta, int N, int L, int H) {
  pdl_sync();
  const int64_t total = (int64_t)N * L * H;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i % H);
    const int64_t tok = i / H;          // n * L + t
    const int64_t off = tok * (H * kD) + h * kD;
    float s = 0.f;
#pragma unroll
    for (int cch = 0; cch < 4; ++cch) {
      const uint4 a = *reinterpret_cast<const uint4*>(o + off + cch * 8);
      const uint4 b = *reinterpret_cast<const uint4*>(dout + off + cch * 8);
      const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = __bfloat1622float2(a2[j]), y = __bfloat1622float2(b2[j]);
        s += x.x * y.x + x.y * y.y;
      }
    }
    const int n = (int)(tok / L), t = (int)(tok % L);
    delta[((int64_t)n * H + h) * L + t] = s;
  }
}

}  // namespace fa
}  // namespace petsyn

using namespace petsyn;

// groups of four warps per CTA (PETSYN_ATTN_GROUPS = 1, 2 or 4; default 2): see the note above attn_fwd_mma_kernel
static int fa_groups() {
  static int g = 0;
  if (g == 0) {
    const char* e = getenv("PETSYN_ATTN_GROUPS");
    g = e != nullptr ? atoi(e) : 2;
    if (g != 1 && g != 2 && g != 4) g = 2;
  }
  return g;
}
#define PETSYN_FA_LAUNCH_G(KERNEL, G, GRID, ST, ...)                                                                  \
  {                                                                                                                   \
    static bool attr_ = false;                                                                                        \
    if (!attr_) {                                                                                                     \
      PETSYN_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * fa::kGroupBytes)); \
      attr_ = true;                                                                                                   \
    }                                                                                                                 \
    PETSYN_CHECK_CUDA(launch_pdl(KERNEL<G>, dim3(GRID), dim3(128 * G), (size_t)(G * fa::kGroupBytes), ST, __VA_ARGS__)); \
  }
#define PETSYN_FA_LAUNCH(KERNEL, GRID, ST, ...)                                       \
  do {                                                                                \
    const int g_ = fa_groups();                                                       \
    if (g_ == 1) PETSYN_FA_LAUNCH_G(KERNEL, 1, GRID, ST, __VA_ARGS__)                  \
    else if (g_ == 2) PETSYN_FA_LAUNCH_G(KERNEL, 2, GRID, ST, __VA_ARGS__)             \
    else PETSYN_FA_LAUNCH_G(KERNEL, 4, GRID, ST, __VA_ARGS__)                          \
  } while (0)

extern "C" {

int32_t petsyn_attention_fwd(const void* qkv, void* out, float* lse, int32_t n, int32_t l, int32_t heads,
                             int32_t head_dim, float scale, void* stream) {
  PETSYN_REQUIRE(qkv && out && lse && n > 0 && l > 0 && heads > 0, "bad argument");
  PETSYN_REQUIRE(head_dim == fa::kD, "attention kernels are specialised for head_dim 32 (num_head_channels=32)");
  dim3 grid((unsigned)((l + fa::kT - 1) / fa::kT), (unsigned)heads, (unsigned)n);
  const auto* qp = reinterpret_cast<const __nv_bfloat16*>(qkv);
  auto* op = reinterpret_cast<__nv_bfloat16*>(out);
  PETSYN_FA_LAUNCH(fa::attn_fwd_mma_kernel, grid, as_stream(stream), qp, op, lse, l, heads, scale);
  return check_launch("attn_fwd_mma_kernel");
}

int32_t petsyn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                             void* dqkv, int32_t n, int32_t l, int32_t heads, int32_t head_dim, float scale,
                             void* stream) {
  PETSYN_REQUIRE(qkv && out && dout && lse && delta && dqkv && n > 0 && l > 0 && heads > 0, "bad argument");
  PETSYN_REQUIRE(head_dim == fa::kD, "attention kernels are specialised for head_dim 32 (num_head_channels=32)");
  cudaStream_t st = as_stream(stream);
  const auto* qp = reinterpret_cast<const __nv_bfloat16*>(qkv);
  const auto* gp = reinterpret_cast<const __nv_bfloat16*>(dout);
  const int64_t total = (int64_t)n * heads * l;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148 * 16));
  PETSYN_CHECK_CUDA(launch_pdl(fa::attn_delta_kernel, dim3(blocks), dim3(256), 0, st, reinterpret_cast<const __nv_bfloat16*>(out), gp, delta, n, l, heads));
  int32_t rc = check_launch("attn_delta_kernel");
  if (rc) return rc;
  dim3 grid((unsigned)((l + fa::kT - 1) / fa::kT), (unsigned)heads, (unsigned)n);
  auto* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  PETSYN_FA_LAUNCH(fa::attn_bwd_dq_mma_kernel, grid, st, qp, gp, lse, delta, dq, l, heads, scale);
  rc = check_launch("attn_bwd_dq_mma_kernel");
  if (rc) return rc;
  PETSYN_FA_LAUNCH(fa::attn_bwd_dkv_mma_kernel, grid, st, qp, gp, lse, delta, dq, l, heads, scale);
  return check_launch("attn_bwd_dkv_mma_kernel");
}

}  // extern "C"
