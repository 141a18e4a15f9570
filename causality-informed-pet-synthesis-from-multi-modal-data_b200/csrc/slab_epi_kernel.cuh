// Depth-folded 3x3x3 slab convolution (slab_conv3_kernel of slab_kernels.cuh) with a FUSED EPILOGUE: the passes over the
// output tensor that a ResnetBlock (atten_unet_model.py:641-662) otherwise runs as separate HBM-bound kernels are done by
// the epilogue warps while the tile is still in registers.
//
//   kEpiSide        y = conv(a) + bias + side          the residual sum  `conv2(...) + skip(x)`  (:662); `side` is a bf16
//                                                      tensor with the output's geometry, one TMA box load per tile
//   kEpiStats       sum y, sum y^2 per (sample, ch)    the statistics of the GroupNorm that consumes y (norm2 after conv1,
//                                                      norm1 of the next block after the residual sum), of the bf16 values
//                                                      as stored
//   kEpiNormReduce  data gradient da = dgrad(dy) and, with z = `side` (the input of the normalisation in front of this
//                   conv: a = act(z * scale + shift)),  S0 = sum g,  S1 = sum g * zhat,  g = da * act'(z * scale + shift):
//                   the reduction pass of the normalisation backward (nx::bwd_reduce_kernel), which then only runs its
//                   apply pass
//
// Sums are kept per thread (one accumulator row = one voxel) in registers over all tiles of a work item, folded over the
// 128 epilogue threads once per item (warp butterfly + four partials in shared memory, fixed order) and added to DOUBLE
// accumulators in global memory (exact for fp32 partials, see det_reduce.cuh: the result does not depend on the order the
// CTAs finish in).
//
// Warp roles as in slab_conv3_kernel: warp 0 = TMA producer (slabs + side tiles), warp 1 = MMA issuer, warps 2..5 = epilogue.
#pragma once
#include "slab_kernels.cuh"

namespace petsyn {

enum { kEpiSide = 1, kEpiStats = 2, kEpiNormReduce = 4 };
#ifndef PETSYN_EPI_MINB
#define PETSYN_EPI_MINB 3       // resident CTAs per SM the 16 -> 16 variants are compiled for
#endif
constexpr int kEpiMaxBatch = 8;                        // samples whose normalisation constants fit the shared-memory table

struct alignas(64) SlabEpi {
  CUtensorMap e_map;        // side tensor: bf16 view (C, W, H, D, N) with the output's dims; box (N, 8, 16, 1, 1); no swizzle
  int32_t flags;
  int32_t side_ring;        // side tiles in flight: 2 or 4
  int32_t nact;             // kEpiNormReduce: activation behind the normalisation
  float nslope;
  double* st1; int32_t st1_c, st1_off;     // kEpiStats: [sample][2][st_c] accumulators, this conv's channels at st_off
  double* st2; int32_t st2_c, st2_off;     // optional second consumer of the same values
  const float *nscale, *nshift, *nmean, *nrstd;   // kEpiNormReduce: [sample][N]
  double* bsums;                                   // kEpiNormReduce: [sample][2][N]  (S0 | S1)
};

__host__ __device__ inline int slab3_epi_extra_bytes(int n, int side_ring, int batch) {
  return side_ring * 128 * n * 2 + batch * 4 * n * 4 + 4 * 2 * n * 4 + 256;   // side_ring 0: statistics only
}

__device__ __forceinline__ float epi_act_grad(float b, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return b > 0.f ? 1.f : 0.f;
    case PETSYN_ACT_LRELU:
    case PETSYN_ACT_PRELU: return b > 0.f ? 1.f : slope;
    case PETSYN_ACT_SILU: { const float s = __fdividef(1.f, 1.f + __expf(-b)); return s * (1.f + b * (1.f - s)); }
    case PETSYN_ACT_TANH: { const float t = tanhf(b); return 1.f - t * t; }
    default: return 1.f;
  }
}

// Sum a[0..31] over the 32 lanes of a warp: afterwards lane l holds the total of a[l] in a[0] (31 shuffles).
__device__ __forceinline__ float warp_transpose_sum32(float (&a)[32], int lane) {
#pragma unroll
  for (int off = 16, k = 16; off >= 1; off >>= 1, k >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < k; ++i) {
      const float send = up ? a[i] : a[i + k];
      const float keep = up ? a[i + k] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}

// FLAGS (compile time) = the epilogue's work: kEpiStats | [kEpiSide]  or  kEpiSide  or  kEpiNormReduce
template <int ATOMS, int NB, int FLAGS>
__global__ void __launch_bounds__(192, (NB == 1 && ATOMS == 1) ? PETSYN_EPI_MINB : (NB * ATOMS <= 2 ? 2 : 1))
    slab_conv3_epi_kernel(const __grid_constant__ SlabParams p, const __grid_constant__ SlabEpi e) {
  constexpr int N = NB * 16;
  constexpr bool kSide = (FLAGS & kEpiSide) != 0, kStats = (FLAGS & kEpiStats) != 0, kNorm = (FLAGS & kEpiNormReduce) != 0;
  static_assert(!(kNorm && (kSide || kStats)), "the norm-backward epilogue belongs to the data gradient alone");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int wbytes = (9 * ATOMS * 5 * N * 32 + 1023) / 1024 * 1024;
  constexpr int stg_bytes = (2 * 128 * N * 2 + 1023) / 1024 * 1024;
  constexpr int side_bytes = 128 * N * 2;
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + wbytes;
  const int R = p.ring;
  const uint32_t Rm = uint32_t(R - 1);
  uint8_t* s_stg = s_ring + R * p.slab_bytes;
  uint8_t* s_side = s_stg + stg_bytes;                                        // [side_ring][128][N] bf16
  float* s_tab = reinterpret_cast<float*>(s_side + e.side_ring * side_bytes);  // [batch][4][N]: scale, shift, mean, rstd
  float* s_x = s_tab + p.batch * 4 * N;                                        // [4 warps][2 N] partial sums
  uint8_t* tail = reinterpret_cast<uint8_t*>(s_x + 4 * 2 * N);
  tail = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tail) + 15) & ~uintptr_t(15));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);           // [ring]
  uint64_t* empty_bar = full_bar + kSlabMaxRing;                    // [ring]
  uint64_t* w_bar = empty_bar + kSlabMaxRing;                       // [1]
  uint64_t* acc_full = w_bar + 1;                                   // [3]
  uint64_t* acc_free = acc_full + 3;                                // [3]
  uint64_t* side_full = acc_free + 3;                               // [4]
  uint64_t* side_empty = side_full + 4;                             // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(side_empty + 4);
  constexpr uint32_t tmem_cols = 3 * N <= 64 ? 64 : (3 * N <= 128 ? 128 : 256);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  constexpr bool use_side = kSide || kNorm;
  constexpr bool want_sums = kStats || kNorm;
  const uint32_t SRm = uint32_t(e.side_ring - 1);
  const uint32_t sr_shift = e.side_ring == 4 ? 2 : 1;

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < R; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    for (int b = 0; b < 3; ++b) {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_free[b], 4);
    }
    for (int b = 0; b < 4; ++b) {
      ptx::mbar_init(&side_full[b], 1);
      ptx::mbar_init(&side_empty[b], 4);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.a_map);
    ptx::prefetch_tmap(&p.b_map);
    ptx::prefetch_tmap(&p.c_map);
    if (use_side) ptx::prefetch_tmap(&e.e_map);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::pdl_sync();         // everything above is on-chip set-up: it overlaps the tail of the previous kernel
  if (kNorm && warp >= 2) {
    // normalisation constants of every sample -> shared memory (read as broadcast 16-byte loads by the epilogue)
    const int cnt = p.batch * N;
    for (int i = tid - 64; i < cnt; i += 128) {
      const int s = i / N, c = i - s * N;
      float* t = s_tab + s * 4 * N;
      t[c] = e.nscale[i];
      t[N + c] = e.nshift[i];
      t[2 * N + c] = e.nmean[i];
      t[3 * N + c] = e.nrstd[i];
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;

  const int G = gridDim.x;
  if (warp == 0) {
    if (ptx::elect_one()) {
      // ---------------- TMA producer: weights once, then per depth step one slab and (real tiles) one side tile ----------------
      ptx::mbar_expect_tx(w_bar, uint32_t(9 * ATOMS * 5 * N * 32));
      for (int j = 0; j < 9; ++j)
        for (int q = 0; q < ATOMS; ++q)
          for (int b = 0; b < 5; ++b)
            ptx::tma_load_2d(s_w + ((j * ATOMS + q) * 5 + b) * N * 32, &p.b_map, w_bar, p.wtap[j][b] * p.kc_pad + q * 16,
                             p.b_row);
      uint32_t seq = 0, nside = 0;
      for (int item = blockIdx.x; item < p.items; item += G) {
        int t = item;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int ch = t % p.nchunks; t /= p.nchunks;
        const int nb = t;
        const int d0 = ch * p.dchunk;
        const int len = min(p.dchunk, p.D - d0);
        for (int s = 0; s < len + 2; ++s, ++seq) {
          const int slot = seq & Rm;
          const uint32_t ph = (seq / uint32_t(R)) & 1;
          ptx::mbar_wait(&empty_bar[slot], ph ^ 1);
          const int d = d0 - 1 + s;
          const bool oob = d < 0 || d >= p.D;
          ptx::mbar_expect_tx(&full_bar[slot], uint32_t(p.slab_tx));
          ptx::tma_load_5d(s_ring + slot * p.slab_bytes, &p.a_map, &full_bar[slot], 0, oob ? p.W + 64 : tw * kSlabW - 1, 0,
                           th * kSlabH - 1, oob ? 0 : nb * p.D + d);
          if (use_side && s >= 2) {
            // the tile this slab closes (depth d0 + s - 2): its side tile is consumed by the epilogue one step later
            const uint32_t ss = nside & SRm;
            ptx::mbar_wait(&side_empty[ss], ((nside >> sr_shift) & 1) ^ 1);
            ptx::mbar_expect_tx(&side_full[ss], uint32_t(side_bytes));
            ptx::tma_load_5d(s_side + ss * side_bytes, &e.e_map, &side_full[ss], 0, tw * kSlabW, th * kSlabH, d0 + s - 2, nb);
            ++nside;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ---------------- MMA issuer (identical to slab_conv3_kernel) ----------------
      const uint64_t a_desc_base = ptx::umma_desc_base(16, uint32_t(ATOMS * kSlabWp * 32), 6);
      const uint64_t b_desc_base = ptx::umma_desc_base(16, 256, 6);
      const uint32_t a_hi = uint32_t(a_desc_base >> 32), a_lo0 = uint32_t(a_desc_base);
      const uint32_t b_hi = uint32_t(b_desc_base >> 32), b_lo0 = uint32_t(b_desc_base);
      const uint32_t idesc = ptx::umma_idesc_bf16(128, uint32_t(3 * N), 0, 0);
      const uint32_t ring_lo = a_lo0 + (ptx::smem_u32(s_ring) >> 4);
      const uint32_t w_lo = b_lo0 + (ptx::smem_u32(s_w) >> 4);
      const uint32_t slab16 = uint32_t(p.slab_bytes) >> 4;
      constexpr uint32_t blk16 = uint32_t(N) * 2;
      const uint32_t rshift = (R == 8) ? 3 : (R == 4 ? 2 : 1);
      ptx::mbar_wait(w_bar, 0);
      for (int b = 0; b < 3; ++b) ptx::mbar_wait(&acc_free[b], 0);
      ptx::tc_fence_after_sync();
      uint32_t g = 0, r = 0;
      for (int item = blockIdx.x; item < p.items; item += G) {
        const int ch = (item / (p.tiles_w * p.tiles_h)) % p.nchunks;
        const int len = min(p.dchunk, p.D - ch * p.dchunk);
        for (int s = 0; s < len + 2; ++s, ++g) {
          ptx::mbar_wait(&full_bar[g & Rm], (g >> rshift) & 1);
          if (g > 0) {
            const uint32_t t2 = g + 2;
            ptx::mbar_wait(&acc_free[t2 % 3], (t2 / 3) & 1);
          }
          ptx::tc_fence_after_sync();
          const uint32_t slab_lo = ring_lo + (g & Rm) * slab16;
          const uint32_t win = (r == 0 ? 0u : (r == 1 ? 2u : 1u)) * blk16;
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            const uint32_t a_lo = slab_lo + uint32_t(p.pair_off[j]);
#pragma unroll
            for (int q = 0; q < ATOMS; ++q)
              umma_bf16_split(tmem_base, a_lo + q * (kSlabWp * 2), a_hi, w_lo + (uint32_t(j * ATOMS + q) * 5) * blk16 + win, b_hi,
                              idesc, 1u);
          }
          ptx::umma_commit(&acc_full[r]);
          ptx::umma_commit(&empty_bar[g & Rm]);
          r = r == 2 ? 0 : r + 1;
        }
      }
    }
  } else {
    // ---------------- epilogue warps (2..5) ----------------
    const int quad = warp & 3;
    const int lane = tid & 31;
    const int row = quad * 32 + lane;
    const int rw = row & 7, rh = row >> 3;                       // voxel (w, h) of this accumulator row inside the tile
    const bool leader = (warp == 2) && (lane == 0);
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    uint32_t zeros[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zeros[i] = 0u;
    for (int c0 = 0; c0 < 3 * N; c0 += 16) ptx::tmem_st_32x16(lane_addr + uint32_t(c0), zeros);
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncwarp();
    if (lane == 0)
      for (int b = 0; b < 3; ++b) ptx::mbar_arrive(&acc_free[b]);
    constexpr int NA = want_sums ? N : 1;
    float a0[NA], a1[NA];                                        // per-row sums over the tiles of the current item
#pragma unroll
    for (int i = 0; i < NA; ++i) a0[i] = a1[i] = 0.f;
    uint32_t g = 0, r = 0, stores = 0, nside = 0;
    for (int item = blockIdx.x; item < p.items; item += G) {
      int t = item;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int ch = t % p.nchunks; t /= p.nchunks;
      const int nb = t;
      const int d0 = ch * p.dchunk;
      const int len = min(p.dchunk, p.D - d0);
      const float vm = (tw * kSlabW + rw < p.W && th * kSlabH + rh < p.H) ? 1.f : 0.f;   // rows outside the volume
      const float* tab = s_tab + nb * 4 * N;
      for (int s = 0; s < len + 2; ++s, ++g) {
        ptx::mbar_wait(&acc_full[r], (g / 3) & 1);
        ptx::tc_fence_after_sync();
        const bool real = s >= 2;
        const uint32_t buf = stores & 1;
        uint8_t* stg = s_stg + buf * (128 * N * 2) + row * (N * 2);
        if (real) {
          if (leader) tma_store_wait_read_1();
          named_bar_sync(1, 128);
        }
        const uint32_t taddr = lane_addr + r * uint32_t(N);
        uint32_t v[NB][16];
        if (real) {
#pragma unroll
          for (int c = 0; c < NB; ++c) ptx::tmem_ld_32x16(taddr + uint32_t(c * 16), v[c]);
          ptx::tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < NB; ++c) ptx::tmem_st_32x16(taddr + uint32_t(c * 16), zeros);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&acc_free[r]);
        if (real) {
          const uint32_t ss = nside & SRm;
          const uint8_t* srow = s_side + ss * side_bytes + row * (N * 2);
          if (use_side) ptx::mbar_wait(&side_full[ss], (nside >> sr_shift) & 1);
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            const int c0 = c * 16;
            float f[16], sd[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[c][i]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + c0 + i);
            }
            if (use_side) {
              const uint4 s0 = *reinterpret_cast<const uint4*>(srow + c0 * 2);
              const uint4 s1 = *reinterpret_cast<const uint4*>(srow + c0 * 2 + 16);
              const uint32_t sw[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 ff = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&sw[i]));
                sd[2 * i] = ff.x;
                sd[2 * i + 1] = ff.y;
              }
              if constexpr (kSide) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] += sd[i];
              }
            }
            if (p.epi_act != PETSYN_ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], p.epi_act, p.epi_slope);
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&b2);
              const float2 back = __bfloat1622float2(b2);        // the values as stored
              f[2 * i] = back.x;
              f[2 * i + 1] = back.y;
            }
            *reinterpret_cast<uint4*>(stg + c0 * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(stg + c0 * 2 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            if constexpr (kStats) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float x = f[i] * vm;
                a0[c0 + i] += x;
                a1[c0 + i] += x * x;
              }
            } else if constexpr (kNorm) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 sc = *reinterpret_cast<const float4*>(tab + c0 + 4 * q);
                const float4 sh = *reinterpret_cast<const float4*>(tab + N + c0 + 4 * q);
                const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const float z = sd[4 * q + i];
                  const float gq = f[4 * q + i] * vm * epi_act_grad(z * scv[i] + shv[i], e.nact, e.nslope);
                  a0[c0 + 4 * q + i] += gq;
                  a1[c0 + 4 * q + i] += gq * z;
                }
              }
            }
          }
          if (use_side) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&side_empty[ss]);
            ++nside;
          }
          ptx::fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (leader) {
            const uint8_t* src = s_stg + buf * (128 * N * 2);
            if (p.reduce)
              ptx::tma_reduce_add_5d(&p.c_map, src, 0, tw * kSlabW, th * kSlabH, d0 + s - 2, nb);
            else
              ptx::tma_store_5d(&p.c_map, src, 0, tw * kSlabW, th * kSlabH, d0 + s - 2, nb);
            ptx::tma_store_commit();
          }
          ++stores;
        }
        r = r == 2 ? 0 : r + 1;
      }
      if constexpr (want_sums) {
        // ---- fold the per-row sums of this item over the 128 epilogue threads and add them to the global accumulators ----
        if constexpr (kNorm) {
#pragma unroll
          for (int i = 0; i < N; ++i) a1[i] = (a1[i] - tab[2 * N + i] * a0[i]) * tab[3 * N + i];   // sum g * (z - mu) * rstd
        }
        if constexpr (N == 16) {
          float a[32];
#pragma unroll
          for (int i = 0; i < 16; ++i) { a[i] = a0[i]; a[16 + i] = a1[i]; }
          s_x[quad * 32 + lane] = warp_transpose_sum32(a, lane);           // lane l: l < 16 -> S0[l], else S1[l - 16]
        } else {
          float a[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = a0[i];
          s_x[quad * 64 + lane] = warp_transpose_sum32(a, lane);
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = a1[i];
          s_x[quad * 64 + 32 + lane] = warp_transpose_sum32(a, lane);
        }
        named_bar_sync(2, 128);
        const int e_tid = tid - 64;
        if (e_tid < 2 * N) {
          const float tot = ((s_x[e_tid] + s_x[2 * N + e_tid]) + s_x[4 * N + e_tid]) + s_x[6 * N + e_tid];
          const int which = e_tid / N, c = e_tid - which * N;
          if constexpr (kNorm) {
            atomicAdd(e.bsums + (int64_t)nb * 2 * N + which * N + c, (double)tot);
          } else {
            if (e.st1) atomicAdd(e.st1 + (int64_t)nb * 2 * e.st1_c + which * e.st1_c + e.st1_off + c, (double)tot);
            if (e.st2) atomicAdd(e.st2 + (int64_t)nb * 2 * e.st2_c + which * e.st2_c + e.st2_off + c, (double)tot);
          }
        }
        named_bar_sync(2, 128);
#pragma unroll
        for (int i = 0; i < NA; ++i) a0[i] = a1[i] = 0.f;
      }
    }
    if (leader) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// host entry points (slab_epi.cu): resident CTAs per SM of the instantiation for `smem` bytes, and the launch
int slab3_epi_occupancy(int atoms, int nb, int flags, int smem);
int32_t slab3_epi_launch(int atoms, int nb, int flags, const SlabParams& p, const SlabEpi& e, int grid, int smem,
                         cudaStream_t st);

}  // namespace petsyn
