// Small-channel 3x3x3 stride-1 convolutions (the AttenUNet full- and half-resolution layers: 16..64 channels).
//
// With so few channels the gather-form kernel of igemm_kernels.cuh is bound by shared-memory traffic, not by HBM or the
// tensor pipe: every kernel tap re-loads its (shifted) activation box through TMA, so an activation voxel crosses the
// L2 -> smem path 27 times.  The kernels here keep a sliding window of input SLABS (one depth slice of the tile plus
// its halo) in shared memory instead and let the 27 taps address shifted views of those slabs directly through the
// UMMA shared-memory descriptor (the hardware swizzle is a function of the absolute shared-memory address, so a
// descriptor may start at any 32-byte voxel row of a 32B-swizzled brick).
//
//  slab_conv3_kernel : fprop / dgrad of 3x3x3 convs (the default).  Persistent CTAs sweep columns of 8(w) x 16(h) output
//                      tiles along depth; each depth step loads ONE halo slab (10 x 18 voxels x Cin) by TMA.  The three
//                      depth taps of every (dh, dw) pair are stacked in the MMA's N dimension (N = 3 * Cout), so a slab
//                      feeds the three output tiles it touches with 9 * Cin/16 tcgen05.mma (M = 128 voxels, K = 16)
//                      against smem-resident weights; the accumulators form a ring of three TMEM blocks that the
//                      epilogue warps read, zero and hand back (see the kernel's own header).
//  slab_conv_kernel  : the same sweep with one MMA per tap (27 * Cin/16 per tile, double-buffered TMEM accumulator):
//                      1x1x1 convs (halo = 0: one tap, one 8 x 16 slab per tile) and the 3x3x3 shapes whose five-block
//                      weight image does not fit shared memory.
//  slab_wgrad_kernel : weight gradient.  dW[(a,b,c), ci, co] = sum_v x[v + (a,b,c) - 1, ci] dy[v, co].  Both operands
//                      are voxel-major (MN-major UMMA operands, K = 16 consecutive voxels along w).  The three w-taps
//                      are folded into M as one-voxel shifts of the dy brick (atom stride 32 B; M = 64 MMAs), the three
//                      h-taps and the Cin/16 channel atoms into N as row shifts of the x brick (atom stride 512 B); the
//                      three d-taps are three accumulators.  Accumulators stay in TMEM for the whole (persistent) CTA
//                      and are added to an fp32 scratch once at the end.  Input-channel groups of <= 48 on grid.z.
//
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2..5 = epilogue.
#pragma once
#include <cuda_bf16.h>

#include "igemm_kernels.cuh"
#include "ptx.cuh"

namespace petsyn {

constexpr int kSlabW = 8, kSlabH = 16;                   // output tile of one depth step (128 voxels = MMA M)
constexpr int kSlabWp = kSlabW + 2, kSlabHp = kSlabH + 2;
constexpr int kSlabMaxRing = 8;                          // slabs in flight: 3 in use + prefetch
constexpr int kSlabMaxTaps = 27;

struct alignas(64) SlabParams {
  CUtensorMap a_map;    // input: dims (16 ch, W, atoms, H, D*N); box (16, 10, atoms, 18, 1); 32B swizzle
  CUtensorMap b_map;    // packed weights [rows][tap * kc_pad + c]; box (16, N); 32B swizzle
  CUtensorMap c_map;    // output: dims (C, W, H, D, N); box (N, 8, 16, 1, 1); no swizzle
  const float* bias;
  int32_t tap_off[27];  // per tap: descriptor start-address delta ((dh+1) * atoms * 320 + (dw+1) * 32) >> 4
  int32_t tap_slab[3];  // taps 9g .. 9g+8 read slab (dd + 1) of the three live slabs
  int32_t ntaps, atoms, kc_pad, b_row;
  int32_t block_n, rows;
  int32_t W, H, D, batch;
  int32_t tiles_w, tiles_h, dchunk, nchunks, items;
  int32_t slab_bytes;   // ring pitch (multiple of 1024)
  int32_t ring;         // slabs in the ring: 4 or 8 (power of two)
  int32_t slab_tx;      // bytes one slab load delivers
  int32_t epi_act;
  float epi_slope;
  int32_t reduce;       // 1: add into the destination (bf16 TMA reduction)
  int32_t out_f32;      // 1: fp32 output (the one-channel network head keeps full precision), c_map is an fp32 map
  int32_t halo;         // 1: 3x3x3 conv (27 taps over three 10x18 slabs); 0: 1x1x1 conv (one tap, one 8x16 slab per tile)
  // depth-folded variant (slab_conv3_kernel): the three depth taps of a (dh, dw) pair are stacked in the MMA's N dimension
  int32_t pair_off[9];  // A start-address delta of pair j = (dh, dw): (((dh+1) * atoms) * 320 + (dw+1) * 32) >> 4
  int32_t wtap[9][5];   // packed-weight tap index of B block b of pair j: the tap (dd5[b], dh, dw), dd5 = {+1, 0, -1, +1, 0}
};

// tcgen05.mma with the 64-bit shared-memory descriptors given as (lo, hi) halves: the MMA-issuing thread only ever
// does 32-bit adds on the start-address field.
__device__ __forceinline__ void umma_bf16_split(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate));
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// smem: [weights: ntaps*atoms*N*32][ring: 4*slab_bytes][staging: 2*128*N*2][barriers]
__host__ __device__ inline int slab_smem_bytes(int ntaps, int atoms, int n, int slab_bytes, int ring, int esz = 2) {
  const int wbytes = (ntaps * atoms * n * 32 + 1023) / 1024 * 1024;
  const int stg = (2 * 128 * n * esz + 1023) / 1024 * 1024;
  return wbytes + ring * slab_bytes + stg + 1024 /*barriers + tap table*/ + 1024 /*align*/;
}

template <int ATOMS>
__global__ void __launch_bounds__(192) slab_conv_kernel(const __grid_constant__ SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = p.block_n;
  const int wbytes = (p.ntaps * p.atoms * N * 32 + 1023) / 1024 * 1024;
  const int esz = p.out_f32 ? 4 : 2;
  const int stg_bytes = (2 * 128 * N * esz + 1023) / 1024 * 1024;
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + wbytes;
  const int R = p.ring;
  const uint32_t Rm = uint32_t(R - 1);
  uint8_t* s_stg = s_ring + R * p.slab_bytes;
  uint8_t* tail = s_stg + stg_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [ring]
  uint64_t* empty_bar = full_bar + kSlabMaxRing;                    // [ring]
  uint64_t* w_bar = empty_bar + kSlabMaxRing;                       // [1]
  uint64_t* acc_full = w_bar + 1;                                   // [2]
  uint64_t* acc_empty = acc_full + 2;                               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < R; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(w_bar, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], 4);
    }
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.a_map);
    ptx::prefetch_tmap(&p.b_map);
    ptx::prefetch_tmap(&p.c_map);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  const int G = gridDim.x;
  if (warp == 0) {
    if (ptx::elect_one()) {
      // ---------------- TMA producer: weights once, then the slab stream of every item ----------------
      ptx::mbar_expect_tx(w_bar, uint32_t(p.ntaps * p.atoms * N * 32));
      for (int t = 0; t < p.ntaps; ++t)
        for (int q = 0; q < p.atoms; ++q)
          ptx::tma_load_2d(s_w + (t * p.atoms + q) * N * 32, &p.b_map, w_bar, t * p.kc_pad + q * 16, p.b_row);
      uint32_t seq = 0;
      for (int item = blockIdx.x; item < p.items; item += G) {
        int t = item;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int ch = t % p.nchunks; t /= p.nchunks;
        const int nb = t;
        const int d0 = ch * p.dchunk;
        const int len = min(p.dchunk, p.D - d0);
        for (int s = 0; s < len + 2 * p.halo; ++s, ++seq) {
          const int slot = seq & Rm;
          const uint32_t ph = (seq / uint32_t(R)) & 1;
          ptx::mbar_wait(&empty_bar[slot], ph ^ 1);
          const int d = d0 - p.halo + s;
          const bool oob = d < 0 || d >= p.D;
          ptx::mbar_expect_tx(&full_bar[slot], uint32_t(p.slab_tx));
          // a depth slice outside the volume must read zeros, not the neighbouring sample: push the box out of range in w
          ptx::tma_load_5d(s_ring + slot * p.slab_bytes, &p.a_map, &full_bar[slot], 0, oob ? p.W + 64 : tw * kSlabW - p.halo,
                           0, th * kSlabH - p.halo, oob ? 0 : nb * p.D + d);
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t wp = kSlabW + 2 * p.halo;             // voxels per slab row
      const uint32_t nslab = 1 + 2 * p.halo;               // slabs a tile reads
      const uint64_t a_desc_base = ptx::umma_desc_base(16, uint32_t(ATOMS) * wp * 32, 6);   // K-major, 32B swizzle
      const uint64_t b_desc_base = ptx::umma_desc_base(16, 256, 6);
      const uint32_t a_hi = uint32_t(a_desc_base >> 32), a_lo0 = uint32_t(a_desc_base);
      const uint32_t b_hi = uint32_t(b_desc_base >> 32), b_lo0 = uint32_t(b_desc_base);
      const uint32_t idesc = ptx::umma_idesc_bf16(128, uint32_t(N), 0, 0);
      const uint32_t ring_lo = a_lo0 + (ptx::smem_u32(s_ring) >> 4);
      const uint32_t w_lo = b_lo0 + (ptx::smem_u32(s_w) >> 4);
      const uint32_t slab16 = uint32_t(p.slab_bytes) >> 4;
      const uint32_t bstep = uint32_t(N) * 2;                 // one weight tile (N x 32 B) in 16-byte units
      const uint32_t rshift = (R == 8) ? 3 : 2;
      ptx::mbar_wait(w_bar, 0);
      uint32_t seq = 0, tile = 0;
      for (int item = blockIdx.x; item < p.items; item += G) {
        const int ch = (item / (p.tiles_w * p.tiles_h)) % p.nchunks;
        const int len = min(p.dchunk, p.D - ch * p.dchunk);
        for (int t = 0; t < len; ++t, ++tile) {
          // slabs seq+t .. seq+t+2 must have landed (the first tile of an item waits for all three)
          for (uint32_t s2 = (t == 0 ? 0 : nslab - 1); s2 < nslab; ++s2) {
            const uint32_t q = seq + t + s2;
            ptx::mbar_wait(&full_bar[q & Rm], (q >> rshift) & 1);
          }
          const uint32_t buf = tile & 1;
          ptx::mbar_wait(&acc_empty[buf], ((tile >> 1) & 1) ^ 1);
          ptx::tc_fence_after_sync();
          const uint32_t acc = tmem_base + buf * 64;
          uint32_t slab_lo[3];
#pragma unroll
          for (int s2 = 0; s2 < 3; ++s2) slab_lo[s2] = ring_lo + ((seq + t + s2) & Rm) * slab16;
          uint32_t b_lo = w_lo;
          if (p.halo == 0) {
#pragma unroll
            for (int q = 0; q < ATOMS; ++q) {
              umma_bf16_split(acc, slab_lo[0] + q * (kSlabW * 2), a_hi, b_lo, b_hi, idesc, q != 0 ? 1u : 0u);
              b_lo += bstep;
            }
          } else {
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const int sl = p.tap_slab[g];
            const uint32_t base = sl == 0 ? slab_lo[0] : (sl == 1 ? slab_lo[1] : slab_lo[2]);
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const uint32_t a_lo = base + uint32_t(p.tap_off[g * 9 + j]);
#pragma unroll
              for (int q = 0; q < ATOMS; ++q) {
                umma_bf16_split(acc, a_lo + q * (kSlabWp * 2), a_hi, b_lo, b_hi, idesc, (g | j | q) != 0 ? 1u : 0u);
                b_lo += bstep;
              }
            }
          }
          }
          ptx::umma_commit(&acc_full[buf]);
          ptx::umma_commit(&empty_bar[(seq + t) & Rm]);     // the oldest slab is no longer needed
          if (t == len - 1 && p.halo) {
            ptx::umma_commit(&empty_bar[(seq + t + 1) & Rm]);
            ptx::umma_commit(&empty_bar[(seq + t + 2) & Rm]);
          }
        }
        seq += len + 2 * p.halo;
      }
    }
  } else {
    // ---------------- epilogue warps (2..5): TMEM -> (+bias, act) -> bf16 -> staging -> TMA store ----------------
    const int quad = warp & 3;                 // TMEM lane quarter this warp may read
    const int row = quad * 32 + (tid & 31);    // accumulator row = voxel h * 8 + w of the tile
    const bool leader = (warp == 2) && ((tid & 31) == 0);
    uint32_t tile = 0;
    for (int item = blockIdx.x; item < p.items; item += G) {
      int t = item;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int ch = t % p.nchunks; t /= p.nchunks;
      const int nb = t;
      const int d0 = ch * p.dchunk;
      const int len = min(p.dchunk, p.D - d0);
      for (int dz = 0; dz < len; ++dz, ++tile) {
        const uint32_t buf = tile & 1;
        ptx::mbar_wait(&acc_full[buf], (tile >> 1) & 1);
        ptx::tc_fence_after_sync();
        // the TMA store that last read this staging buffer (two tiles ago) must have finished reading it
        if (leader) tma_store_wait_read_1();
        named_bar_sync(1, 128);
        uint8_t* stg = s_stg + buf * (128 * N * esz) + row * (N * esz);
        const uint32_t taddr = tmem_base + buf * 64 + (uint32_t(quad * 32) << 16);
        for (int c0 = 0; c0 < N; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + uint32_t(c0), v);
          ptx::tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
          if (p.bias != nullptr) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + c0 + i);
          }
          if (p.epi_act != PETSYN_ACT_NONE) {
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], p.epi_act, p.epi_slope);
          }
          if (p.out_f32) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<float4*>(stg + c0 * 4 + q * 16) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          } else {
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
              pk[i] = *reinterpret_cast<uint32_t*>(&b2);
            }
            *reinterpret_cast<uint4*>(stg + c0 * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            *reinterpret_cast<uint4*>(stg + c0 * 2 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          }
        }
        // accumulator drained: hand the TMEM buffer back to the MMA warp
        ptx::tc_fence_before_sync();
        __syncwarp();
        if ((tid & 31) == 0) ptx::mbar_arrive(&acc_empty[buf]);
        ptx::fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (leader) {
          const uint8_t* src = s_stg + buf * (128 * N * esz);
          if (p.reduce)
            ptx::tma_reduce_add_5d(&p.c_map, src, 0, tw * kSlabW, th * kSlabH, d0 + dz, nb);
          else
            ptx::tma_store_5d(&p.c_map, src, 0, tw * kSlabW, th * kSlabH, d0 + dz, nb);
          ptx::tma_store_commit();
        }
      }
    }
    if (leader) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 128);
}

// ------------------------------------------------------------------------------------------------------------
// Depth-folded fprop / dgrad: ONE A read per (slab, (dh, dw) pair, channel atom) feeds all three depth taps.
//
// A slab at depth s contributes to the output tiles at depths s-1, s, s+1 through the taps dd = +1, 0, -1.  Stacking those
// three weight tiles in N (N = 3 * Cout) turns them into ONE tcgen05.mma whose 3 * Cout accumulator columns are three
// DIFFERENT output tiles: a ring of three TMEM blocks, tile t living in block t % 3.  The weights are stored as five
// blocks per pair (dd = +1, 0, -1, +1, 0) so that the three cyclic rotations the ring needs are plain windows.  Every MMA
// accumulates; a tile is complete once the slab after its own has been issued, then the epilogue warps read its block
// AND zero it (tcgen05.st) before the MMA warp may open the tile three steps later in the same block.  Shared-memory
// reads per output tile drop from 27 * (4 KB + B) to 9 * (4 KB + 3 B).
// ------------------------------------------------------------------------------------------------------------
__host__ __device__ inline int slab3_smem_bytes(int atoms, int n, int slab_bytes, int ring, int esz = 2) {
  const int wbytes = (9 * atoms * 5 * n * 32 + 1023) / 1024 * 1024;
  const int stg = (2 * 128 * n * esz + 1023) / 1024 * 1024;
  return wbytes + ring * slab_bytes + stg + 1024 + 1024;
}

template <int ATOMS>
__global__ void __launch_bounds__(192) slab_conv3_kernel(const __grid_constant__ SlabParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int N = p.block_n;
  const int wbytes = (9 * ATOMS * 5 * N * 32 + 1023) / 1024 * 1024;
  const int esz = p.out_f32 ? 4 : 2;
  const int stg_bytes = (2 * 128 * N * esz + 1023) / 1024 * 1024;
  uint8_t* s_w = smem;
  uint8_t* s_ring = smem + wbytes;
  const int R = p.ring;
  const uint32_t Rm = uint32_t(R - 1);
  uint8_t* s_stg = s_ring + R * p.slab_bytes;
  uint8_t* tail = s_stg + stg_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);          // [ring]
  uint64_t* empty_bar = full_bar + kSlabMaxRing;                    // [ring]
  uint64_t* w_bar = empty_bar + kSlabMaxRing;                       // [1]
  uint64_t* acc_full = w_bar + 1;                                   // [3]
  uint64_t* acc_free = acc_full + 3;                                // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 3);
  const uint32_t tmem_cols = 3 * N <= 64 ? 64 : (3 * N <= 128 ? 128 : 256);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

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
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.a_map);
    ptx::prefetch_tmap(&p.b_map);
    ptx::prefetch_tmap(&p.c_map);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  const int G = gridDim.x;
  if (warp == 0) {
    if (ptx::elect_one()) {
      // ---------------- TMA producer: 45 * ATOMS weight tiles once, then one slab per depth step ----------------
      ptx::mbar_expect_tx(w_bar, uint32_t(9 * ATOMS * 5 * N * 32));
      for (int j = 0; j < 9; ++j)
        for (int q = 0; q < ATOMS; ++q)
          for (int b = 0; b < 5; ++b)
            ptx::tma_load_2d(s_w + ((j * ATOMS + q) * 5 + b) * N * 32, &p.b_map, w_bar, p.wtap[j][b] * p.kc_pad + q * 16,
                             p.b_row);
      uint32_t seq = 0;
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
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ---------------- MMA issuer: slab g feeds tiles g, g+1, g+2 (tile t is centred on slab t - 1) ----------------
      const uint64_t a_desc_base = ptx::umma_desc_base(16, uint32_t(ATOMS * kSlabWp * 32), 6);   // K-major, 32B swizzle
      const uint64_t b_desc_base = ptx::umma_desc_base(16, 256, 6);
      const uint32_t a_hi = uint32_t(a_desc_base >> 32), a_lo0 = uint32_t(a_desc_base);
      const uint32_t b_hi = uint32_t(b_desc_base >> 32), b_lo0 = uint32_t(b_desc_base);
      const uint32_t idesc = ptx::umma_idesc_bf16(128, uint32_t(3 * N), 0, 0);
      const uint32_t ring_lo = a_lo0 + (ptx::smem_u32(s_ring) >> 4);
      const uint32_t w_lo = b_lo0 + (ptx::smem_u32(s_w) >> 4);
      const uint32_t slab16 = uint32_t(p.slab_bytes) >> 4;
      const uint32_t blk16 = uint32_t(N) * 2;                  // one weight block (N rows x 32 B) in 16-byte units
      const uint32_t rshift = (R == 8) ? 3 : (R == 4 ? 2 : 1);
      ptx::mbar_wait(w_bar, 0);
      for (int b = 0; b < 3; ++b) ptx::mbar_wait(&acc_free[b], 0);      // the epilogue warps zeroed the three blocks
      ptx::tc_fence_after_sync();
      uint32_t g = 0, r = 0;                                    // r = g % 3
      for (int item = blockIdx.x; item < p.items; item += G) {
        const int ch = (item / (p.tiles_w * p.tiles_h)) % p.nchunks;
        const int len = min(p.dchunk, p.D - ch * p.dchunk);
        for (int s = 0; s < len + 2; ++s, ++g) {
          ptx::mbar_wait(&full_bar[g & Rm], (g >> rshift) & 1);
          if (g > 0) {
            // this slab opens tile g + 2 in the block tile g - 1 used: its epilogue must have read and zeroed it
            const uint32_t t2 = g + 2;
            ptx::mbar_wait(&acc_free[t2 % 3], (t2 / 3) & 1);
          }
          ptx::tc_fence_after_sync();
          const uint32_t slab_lo = ring_lo + (g & Rm) * slab16;
          const uint32_t win = (r == 0 ? 0u : (r == 1 ? 2u : 1u)) * blk16;   // rotation window into the 5 blocks
#pragma unroll
          for (int j = 0; j < 9; ++j) {
            const uint32_t a_lo = slab_lo + uint32_t(p.pair_off[j]);
#pragma unroll
            for (int q = 0; q < ATOMS; ++q)
              umma_bf16_split(tmem_base, a_lo + q * (kSlabWp * 2), a_hi, w_lo + (uint32_t(j * ATOMS + q) * 5) * blk16 + win, b_hi,
                              idesc, 1u);
          }
          ptx::umma_commit(&acc_full[r]);                       // tile g (block g % 3) is complete
          ptx::umma_commit(&empty_bar[g & Rm]);
          r = r == 2 ? 0 : r + 1;
        }
      }
    }
  } else {
    // ---------------- epilogue warps (2..5) ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + (tid & 31);
    const bool leader = (warp == 2) && ((tid & 31) == 0);
    const uint32_t lane_addr = tmem_base + (uint32_t(quad * 32) << 16);
    uint32_t zeros[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) zeros[i] = 0u;
    for (int c0 = 0; c0 < 3 * N; c0 += 16) ptx::tmem_st_32x16(lane_addr + uint32_t(c0), zeros);
    ptx::tmem_st_wait();
    ptx::tc_fence_before_sync();
    __syncwarp();
    if ((tid & 31) == 0)
      for (int b = 0; b < 3; ++b) ptx::mbar_arrive(&acc_free[b]);
    uint32_t g = 0, r = 0, stores = 0;
    for (int item = blockIdx.x; item < p.items; item += G) {
      int t = item;
      const int tw = t % p.tiles_w; t /= p.tiles_w;
      const int th = t % p.tiles_h; t /= p.tiles_h;
      const int ch = t % p.nchunks; t /= p.nchunks;
      const int nb = t;
      const int d0 = ch * p.dchunk;
      const int len = min(p.dchunk, p.D - d0);
      for (int s = 0; s < len + 2; ++s, ++g) {
        // tile g was closed by slab g; it is a real output tile iff s >= 2 (depth d0 + s - 2)
        ptx::mbar_wait(&acc_full[r], (g / 3) & 1);
        ptx::tc_fence_after_sync();
        const bool real = s >= 2;
        const uint32_t buf = stores & 1;
        uint8_t* stg = s_stg + buf * (128 * N * esz) + row * (N * esz);
        if (real) {
          // the TMA store that last read this staging buffer (two stores ago) must have finished reading it
          if (leader) tma_store_wait_read_1();
          named_bar_sync(1, 128);
        }
        const uint32_t taddr = lane_addr + r * uint32_t(N);
        // TMEM handoff first (it is on the MMA warp's critical path): read the whole block into registers, zero it, release
        uint32_t v[4][16];
        if (real) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            if (c * 16 < N) ptx::tmem_ld_32x16(taddr + uint32_t(c * 16), v[c]);
          ptx::tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
          if (c * 16 < N) ptx::tmem_st_32x16(taddr + uint32_t(c * 16), zeros);
        ptx::tmem_st_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if ((tid & 31) == 0) ptx::mbar_arrive(&acc_free[r]);
        if (real) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            if (c * 16 >= N) break;
            const int c0 = c * 16;
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[c][i]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] += __ldg(p.bias + c0 + i);
            }
            if (p.epi_act != PETSYN_ACT_NONE) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], p.epi_act, p.epi_slope);
            }
            if (p.out_f32) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<float4*>(stg + c0 * 4 + q * 16) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            } else {
              uint32_t pk[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
                pk[i] = *reinterpret_cast<uint32_t*>(&b2);
              }
              *reinterpret_cast<uint4*>(stg + c0 * 2) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(stg + c0 * 2 + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
          }
        }
        if (real) {
          ptx::fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (leader) {
            const uint8_t* src = s_stg + buf * (128 * N * esz);
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
    }
    if (leader) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient over slabs.
// ------------------------------------------------------------------------------------------------------------
constexpr int kWgW = 16, kWgH = 16;            // voxel tile of one depth step: 16 (w, = MMA K) x 16 (h) rows
constexpr int kWgMaxRing = 8;

struct alignas(64) SlabWgradParams {
  CUtensorMap x_map;    // dims (16 ch, W, atoms, H, D*N); box (16, 16, atoms, 18, 1); 32B swizzle -> smem [h][atom][w][16]
  CUtensorMap g_map;    // dy co-atom view: dims (16 ch, W, H, D*N); box (16, 18, 16, 1); 32B swizzle -> smem [h][w 18][16]
  float* scratch;       // fp32 [persistent CTA = gridDim.x][ci_groups][co_atoms][3 (c)][48 (i, co)][ncols = 3 (b) * atoms * 16]:
                        // every CTA STORES its accumulators into its own image; the unpack kernel adds the images in CTA order
  int64_t image_floats; // floats of one CTA's image (ci_groups * co_atoms * nacc * 48 * ncols)
  int32_t atoms, co_atoms;   // atoms = 16-channel atoms of x PER channel group (grid.z = group), <= 3
  int32_t W, H, D, batch;
  int32_t tiles_w, tiles_h, dchunk, nchunks, items;
  int32_t xslab_bytes, gslab_bytes;   // ring pitches (multiples of 1024)
  int32_t xslab_tx, gslab_tx;
  int32_t xring, gring;               // ring depths: 4 or 8 / 2 or 4 (powers of two)
  int32_t tmem_cols, acc_stride;      // TMEM allocation (power of two) and column pitch of the three accumulators
  int32_t halo;                       // 1: 3x3x3 (27 taps); 0: 1x1x1 (one tap: no halo, one accumulator, ncols = atoms * 16)
  int32_t m64;                        // 1: M = 64 MMAs (4 dy shift atoms instead of 8: half the A-operand smem reads);
                                      //    accumulator row m then sits in TMEM lane (m / 16) * 32 + m % 16
};

__host__ __device__ inline int slab_wgrad_smem_bytes(int xslab_bytes, int gslab_bytes, int ncols, int xring, int gring) {
  const int stg = (48 * ncols * 4 + 1023) / 1024 * 1024;
  const int ring = xring * xslab_bytes + gring * gslab_bytes;
  return (ring > stg ? ring : stg) + 1024 + 1024;
}

__device__ __forceinline__ void bulk_store(float* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ptx::smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}

// grid: x = persistent CTAs, y = co atom, z = input-channel group (atoms * 16 channels each)
static __global__ void __launch_bounds__(192) slab_wgrad_kernel(const __grid_constant__ SlabWgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int hh = p.halo;
  const int nacc = hh ? 3 : 1;
  const int ncols = nacc * p.atoms * 16;
  const int stg_bytes = (48 * ncols * 4 + 1023) / 1024 * 1024;
  const int XR = p.xring, GR = p.gring;
  const int ring_bytes = XR * p.xslab_bytes + GR * p.gslab_bytes;
  uint8_t* s_x = smem;
  uint8_t* s_g = smem + XR * p.xslab_bytes;
  uint8_t* tail = smem + (ring_bytes > stg_bytes ? ring_bytes : stg_bytes);
  uint64_t* xfull = reinterpret_cast<uint64_t*>(tail);
  uint64_t* xempty = xfull + kWgMaxRing;
  uint64_t* gfull = xempty + kWgMaxRing;
  uint64_t* gempty = gfull + kWgMaxRing;
  uint64_t* done_bar = gempty + kWgMaxRing;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int coa = blockIdx.y;
  const int cig = blockIdx.z;
  const int G = gridDim.x;
  const bool has_work = (int)blockIdx.x < p.items;

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < XR; ++s) { ptx::mbar_init(&xfull[s], 1); ptx::mbar_init(&xempty[s], 1); }
    for (int s = 0; s < GR; ++s) { ptx::mbar_init(&gfull[s], 1); ptx::mbar_init(&gempty[s], 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.x_map);
    ptx::prefetch_tmap(&p.g_map);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, uint32_t(p.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  if (warp == 0) {
    if (ptx::elect_one() && has_work) {
      uint32_t xseq = 0, gseq = 0;
      for (int item = blockIdx.x; item < p.items; item += G) {
        int t = item;
        const int tw = t % p.tiles_w; t /= p.tiles_w;
        const int th = t % p.tiles_h; t /= p.tiles_h;
        const int ch = t % p.nchunks; t /= p.nchunks;
        const int nb = t;
        const int d0 = ch * p.dchunk;
        const int len = min(p.dchunk, p.D - d0);
        // interleave: x slabs d0-1, d0, then per output slice: x slab d+1 and dy slab d
        for (int s = 0; s < len + 2 * hh; ++s) {
          {
            const int slot = xseq & (XR - 1);
            ptx::mbar_wait(&xempty[slot], ((xseq / uint32_t(XR)) & 1) ^ 1);
            const int d = d0 - hh + s;
            const bool oob = d < 0 || d >= p.D;
            ptx::mbar_expect_tx(&xfull[slot], uint32_t(p.xslab_tx));
            ptx::tma_load_5d(s_x + slot * p.xslab_bytes, &p.x_map, &xfull[slot], 0, oob ? p.W + 64 : tw * kWgW,
                             cig * p.atoms, th * kWgH - hh, oob ? 0 : nb * p.D + d);
            ++xseq;
          }
          if (s >= 2 * hh) {
            const int slot = gseq & (GR - 1);
            ptx::mbar_wait(&gempty[slot], ((gseq / uint32_t(GR)) & 1) ^ 1);
            const int d = d0 + s - 2 * hh;
            ptx::mbar_expect_tx(&gfull[slot], uint32_t(p.gslab_tx));
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
                "[%2];" ::"r"(ptx::smem_u32(s_g + slot * p.gslab_bytes)),
                "l"(reinterpret_cast<uint64_t>(&p.g_map)), "r"(ptx::smem_u32(&gfull[slot])), "r"(coa * 16),
                "r"(tw * kWgW - hh), "r"(th * kWgH), "r"(nb * p.D + d)
                : "memory");
            ++gseq;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one() && has_work) {
      // A = dy brick [h][w 18][16 ch]: MN-major 32B swizzle; M atoms (one-voxel shifts) 32 B apart, 8-voxel K groups 256 B
      const uint64_t a_desc_base = ptx::umma_desc_base(32, 256, 6);
      // B = x brick [h][atom][w 16][16 ch]: N atoms ((h-shift, channel atom)) 512 B apart
      const uint64_t b_desc_base = ptx::umma_desc_base(512, 256, 6);
      const uint32_t a_hi = uint32_t(a_desc_base >> 32), b_hi = uint32_t(b_desc_base >> 32);
      const uint32_t idesc = ptx::umma_idesc_bf16(p.m64 ? 64 : 128, uint32_t(ncols), 1, 1);
      const uint32_t x_lo = uint32_t(b_desc_base) + (ptx::smem_u32(s_x) >> 4);
      const uint32_t g_lo = uint32_t(a_desc_base) + (ptx::smem_u32(s_g) >> 4);
      const uint32_t xrow16 = uint32_t(p.atoms) * 32;   // one h row of the x brick (atoms * 512 B) in 16-byte units
      const uint32_t xslab16 = uint32_t(p.xslab_bytes) >> 4, gslab16 = uint32_t(p.gslab_bytes) >> 4;
      const uint32_t XRm = uint32_t(XR - 1), GRm = uint32_t(GR - 1);
      const uint32_t xshift = (XR == 8) ? 3 : 2, gshift = (GR == 4) ? 2 : 1;
      uint32_t xseq = 0, gseq = 0;
      uint32_t started = 0;                             // 0 until the accumulators hold data
      for (int item = blockIdx.x; item < p.items; item += G) {
        const int ch = (item / (p.tiles_w * p.tiles_h)) % p.nchunks;
        const int len = min(p.dchunk, p.D - ch * p.dchunk);
        for (int t = 0; t < len; ++t) {
          for (int s2 = (t == 0 ? 0 : nacc - 1); s2 < nacc; ++s2) {
            const uint32_t q = xseq + t + s2;
            ptx::mbar_wait(&xfull[q & XRm], (q >> xshift) & 1);
          }
          ptx::mbar_wait(&gfull[gseq & GRm], (gseq >> gshift) & 1);
          ptx::tc_fence_after_sync();
          const uint32_t g0 = g_lo + (gseq & GRm) * gslab16;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            if (c >= nacc) break;
            const uint32_t x0 = x_lo + ((xseq + t + c) & XRm) * xslab16;
            const uint32_t acc = tmem_base + c * p.acc_stride;
#pragma unroll
            for (int h = 0; h < kWgH; ++h) {
              // dy row h (18 voxels, starting one voxel left of the tile) x rows h-1..h+1 (= brick rows h..h+2)
              umma_bf16_split(acc, g0 + h * (18 * 2), a_hi, x0 + h * xrow16, b_hi, idesc, h == 0 ? started : 1u);
            }
          }
          started = 1;
          ptx::umma_commit(&gempty[gseq & GRm]);
          ptx::umma_commit(&xempty[(xseq + t) & XRm]);
          if (t == len - 1 && hh) {
            ptx::umma_commit(&xempty[(xseq + t + 1) & XRm]);
            ptx::umma_commit(&xempty[(xseq + t + 2) & XRm]);
          }
          ++gseq;
        }
        xseq += len + 2 * hh;
      }
      ptx::umma_commit(done_bar);
    }
  }
  __syncthreads();
  if (has_work && warp >= 2) {
    ptx::mbar_wait(done_bar, 0);
    ptx::tc_fence_after_sync();
    // rows (lanes) 0..47 hold (i = 2 - a, co); warps 2 and 3 own TMEM lanes 64..127 / 96..127 -> use quads 0 and 1:
    // warp 4 reads lanes 0..31 (quad 0), warp 5 reads lanes 32..63 (quad 1, only 32..47 are useful).
    float* stg = reinterpret_cast<float*>(smem);
    for (int c = 0; c < nacc; ++c) {
      // M = 128: accumulator row m = TMEM lane m (rows 0..47 useful: quads 0 and 1 = warps 4 and 5).
      // M = 64 : row m sits in lane (m / 16) * 32 + m % 16: rows 0..47 = the first 16 lanes of quads 0, 1, 2 (warps 4, 5, 2).
      const int quad = warp & 3;
      const int lane = tid & 31;
      const bool reader = p.m64 ? (quad <= 2) : (quad <= 1);
      if (reader) {
        const int row = p.m64 ? quad * 16 + lane : quad * 32 + lane;
        const bool valid = p.m64 ? lane < 16 : row < 48;
        const uint32_t taddr = tmem_base + c * p.acc_stride + (uint32_t(quad * 32) << 16);
        for (int c0 = 0; c0 < ncols; c0 += 16) {
          uint32_t v[16];
          ptx::tmem_ld_32x16(taddr + uint32_t(c0), v);
          ptx::tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(stg + row * ncols + c0 + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          }
        }
        ptx::fence_proxy_async_smem();
      }
      named_bar_sync(1, 128);
      if (warp == 4 && (tid & 31) == 0) {
        bulk_store(p.scratch + (size_t)blockIdx.x * p.image_floats + ((size_t)((cig * p.co_atoms + coa) * nacc + c) * 48) * ncols,
                   stg, uint32_t(48 * ncols * 4));
        ptx::tma_store_commit();
        ptx::tma_store_wait_read();
      }
      named_bar_sync(1, 128);
    }
    if (warp == 4 && (tid & 31) == 0) ptx::tma_store_wait_all();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, uint32_t(p.tmem_cols));
}

// scratch [cta][ci_group][co_atoms][nacc (c = kd)][48 = (2 - kw) * 16 + co % 16][(kh * apg + q) * 16 + ci % 16], ci atom = group * apg + q
//   -> dw[co][ci][kd][kh][kw]            (k3 = 27; for 1x1x1 convs k3 = 1, nacc = 1 and the only tap sits in row block 0)
// One thread per element of ONE image (coalesced over the image's own layout); the `nimages` per-CTA images are added in CTA
// order, so the weight gradient is reproducible from run to run (no atomics, no memset of the scratch).
static __global__ void __launch_bounds__(256) slab_wgrad_unpack_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                                int cout, int cin, int apg, int k3, int accumulate,
                                                                int nimages, int64_t image_floats,
                                                                const double* __restrict__ dbias_acc,
                                                                float* __restrict__ dbias, int nbias) {
  ptx::pdl_sync();
  if (dbias != nullptr && blockIdx.x == 0)                 // bias gradient: double accumulator -> fp32 gradient slot
    for (int i = threadIdx.x; i < nbias; i += 256) dbias[i] = accumulate ? dbias[i] + (float)dbias_acc[i] : (float)dbias_acc[i];
  const int nacc = k3 == 27 ? 3 : 1;
  const int ncols = nacc * apg * 16;
  const int co_atoms = cout >> 4;
  for (int64_t j = blockIdx.x * 256 + threadIdx.x; j < image_floats; j += (int64_t)gridDim.x * 256) {
    int64_t t = j;
    const int col = (int)(t % ncols); t /= ncols;
    const int row = (int)(t % 48); t /= 48;
    const int kd = (int)(t % nacc); t /= nacc;
    const int coa = (int)(t % co_atoms); t /= co_atoms;
    const int grp = (int)t;
    const int kh = col / (apg * 16), q = (col / 16) % apg, cil = col & 15;
    const int rowblk = row >> 4, col_ = row & 15;
    if (k3 != 27 && rowblk != 0) continue;                 // 1x1x1: only row block 0 holds the tap
    const int kw = k3 == 27 ? 2 - rowblk : 0;
    const int ci = (grp * apg + q) * 16 + cil, co = coa * 16 + col_;
    if (ci >= cin) continue;                                // zero-filled atoms of a short last channel group
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    int i = 0;
    for (; i + 3 < nimages; i += 4) {
      v0 += __ldg(scratch + (int64_t)i * image_floats + j);
      v1 += __ldg(scratch + (int64_t)(i + 1) * image_floats + j);
      v2 += __ldg(scratch + (int64_t)(i + 2) * image_floats + j);
      v3 += __ldg(scratch + (int64_t)(i + 3) * image_floats + j);
    }
    for (; i < nimages; ++i) v0 += __ldg(scratch + (int64_t)i * image_floats + j);
    const float v = (v0 + v1) + (v2 + v3);
    const int64_t o = ((int64_t)co * cin + ci) * k3 + (k3 == 27 ? (kd * 3 + kh) * 3 + kw : 0);
    dw[o] = accumulate ? dw[o] + v : v;
  }
}

}  // namespace petsyn
