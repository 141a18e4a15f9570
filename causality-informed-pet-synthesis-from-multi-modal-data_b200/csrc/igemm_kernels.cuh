// Implicit-GEMM convolution kernels for sm_100a.
//
//  igemm_kernel  : "gather" form used by fprop, dgrad and ConvTranspose.  One CTA computes a 128-voxel x BLOCK_N-channel
//                  output tile.  The 128 voxels are a (box_w x box_h x box_d) brick of the output view, so for every
//                  kernel tap the A operand is ONE 5-D TMA box load from the (possibly stride-2 phase-) view of the
//                  input, shifted by the tap offset; out-of-bounds voxels are zero-filled by the TMA unit, which is the
//                  convolution's zero padding.  B (packed weights, K-major) comes in by 2-D TMA.  Both land in 128B-
//                  swizzled shared memory and feed tcgen05.mma (M=128, N=BLOCK_N, K=16 per instruction) accumulating
//                  fp32 in TMEM.  Epilogue: tcgen05.ld -> (+bias, activation) -> bf16/fp32 -> swizzled smem -> TMA
//                  store (clipped at the tensor edge) or TMA add-reduction (split-K).
//  wgrad_kernel  : "reduce over voxels" form: dW'[r, tap, c] = sum_vox dy[vox, r] * x[vox + tap, c].  Both operands
//                  are voxel-major in memory, i.e. MN-major UMMA operands; K = 64 voxels per pipeline stage.
//
// Warp roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer (warp 1 owns the TMEM
// allocation), then all four warps run the epilogue (warp w reads TMEM lanes 32w..32w+31).
#pragma once
#include <cuda_bf16.h>

#include "ptx.cuh"

namespace petsyn {

constexpr int kMaxViews = 8;
constexpr int kMaxSubs = 8;
constexpr int kMaxTaps = 64;

struct IgemmTap {  // one kernel tap of a sub-problem
  int32_t a_view;  // which A tensor map (phase view) to read
  int32_t dw, dh, dd;  // voxel offset of the A box relative to the output tile origin
};

struct IgemmSub {
  int32_t c_view;      // which C tensor map (output phase view)
  int32_t b_row;       // first row of this sub-problem in the packed weight matrix
  int32_t tap_begin;   // index into the tap table
  int32_t tap_count;
};

struct alignas(64) IgemmParams {
  CUtensorMap a_maps[kMaxViews];
  CUtensorMap c_maps[kMaxViews];
  CUtensorMap b_map;
  IgemmSub subs[kMaxSubs];
  const IgemmTap* taps;   // device table
  const float* bias;      // [rows] or nullptr
  int32_t tiles_w, tiles_h, tiles_d, batch;
  int32_t box_w, box_h, box_d;
  int32_t a_stage_bytes;  // bytes one A box load delivers
  int32_t kc_chunks;      // channel chunks (of KCH) per tap
  int32_t kc_pad;         // padded reduction channels per tap in B
  int32_t rows;           // valid output channels (for bias bounds)
  int32_t epi_act;
  float epi_slope;
  int32_t ksplit;         // CTAs sharing one output tile's K loop (split-K; >1 only with OUT_F32_REDUCE)
  // STATS variant: per (sample, channel) sum and sum of squares of the STORED (bf16) outputs, added to up to two
  // [sample][2][st_c] double accumulators (the GroupNorm / InstanceNorm that consumes the output skips its statistics pass)
  double* st1;
  double* st2;
  int32_t st1_c, st1_off, st2_c, st2_off;
  int32_t ext_w, ext_h, ext_d;   // extent of the output view: rows of a tile beyond it are clipped by the store and not summed
};

enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_F32_REDUCE = 2, OUT_BF16_REDUCE = 3 };

__device__ __forceinline__ float apply_act(float x, int act, float slope) {
  switch (act) {
    case PETSYN_ACT_RELU: return fmaxf(x, 0.f);
    case PETSYN_ACT_LRELU: return x > 0.f ? x : x * slope;
    case PETSYN_ACT_SILU: return x / (1.f + __expf(-x));
    case PETSYN_ACT_TANH: return tanhf(x);
    default: return x;
  }
}

template <int KCH>
struct SwizzleOf {
  static constexpr int kBytes = KCH * 2;                                   // bytes per K-chunk row: 128 / 64 / 32
  static constexpr uint32_t kLayout = (kBytes == 128) ? 2u : (kBytes == 64 ? 4u : 6u);
  static constexpr uint32_t kSbo = 8u * kBytes;                            // 8-row core-matrix group pitch
};

template <int BLOCK_N, int KCH, int STAGES, int OUT_MODE>
struct IgemmCfg {
  static constexpr int kABytes = 128 * KCH * 2;
  static constexpr int kBBytesRaw = BLOCK_N * KCH * 2;
  static constexpr int kBBytes = (kBBytesRaw + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kEsz = (OUT_MODE == OUT_BF16 || OUT_MODE == OUT_BF16_REDUCE) ? 2 : 4;
  // epilogue staging: chunks of kChunkC channels with 128-byte (swizzled) rows when BLOCK_N allows, else one dense chunk
  static constexpr bool kSwz = (BLOCK_N * kEsz) % 128 == 0;
  static constexpr int kChunkC = kSwz ? 128 / kEsz : BLOCK_N;
  static constexpr int kNChunk = BLOCK_N / kChunkC;
  static constexpr int kRowPitch = kChunkC * kEsz;
  static constexpr int kStagingBytes = kNChunk * 128 * kRowPitch;
  static constexpr int kPipeBytes = STAGES * kStageBytes;
  static constexpr int kMainBytes = kPipeBytes > kStagingBytes ? kPipeBytes : kStagingBytes;
  static constexpr int kSmemBytes = kMainBytes + 1024 /*align slack*/ + 2048 /*barriers (256 B) + tap table*/;
  static constexpr uint32_t kTmemCols = BLOCK_N <= 32 ? 32 : (BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256));
};

template <int BLOCK_N, int KCH, int STAGES, int OUT_MODE, bool STATS = false>
__global__ void __launch_bounds__(128) igemm_kernel(const __grid_constant__ IgemmParams p) {
  static_assert(!STATS || (OUT_MODE == OUT_BF16 && (BLOCK_N == 32 || BLOCK_N == 64 || BLOCK_N == 128)),
                "output statistics: bf16 stores, 32 / 64 / 128 channels per tile");
  using Cfg = IgemmCfg<BLOCK_N, KCH, STAGES, OUT_MODE>;
  using Sw = SwizzleOf<KCH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + Cfg::kMainBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                            // [STAGES]
  uint64_t* accum_bar = empty_bar + STAGES;                           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);   // [1]
  IgemmTap* s_taps = reinterpret_cast<IgemmTap*>(tail + 256);         // kMaxTaps x 16 B = 1024 B

  const int tid = threadIdx.x;
  const int warp = tid >> 5;

  const IgemmSub sub = p.subs[blockIdx.z / p.ksplit];
  const int split = blockIdx.z % p.ksplit;
  int t = blockIdx.x;
  const int tw = t % p.tiles_w; t /= p.tiles_w;
  const int th = t % p.tiles_h; t /= p.tiles_h;
  const int td = t % p.tiles_d; t /= p.tiles_d;
  const int nb = t;
  const int w0 = tw * p.box_w, h0 = th * p.box_h, d0 = td * p.box_d;
  const int n0 = blockIdx.y * BLOCK_N;
  const int total_steps = sub.tap_count * p.kc_chunks;
  // balanced partition of the K steps: no split is empty as long as ksplit <= total_steps (the plan guarantees it), so every
  // partial image of the split-K workspace is written in full and needs no clearing
  const int s_begin = (int)(((int64_t)split * total_steps) / p.ksplit);
  const int s_end = (int)(((int64_t)(split + 1) * total_steps) / p.ksplit);
  const int nsteps = s_end - s_begin;
  if (nsteps <= 0) return;   // uniform across the CTA (cannot happen for a planned split)

  for (int i = tid; i < sub.tap_count; i += 128) s_taps[i] = p.taps[sub.tap_begin + i];

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
    ptx::prefetch_tmap(&p.b_map);
    ptx::prefetch_tmap(&p.c_maps[sub.c_view]);
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ---------------- TMA producer ----------------
      const uint32_t tx_bytes = uint32_t(p.a_stage_bytes) + uint32_t(Cfg::kBBytesRaw);
      int tap = s_begin / p.kc_chunks, kc = s_begin - tap * p.kc_chunks;
      for (int s = 0; s < nsteps; ++s) {
        const IgemmTap tp = s_taps[tap];
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
        uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
        uint8_t* b_dst = a_dst + Cfg::kABytes;
        ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
        ptx::tma_load_5d(a_dst, &p.a_maps[tp.a_view], &full_bar[stage], kc * KCH, w0 + tp.dw, h0 + tp.dh, d0 + tp.dd, nb);
        ptx::tma_load_2d(b_dst, &p.b_map, &full_bar[stage], tap * p.kc_pad + kc * KCH, sub.b_row + n0);
        if (++kc == p.kc_chunks) { kc = 0; ++tap; }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ---------------- MMA issuer ----------------
      constexpr uint64_t desc_base = ptx::umma_desc_base(16, Sw::kSbo, Sw::kLayout);
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BLOCK_N, 0, 0);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&full_bar[stage], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < KCH / 16; ++k) {
          ptx::umma_bf16(tmem_base, ptx::umma_desc(desc_base, a_addr + k * 32), ptx::umma_desc(desc_base, b_addr + k * 32),
                         idesc, (s | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);   // frees the smem slot once these MMAs have read it
      }
      ptx::umma_commit(accum_bar);             // accumulator complete
    }
  }
  __syncwarp();

  // ---------------- epilogue: TMEM -> registers -> (bias, act, convert) -> swizzled smem -> TMA store ----------------
  ptx::mbar_wait(accum_bar, 0);
  ptx::tc_fence_after_sync();
  const int row = tid;  // accumulator row == TMEM lane == voxel index inside the box
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
#pragma unroll 1
  for (int c0 = 0; c0 < BLOCK_N; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x16(tmem_base + lane_base + uint32_t(c0), v);
    ptx::tmem_ld_wait();
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
    if (p.bias != nullptr) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int ch = n0 + c0 + i;
        f[i] += (ch < p.rows) ? __ldg(p.bias + ch) : 0.f;
      }
    }
    if (p.epi_act != PETSYN_ACT_NONE) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = apply_act(f[i], p.epi_act, p.epi_slope);
    }
    const int chunk = c0 / Cfg::kChunkC;
    const int cin = c0 % Cfg::kChunkC;
    uint8_t* rowp = smem + chunk * (128 * Cfg::kRowPitch) + row * Cfg::kRowPitch;
    const int j0 = (cin * Cfg::kEsz) >> 4;
    if constexpr (OUT_MODE == OUT_BF16 || OUT_MODE == OUT_BF16_REDUCE) {
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 b2 = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        pk[i] = *reinterpret_cast<uint32_t*>(&b2);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int j = j0 + q;
        const int js = Cfg::kSwz ? (j ^ (row & 7)) : j;
        *reinterpret_cast<uint4*>(rowp + (js << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int j = j0 + q;
        const int js = Cfg::kSwz ? (j ^ (row & 7)) : j;
        *reinterpret_cast<float4*>(rowp + (js << 4)) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      }
    }
  }
  // STATS scratch sits behind the staging tile, in pipeline stages no MMA reads any more (accum_bar has completed)
  uint8_t* s_valid = smem + Cfg::kStagingBytes;                              // [128] row inside the output extent?
  float* s_part = reinterpret_cast<float*>(smem + Cfg::kStagingBytes + 128); // [row groups][BLOCK_N][2]
  if constexpr (STATS) {
    static_assert(Cfg::kStagingBytes + 128 + (256 / BLOCK_N) * BLOCK_N * 2 * 4 <= Cfg::kMainBytes, "no room for the statistics scratch");
    const int bw = row % p.box_w, bh = (row / p.box_w) % p.box_h, bd = row / (p.box_w * p.box_h);
    // (a box may hold fewer than 128 voxels: the accumulator rows behind it are computed from stale shared memory)
    s_valid[row] = (bd < p.box_d && w0 + bw < p.ext_w && h0 + bh < p.ext_h && d0 + bd < p.ext_d) ? 1 : 0;
  }
  ptx::tc_fence_before_sync();
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  if (tid == 0) {
    const CUtensorMap* cmap = &p.c_maps[sub.c_view];
#pragma unroll 1
    for (int chunk = 0; chunk < Cfg::kNChunk; ++chunk) {
      const int ch = n0 + chunk * Cfg::kChunkC;
      if (ch >= p.rows) break;
      const uint8_t* src = smem + chunk * (128 * Cfg::kRowPitch);
      if constexpr (OUT_MODE == OUT_F32_REDUCE)        // split-K: a plain store into this split's own partial image
        ptx::tma_store_5d(cmap, src, ch, w0, h0, d0, nb + split * p.batch);
      else if constexpr (OUT_MODE == OUT_BF16_REDUCE)
        ptx::tma_reduce_add_5d(cmap, src, ch, w0, h0, d0, nb);
      else
        ptx::tma_store_5d(cmap, src, ch, w0, h0, d0, nb);
    }
    ptx::tma_store_commit();
    if constexpr (!STATS) ptx::tma_store_wait_all();
  }
  if constexpr (STATS) {
    // column sums of the staged tile (what the TMA store is writing): thread = (channel pair, group of rows)
    constexpr int kPairs = BLOCK_N / 2, kGroups = 128 / kPairs, kRowsPer = 128 / kGroups;
    const int pair = tid % kPairs, grp = tid / kPairs;
    const int c = 2 * pair;
    const uint8_t* colp = smem + (c / Cfg::kChunkC) * (128 * Cfg::kRowPitch);
    const int cin = c % Cfg::kChunkC;
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 4
    for (int r = grp * kRowsPer; r < (grp + 1) * kRowsPer; ++r) {
      const int j = (cin * 2) >> 4;
      const int js = Cfg::kSwz ? (j ^ (r & 7)) : j;
      const uint32_t w = *reinterpret_cast<const uint32_t*>(colp + r * Cfg::kRowPitch + (js << 4) + ((cin * 2) & 15));
      const bool ok = s_valid[r] != 0;              // a select, not a product: a clipped row may hold NaN
      const float a = ok ? __uint_as_float(w << 16) : 0.f, b = ok ? __uint_as_float(w & 0xffff0000u) : 0.f;
      s0 += a; q0 += a * a;
      s1 += b; q1 += b * b;
    }
    float* pp = s_part + (grp * BLOCK_N + c) * 2;
    pp[0] = s0; pp[1] = q0; pp[2] = s1; pp[3] = q1;
    __syncthreads();
    if (tid < BLOCK_N && n0 + tid < p.rows) {
      float ts = 0.f, tq = 0.f;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) { ts += s_part[(g * BLOCK_N + tid) * 2]; tq += s_part[(g * BLOCK_N + tid) * 2 + 1]; }
      const int ch = n0 + tid;
      if (p.st1 != nullptr) {
        double* d = p.st1 + (int64_t)nb * 2 * p.st1_c + p.st1_off + ch;
        atomicAdd(d, (double)ts);
        atomicAdd(d + p.st1_c, (double)tq);
      }
      if (p.st2 != nullptr) {
        double* d = p.st2 + (int64_t)nb * 2 * p.st2_c + p.st2_off + ch;
        atomicAdd(d, (double)ts);
        atomicAdd(d + p.st2_c, (double)tq);
      }
    }
    if (tid == 0) ptx::tma_store_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad: D[r, c] (128 x BLOCK_N, fp32) = sum over voxels of G[vox, r] * X[vox + tap, c]
// ------------------------------------------------------------------------------------------------------------
struct alignas(64) WgradParams {
  CUtensorMap g_maps[kMaxViews];   // dy views (per sub-problem output view); box = (64 ch, bw, bh, bd, 1)
  CUtensorMap x_maps[kMaxViews];   // x views (per tap a_view);               box = (64 ch, bw, bh, bd, 1)
  CUtensorMap d_map;               // fp32 [rows_total, ldb] scratch; box = (32, 128)
  IgemmSub subs[kMaxSubs];
  const IgemmTap* taps;
  int32_t tiles_w, tiles_h, tiles_d, batch;   // voxel boxes (K blocks) of the sub-problem's output view
  int32_t box_w, box_h, box_d;
  int32_t box_bytes;               // bytes one (64 ch x box) load delivers
  int32_t kc_pad;                  // padded Cin per tap in the scratch matrix
  int32_t ksplit;                  // number of CTAs sharing the voxel reduction
  int32_t r_tiles, c_tiles;
  int32_t image_rows;              // rows of ONE partial image of the scratch (split k adds into rows [k, k+1) * image_rows)
};

// grid: x = tap (within sub) * r_tiles * c_tiles flattened, y = ksplit index, z = sub
template <int BLOCK_N, int STAGES>
__global__ void __launch_bounds__(128) wgrad_kernel(const __grid_constant__ WgradParams p) {
  constexpr int kGBytes = 2 * 64 * 64 * 2;            // A operand: 64 voxels x 128 r  (two 64-channel slabs)
  constexpr int kXBytes = (BLOCK_N / 64) * 64 * 64 * 2;
  constexpr int kStageBytes = kGBytes + kXBytes;
  constexpr int kStagingBytes = (BLOCK_N / 32) * 128 * 128;  // fp32, 32-channel chunks with 128B rows
  constexpr int kPipeBytes = STAGES * kStageBytes;
  constexpr int kMainBytes = kPipeBytes > kStagingBytes ? kPipeBytes : kStagingBytes;
  constexpr uint32_t kTmemCols = BLOCK_N <= 64 ? 64 : (BLOCK_N <= 128 ? 128 : 256);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + kMainBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const IgemmSub sub = p.subs[blockIdx.z];
  int t = blockIdx.x;
  const int ct = t % p.c_tiles; t /= p.c_tiles;
  const int rt = t % p.r_tiles; t /= p.r_tiles;
  const int tap = t;
  if (tap >= sub.tap_count) return;   // sub-problems may have fewer taps than the grid's maximum (uniform exit)
  const IgemmTap tp = p.taps[sub.tap_begin + tap];
  const int r0 = rt * 128, c0 = ct * BLOCK_N;
  const int nboxes = p.tiles_w * p.tiles_h * p.tiles_d * p.batch;
  const int per = (nboxes + p.ksplit - 1) / p.ksplit;
  const int kb_begin = blockIdx.y * per;
  const int kb_end = min(nboxes, kb_begin + per);
  const int nsteps = kb_end - kb_begin;
  if (nsteps <= 0) return;

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  if (warp == 0) {
    if (ptx::elect_one()) {
      const CUtensorMap* gmap = &p.g_maps[sub.c_view];
      const CUtensorMap* xmap = &p.x_maps[tp.a_view];
      const uint32_t tx_bytes = uint32_t(p.box_bytes) * (2 + BLOCK_N / 64);
      for (int s = 0; s < nsteps; ++s) {
        int kb = kb_begin + s;
        const int bw = kb % p.tiles_w; kb /= p.tiles_w;
        const int bh = kb % p.tiles_h; kb /= p.tiles_h;
        const int bd = kb % p.tiles_d; kb /= p.tiles_d;
        const int nb = kb;
        const int w0 = bw * p.box_w, h0 = bh * p.box_h, d0 = bd * p.box_d;
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
        uint8_t* g_dst = smem + stage * kStageBytes;
        uint8_t* x_dst = g_dst + kGBytes;
        ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
        ptx::tma_load_5d(g_dst, gmap, &full_bar[stage], r0, w0, h0, d0, nb);
        ptx::tma_load_5d(g_dst + 8192, gmap, &full_bar[stage], r0 + 64, w0, h0, d0, nb);
#pragma unroll
        for (int j = 0; j < BLOCK_N / 64; ++j)
          ptx::tma_load_5d(x_dst + j * 8192, xmap, &full_bar[stage], c0 + j * 64, w0 + tp.dw, h0 + tp.dh, d0 + tp.dd, nb);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // MN-major, 128B swizzle: 64 contiguous channels per voxel row (128 B), 8 voxel rows per 1024 B group (SBO),
      // next 64-channel slab 8192 B further (LBO).
      constexpr uint64_t desc_base = ptx::umma_desc_base(8192, 1024, 2);
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, BLOCK_N, 1, 1);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&full_bar[stage], ph);
        ptx::tc_fence_after_sync();
        const uint32_t g_addr = ptx::smem_u32(smem + stage * kStageBytes);
        const uint32_t x_addr = g_addr + kGBytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 16 voxels per MMA = two 8-row groups = 2048 B
          ptx::umma_bf16(tmem_base, ptx::umma_desc(desc_base, g_addr + k * 2048),
                         ptx::umma_desc(desc_base, x_addr + k * 2048), idesc, (s | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
      }
      ptx::umma_commit(accum_bar);
    }
  }
  __syncwarp();

  ptx::mbar_wait(accum_bar, 0);
  ptx::tc_fence_after_sync();
  const int row = tid;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
#pragma unroll 1
  for (int cc = 0; cc < BLOCK_N; cc += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x16(tmem_base + lane_base + uint32_t(cc), v);
    ptx::tmem_ld_wait();
    const int chunk = cc / 32;
    uint8_t* rowp = smem + chunk * (128 * 128) + row * 128;
    const int j0 = ((cc % 32) * 4) >> 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int js = (j0 + q) ^ (row & 7);
      *reinterpret_cast<uint4*>(rowp + (js << 4)) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
  }
  ptx::tc_fence_before_sync();
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
  if (tid == 0) {
#pragma unroll 1
    for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
      // split-private partial image: the only non-zero contribution to an element of image k comes from ONE CTA (tiles
      // overlap only where the operands are zero-filled), so the reduction order cannot change the sum
      ptx::tma_reduce_add_2d(&p.d_map, smem + chunk * (128 * 128), tap * p.kc_pad + c0 + chunk * 32,
                             int(blockIdx.y) * p.image_rows + sub.b_row + r0);
    }
    ptx::tma_store_commit();
    ptx::tma_store_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------------------
// wgrad for SMALL channel counts (Cin, Cout <= 64, multiples of 16): the AttenUNet full-resolution layers.
// With 16..64 channels a (Cout x Cin) weight-gradient tile would fill 1/32..1/2 of a 128 x 64 MMA.  Instead the kernel
// taps are folded into the MMA's M dimension:   D[(tap, ci), co] = sum_vox X_tap[vox, ci] * dY[vox, co]
// A operand = TPM = 128/Cin shifted boxes of x side by side (16-channel MN-major atoms, 32-byte swizzle),
// B operand = dy (Cout/16 atoms).  One K step = 64 voxels; split over `ksplit` CTAs; fp32 tiles are TMA-add-reduced into
// a scratch laid out [(sub, m_tile) * 128 + tap_local * Cin + ci][co].
// ------------------------------------------------------------------------------------------------------------
struct alignas(64) WgradSmallParams {
  CUtensorMap g_maps[kMaxViews];   // dy views; box = (16 ch, bw, bh, bd, 1), 32B swizzle
  CUtensorMap x_maps[kMaxViews];   // x views;  box = (16 ch, bw, bh, bd, 1), 32B swizzle
  CUtensorMap d_map;               // fp32 [rows_total, npad] scratch; box = (npad, 128), no swizzle
  IgemmSub subs[kMaxSubs];
  const IgemmTap* taps;
  int32_t tiles_w, tiles_h, tiles_d, batch;
  int32_t box_w, box_h, box_d;
  int32_t cin, cin_atoms;          // Cin, Cin/16
  int32_t n_atoms;                 // Cout/16 (padded)
  int32_t tpm;                     // taps per M tile
  int32_t m_tiles;                 // M tiles per sub-problem (grid.x)
  int32_t ksplit;
  int32_t image_rows;              // rows of ONE partial image of the scratch (split k adds into image k)
};

template <int STAGES>
__global__ void __launch_bounds__(128) wgrad_small_kernel(const __grid_constant__ WgradSmallParams p) {
  constexpr int kAtom = 64 * 32;                 // one [64 voxels x 16 channels] box: 2 KB
  constexpr int kABytes = 8 * kAtom;             // 128 M rows = 8 atoms
  constexpr int kBBytes = 4 * kAtom;             // up to N = 64
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr int kStagingBytes = 128 * 64 * 4;    // fp32 [128][<=64]
  constexpr int kPipeBytes = STAGES * kStageBytes;
  constexpr int kMainBytes = kPipeBytes > kStagingBytes ? kPipeBytes : kStagingBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tail = smem + kMainBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const IgemmSub sub = p.subs[blockIdx.z];
  const int mt = blockIdx.x;
  const int tap0 = mt * p.tpm;
  const int ntaps = min(p.tpm, sub.tap_count - tap0);
  if (ntaps <= 0) return;
  const int nboxes = p.tiles_w * p.tiles_h * p.tiles_d * p.batch;
  const int per = (nboxes + p.ksplit - 1) / p.ksplit;
  const int kb_begin = blockIdx.y * per;
  const int nsteps = min(nboxes, kb_begin + per) - kb_begin;
  if (nsteps <= 0) return;
  const int npad = p.n_atoms * 16;

  if (warp == 0 && ptx::elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(accum_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_sync();       // everything above is on-chip set-up: it overlaps the tail of the previous kernel

  if (warp == 0) {
    if (ptx::elect_one()) {
      const CUtensorMap* gmap = &p.g_maps[sub.c_view];
      const uint32_t tx_bytes = uint32_t(kAtom) * uint32_t(ntaps * p.cin_atoms + p.n_atoms);
      for (int s = 0; s < nsteps; ++s) {
        int kb = kb_begin + s;
        const int bw = kb % p.tiles_w; kb /= p.tiles_w;
        const int bh = kb % p.tiles_h; kb /= p.tiles_h;
        const int bd = kb % p.tiles_d; kb /= p.tiles_d;
        const int nb = kb;
        const int w0 = bw * p.box_w, h0 = bh * p.box_h, d0 = bd * p.box_d;
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[stage], ph ^ 1);
        uint8_t* a_dst = smem + stage * kStageBytes;
        uint8_t* b_dst = a_dst + kABytes;
        ptx::mbar_expect_tx(&full_bar[stage], tx_bytes);
        for (int t = 0; t < ntaps; ++t) {
          const IgemmTap tp = p.taps[sub.tap_begin + tap0 + t];
          for (int c = 0; c < p.cin_atoms; ++c)
            ptx::tma_load_5d(a_dst + (t * p.cin_atoms + c) * kAtom, &p.x_maps[tp.a_view], &full_bar[stage], c * 16,
                             w0 + tp.dw, h0 + tp.dh, d0 + tp.dd, nb);
        }
        for (int c = 0; c < p.n_atoms; ++c)
          ptx::tma_load_5d(b_dst + c * kAtom, gmap, &full_bar[stage], c * 16, w0, h0, d0, nb);
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // MN-major, 32B swizzle: 16 contiguous channels per voxel row (32 B), 8 voxel rows per 256 B group (SBO),
      // next 16-channel atom 2048 B further (LBO)
      constexpr uint64_t desc_base = ptx::umma_desc_base(kAtom, 256, 6);
      const uint32_t idesc = ptx::umma_idesc_bf16(128, uint32_t(npad), 1, 1);
      for (int s = 0; s < nsteps; ++s) {
        const int stage = s % STAGES;
        const uint32_t ph = (s / STAGES) & 1;
        ptx::mbar_wait(&full_bar[stage], ph);
        ptx::tc_fence_after_sync();
        const uint32_t a_addr = ptx::smem_u32(smem + stage * kStageBytes);
        const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {   // 16 voxels per MMA = two 8-row groups = 512 B
          ptx::umma_bf16(tmem_base, ptx::umma_desc(desc_base, a_addr + k * 512), ptx::umma_desc(desc_base, b_addr + k * 512),
                         idesc, (s | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);
      }
      ptx::umma_commit(accum_bar);
    }
  }
  __syncwarp();

  ptx::mbar_wait(accum_bar, 0);
  ptx::tc_fence_after_sync();
  const int row = tid;
  const uint32_t lane_base = uint32_t(warp * 32) << 16;
  float* stg = reinterpret_cast<float*>(smem);
  for (int cc = 0; cc < npad; cc += 16) {
    uint32_t v[16];
    ptx::tmem_ld_32x16(tmem_base + lane_base + uint32_t(cc), v);
    ptx::tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(stg + row * npad + cc + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
  ptx::tc_fence_before_sync();
  ptx::fence_proxy_async_smem();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, 64);
  if (tid == 0) {
    ptx::tma_reduce_add_2d(&p.d_map, smem, 0, int(blockIdx.y) * p.image_rows + (int(blockIdx.z) * p.m_tiles + mt) * 128);
    ptx::tma_store_commit();
    ptx::tma_store_wait_all();
  }
}

}  // namespace petsyn
