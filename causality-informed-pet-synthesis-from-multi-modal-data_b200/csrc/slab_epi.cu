// Instantiations of the fused-epilogue slab convolution (slab_epi_kernel.cuh) -- a translation unit of its own so that the
// kernel variants compile beside conv_plan.cu instead of inside it.
#include "common.h"
#include "slab_epi_kernel.cuh"

namespace petsyn {

namespace {

template <int ATOMS, int NB, int FLAGS>
int occupancy_of(int smem) {
  auto kern = slab_conv3_epi_kernel<ATOMS, NB, FLAGS>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024) != cudaSuccess) return 0;
    attr_set = true;
  }
  // resident CTAs per SM from the two resources that bind here: 227 KB of shared memory (1 KB reserved per CTA) and the 64 K
  // registers (allocated per warp in units of 512); the runtime's occupancy query answers for the default carve-out, not
  // for the one the launch will get
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) return 0;
  const int regs_per_warp = (fa.numRegs * 32 + 511) / 512 * 512;
  const int by_regs = 65536 / (regs_per_warp * 6);
  const int by_smem = (227 * 1024) / (smem + 1024);
  return by_regs < by_smem ? by_regs : by_smem;
}

template <int ATOMS, int NB, int FLAGS>
int32_t launch_of(const SlabParams& p, const SlabEpi& e, int grid, int smem, cudaStream_t st) {
  auto kern = slab_conv3_epi_kernel<ATOMS, NB, FLAGS>;
  static bool attr_set = false;
  if (!attr_set) {
    PETSYN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
  PETSYN_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(192), smem, st, p, e));
  return check_launch("slab_conv3_epi_kernel");
}

}  // namespace

#define PETSYN_EPI_FOR_EACH(X)                                                                                     \
  X(1, 1, 0) X(2, 1, 0) X(3, 1, 0) X(1, 2, 0) X(2, 2, 0) X(3, 2, 0)                                                  \
  X(1, 1, kEpiStats) X(1, 1, kEpiSide) X(1, 1, kEpiSide | kEpiStats) X(1, 1, kEpiNormReduce)                       \
  X(2, 1, kEpiStats) X(2, 1, kEpiSide) X(2, 1, kEpiSide | kEpiStats) X(2, 1, kEpiNormReduce)                       \
  X(3, 1, kEpiStats) X(3, 1, kEpiSide) X(3, 1, kEpiSide | kEpiStats) X(3, 1, kEpiNormReduce)                       \
  X(1, 2, kEpiStats) X(1, 2, kEpiSide) X(1, 2, kEpiSide | kEpiStats) X(1, 2, kEpiNormReduce)                       \
  X(2, 2, kEpiStats) X(2, 2, kEpiSide) X(2, 2, kEpiSide | kEpiStats) X(2, 2, kEpiNormReduce)                       \
  X(3, 2, kEpiStats) X(3, 2, kEpiSide) X(3, 2, kEpiSide | kEpiStats) X(3, 2, kEpiNormReduce)

int slab3_epi_occupancy(int atoms, int nb, int flags, int smem) {
#define X(A, B, F) if (atoms == A && nb == B && flags == (F)) return occupancy_of<A, B, (F)>(smem);
  PETSYN_EPI_FOR_EACH(X)
#undef X
  return 0;
}

int32_t slab3_epi_launch(int atoms, int nb, int flags, const SlabParams& p, const SlabEpi& e, int grid, int smem,
                         cudaStream_t st) {
#define X(A, B, F) if (atoms == A && nb == B && flags == (F)) return launch_of<A, B, (F)>(p, e, grid, smem, st);
  PETSYN_EPI_FOR_EACH(X)
#undef X
  return fail(PETSYN_EINVAL, "fused epilogue: no kernel for %d input atoms, %d output channels, flags %d", atoms, nb * 16, flags);
}

}  // namespace petsyn
