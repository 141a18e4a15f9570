/*
 * libpetsyn -- C ABI of the B200-native (sm_100a) kernels behind the 3D T1->PET generator hot path.
 *
 * The reference (jessyblues/Causality-Informed-PET-Synthesis-from-Multi-modal-Data) is pure PyTorch and has no
 * FFI/plugin layer of its own: its "operator interface" for this path is the set of ATen/cuDNN calls made by the
 * nn.Module graphs in
 *     unet/utils/unet_model.py:37-99            (Conv3d k4 s2 p1, Upsample x2 + Conv3d k3 p1, BatchNorm3d,
 *                                                LeakyReLU/ReLU/Tanh, skip concat)
 *     unet/utils/atten_unet_model.py:565-662    (GroupNorm + SiLU + Conv3d k3, AvgPool/nearest resampling)
 *     bl_methods/BMGAN/bmgan_model.py:12-144    (Conv3d k3 s1/s2, ConvTranspose3d k4 s2 p1, InstanceNorm3d,
 *                                                LeakyReLU/PReLU)
 *     unet/scripts/train_unet.py:106,149,155    (L1 / LSGAN-MSE losses), train_unify_causal_gen.py:57-73 (KL)
 * Every entry point below names the reference call it replaces.  A maintainer binds them from Python with ctypes
 * (see INTEGRATION.md); the package's host layer (`_cabi.py`) is exactly such a binding.
 *
 * Conventions
 *   - All functions return 0 on success or a negative PETSYN_E* code; petsyn_last_error() returns a thread-local
 *     human-readable message for the last failure on the calling thread.
 *   - The caller owns every buffer (inputs, outputs, workspaces); the library allocates no tensor memory.  A plan
 *     object owns only small device-side tables (tap programs) and its cached TMA descriptors.
 *   - All work is enqueued on the cudaStream_t passed in (as void*); no call synchronises the device.
 *   - Activations are channels-last (N, D, H, W, C) bf16; a tensor may be a channel slice [coff, coff+C) of a wider
 *     buffer with `cstride` channels per voxel -- this is how skip concatenation is made copy-free.
 *   - Master weights are fp32 in PyTorch layout (Cout, Cin, kD, kH, kW); kernels consume packed bf16 copies made by
 *     petsyn_conv_pack_weights().
 *   - There is no CPU implementation behind this ABI.
 */
#ifndef PETSYN_H_
#define PETSYN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PETSYN_VERSION 100

/* error codes */
#define PETSYN_OK 0
#define PETSYN_EINVAL (-1)   /* bad descriptor / unsupported shape (Python layer raises ValueError) */
#define PETSYN_ECUDA (-2)    /* CUDA runtime/driver error (RuntimeError) */
#define PETSYN_ENOMEM (-3)   /* workspace too small */

int32_t petsyn_version(void);
const char* petsyn_last_error(void);
/* Number of CUDA kernels this library has launched in the calling process (monotonic; for benchmarks' bookkeeping). */
uint64_t petsyn_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Convolution family (implicit GEMM on tcgen05 tensor cores, TMA-fed, fp32 accumulation in TMEM)
 * ---------------------------------------------------------------------------------------------------------- */

/* operator kinds */
#define PETSYN_OP_CONV 0    /* nn.Conv3d(k, stride 1|2, pad)                  unet_model.py:47, bmgan_model.py:34-51 */
#define PETSYN_OP_UPCONV 1  /* nn.Upsample(scale_factor=2) -> nn.Conv3d(k3,p1)  unet_model.py:59-60,71-72,81-82     */
#define PETSYN_OP_CONVT 2   /* nn.ConvTranspose3d(k4, s2, p1)                 bmgan_model.py:57-64                  */

/* epilogue activations (fused after bias) */
#define PETSYN_ACT_NONE 0
#define PETSYN_ACT_RELU 1
#define PETSYN_ACT_LRELU 2   /* slope given separately */
#define PETSYN_ACT_SILU 3
#define PETSYN_ACT_TANH 4
#define PETSYN_ACT_PRELU 5   /* nn.PReLU() with one learnable slope read from device memory (normact descriptor only) */

typedef struct petsyn_conv_desc {
  int32_t op;                 /* PETSYN_OP_* */
  int32_t n, d, h, w;         /* dims of the stored input x (for UPCONV: before the nearest x2 upsample) */
  int32_t cin, cout;          /* multiples of 8 (TMA 16-byte stride rule) */
  int32_t ksize, stride, pad; /* CONV: k in {1,3,4}, stride in {1,2}; UPCONV: 3,1,1; CONVT: 4,2,1 */
  int32_t x_cstride, x_coff;  /* channel pitch / offset of x inside its NDHWC buffer (bf16) */
  int32_t y_cstride, y_coff;  /* same for the forward output y (bf16) */
  int32_t dy_cstride, dy_coff;/* gradient w.r.t. y (bf16) */
  int32_t dx_cstride, dx_coff;/* gradient w.r.t. x (bf16) */
  int32_t epi_act;            /* PETSYN_ACT_* applied to y in the fprop epilogue (after bias) */
  float epi_slope;            /* LeakyReLU slope */
  int32_t y_fp32;             /* != 0: the forward output y is fp32 instead of bf16 (pitches still in elements) */
} petsyn_conv_desc;

typedef struct petsyn_conv_plan petsyn_conv_plan;

/* Validates the descriptor, chooses tiles, builds the device-side tap programs.  ValueError-class failures
 * (PETSYN_EINVAL) mirror the reference's channel/shape checks (atten_unet_model.py:502-506,545-546). */
int32_t petsyn_conv_plan_create(const petsyn_conv_desc* desc, petsyn_conv_plan** plan);
void petsyn_conv_plan_destroy(petsyn_conv_plan* plan);

/* Output spatial dims of the forward op. */
int32_t petsyn_conv_out_dims(const petsyn_conv_plan* plan, int32_t* od, int32_t* oh, int32_t* ow);
/* Direct-convolution FLOPs of one forward call (2*M*Cout*Cin*k^3) and the FLOPs the kernel executes
 * (smaller for UPCONV, whose 27 taps collapse to 8 per output phase). */
int32_t petsyn_conv_flops(const petsyn_conv_plan* plan, double* algorithmic, double* executed);

/* Sizes (bytes) of the packed bf16 weight images and of the fp32 wgrad scratch. */
size_t petsyn_conv_packed_fprop_bytes(const petsyn_conv_plan* plan);
size_t petsyn_conv_packed_dgrad_bytes(const petsyn_conv_plan* plan);
size_t petsyn_conv_wgrad_scratch_bytes(const petsyn_conv_plan* plan);

/* Which kernel family a pass of this plan runs on (for profiling / reporting): pass 0 = fprop, 1 = dgrad, 2 = wgrad.
 * Returns 0 = gather-form tcgen05 implicit GEMM (igemm_kernel / wgrad_kernel), 1 = slab kernels (smem-resident halo
 * slabs, small channel counts), 2 = wgrad_small_kernel; -1 on a bad argument. */
int32_t petsyn_conv_kernel_path(const petsyn_conv_plan* plan, int32_t pass);

/* Deep layers with few output voxels (M = 252 at the U-Net bottleneck) split their K loop over several CTAs that
 * add-reduce fp32 partial tiles (TMA reduction) into a caller-owned workspace; 0 bytes when no split is planned.
 * The workspace (shared by fprop and dgrad, which never overlap on one stream) must be set before the first call. */
size_t petsyn_conv_workspace_bytes(const petsyn_conv_plan* plan);
int32_t petsyn_conv_set_workspace(petsyn_conv_plan* plan, void* workspace, size_t bytes);

/* fp32 (Cout,Cin,k,k,k) [CONVT: (Cin,Cout,k,k,k)] -> packed bf16 GEMM operands.  Either destination may be NULL. */
int32_t petsyn_conv_pack_weights(petsyn_conv_plan* plan, const float* w, void* packed_fprop, void* packed_dgrad,
                                 void* stream);

/* Batched weight packing: every weight tensor of a network in a handful of launches (one per kernel-volume / layout
 * class) instead of one launch per tensor.  Entry i packs weights[i] for plans[i] into packed_fprop[i] and, when not
 * NULL, packed_dgrad[i].  All pointers are captured at creation and must stay valid (flat parameter arenas).
 * No reference counterpart: the reference re-reads nn.Parameter storage through cuDNN every step. */
typedef struct petsyn_pack_batch petsyn_pack_batch;
int32_t petsyn_pack_batch_create(int32_t n, petsyn_conv_plan* const* plans, const float* const* weights,
                                 void* const* packed_fprop, void* const* packed_dgrad, petsyn_pack_batch** out);
int32_t petsyn_pack_batch_run(petsyn_pack_batch* batch, void* stream);
void petsyn_pack_batch_destroy(petsyn_pack_batch* batch);

/* y = act(conv(x, W) + bias).  Replaces F.conv3d / F.conv_transpose3d (+ the preceding F.interpolate for UPCONV).
 * bias may be NULL (BatchNorm'd convs are bias-free, unet_model.py:42-45). */
int32_t petsyn_conv_fprop(petsyn_conv_plan* plan, const void* x, const void* packed_fprop, const float* bias, void* y,
                          void* stream);
/* dx = conv_backward_input(dy, W).  Replaces cuDNN dgrad (+ the upsample's backward for UPCONV). */
int32_t petsyn_conv_dgrad(petsyn_conv_plan* plan, const void* dy, const void* packed_dgrad, void* dx, void* stream);
/* dx += conv_backward_input(dy, W): gradient fan-in for tensors with several consumers (dense concatenation and
 * residual sums in bmgan_model.py:12-23).  The tile epilogue add-reduces through the TMA unit (bf16). */
int32_t petsyn_conv_dgrad_accumulate(petsyn_conv_plan* plan, const void* dy, const void* packed_dgrad, void* dx,
                                     void* stream);
/* Fused epilogues of the small-channel 3x3x3 convolutions (the full- and half-resolution ResnetBlock convs of AttenUNet,
 * atten_unet_model.py:641-662): passes over the convolution's OUTPUT that would otherwise be HBM-bound kernels of their own are
 * done on the tile while it is in registers.
 *   side          fprop: added to the output -- the residual sum `conv2(...) + skip_connection(x)` (:662);
 *                 dgrad: z, the input of the normalisation in front of this convolution (a = act(norm(z)), :644-645)
 *   stats1/2      fprop: [sample][2][stats_c] double accumulators += (sum y, sum y^2) of the stored bf16 output at channel offset
 *                 stats_coff -- the statistics of the GroupNorm that reads the output (norm2 after conv1, norm1 of the next block)
 *   norm_*, bsums dgrad: the reduction pass of that normalisation's backward, bsums[sample][2][cin] += (sum g, sum g * zhat) with
 *                 g = dx * act'(z * scale + shift); petsyn_normact_bwd then runs with `sums_precomputed` = 1 (apply pass only)
 * Available when petsyn_conv_epilogue_supported() says so (depth-folded slab kernel, 16 or 32 output channels of the pass);
 * the plain entry points stay valid for every plan. */
typedef struct petsyn_conv_epilogue {
  const void* side;              /* bf16 NDHWC, spatial dims of the pass's output */
  int32_t side_cstride, side_coff;
  int32_t add_side;              /* fprop: y += side */
  double* stats1; int32_t stats1_c, stats1_coff;
  double* stats2; int32_t stats2_c, stats2_coff;
  const float* norm_scale;       /* dgrad: [sample][cin] each */
  const float* norm_shift;
  const float* norm_mean;
  const float* norm_rstd;
  int32_t norm_act;
  float norm_slope;
  double* bsums;
} petsyn_conv_epilogue;
/* pass: 0 = fprop, 1 = dgrad.  1 if petsyn_conv_{fprop,dgrad}_epi can run this plan's pass with any epilogue, 2 if only with
 * statistics targets (stats1 / stats2; the gather-form kernel, forward pass), else 0. */
int32_t petsyn_conv_epilogue_supported(const petsyn_conv_plan* plan, int32_t pass);
int32_t petsyn_conv_fprop_epi(petsyn_conv_plan* plan, const void* x, const void* packed_fprop, const float* bias, void* y,
                              const petsyn_conv_epilogue* epi, void* stream);
int32_t petsyn_conv_dgrad_epi(petsyn_conv_plan* plan, const void* dy, const void* packed_dgrad, void* dx,
                              const petsyn_conv_epilogue* epi, void* stream);
/* dw (fp32, PyTorch layout) = conv_backward_weight(x, dy); optional dbias (fp32 [cout]) = sum(dy).
 * `scratch` must hold petsyn_conv_wgrad_scratch_bytes(); it is zeroed and reduced into by the kernel.
 * accumulate != 0 adds into dw instead of overwriting (autograd .grad accumulation). */
int32_t petsyn_conv_wgrad(petsyn_conv_plan* plan, const void* x, const void* dy, void* scratch, float* dw,
                          int32_t accumulate, void* stream);
/* The same, and -- in the launch that lays the weight gradient out -- the bias gradient: dbias[0:nbias] (+)= dbias_acc[0:nbias]
 * (the double accumulator that petsyn_colsum / a normalisation's backward filled earlier on this stream -> the fp32 gradient
 * slot; added when `accumulate`).  Saves a conversion launch per convolution. */
int32_t petsyn_conv_wgrad_bias(petsyn_conv_plan* plan, const void* x, const void* dy, void* scratch, float* dw,
                               int32_t accumulate, const double* dbias_acc, float* dbias, int32_t nbias, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Edge layers of the pix2pix U-Net (Cin = 1 / Cout = 1): small layout kernels that turn them into 1x1x1 GEMMs for the
 * conv family above, so that they too run on the tensor cores.
 * ---------------------------------------------------------------------------------------------------------- */

/* First layer, Conv3d(1 -> C, k4 s2 p1) on the fp32 NCDHW network input (unet_model.py:47,62): explicit im2col of the
 * single-channel volume.  x fp32 [n,d,h,w] -> patches bf16 [n*(d/2)*(h/2)*(w/2), 64] (tap = (kd*4+kh)*4+kw, zero
 * padded at the borders).  The conv itself is then petsyn_conv_fprop with k=1, cin=64, and its weight gradient
 * petsyn_conv_wgrad on the same patches. */
int32_t petsyn_stem_im2col_k4s2(const float* x, void* patches, int32_t n, int32_t d, int32_t h, int32_t w, void* stream);

/* Backward-data of the first layer (needed when a generator's output feeds the PatchGAN discriminator, whose first
 * conv is exactly this stem): dpatches bf16 [n*(d/2)*(h/2)*(w/2), 64] -> dx fp32 [n,d,h,w] (col2im, gather form). */
int32_t petsyn_stem_col2im_k4s2(const void* dpatches, float* dx, int32_t n, int32_t d, int32_t h, int32_t w, void* stream);

/* BMGAN generator input: cat([t1, z.view(N,nz,1,1,1).expand(...)], 1) (bmgan_model.py:76-79) written as NDHWC bf16
 * with the 1+nz channels zero-padded to cpad.  x fp32 [n, rows_per_sample]; zvec fp32 [n, nz]. */
int32_t petsyn_concat_latent(const float* x, const float* zvec, void* out, int64_t rows_per_sample, int32_t n,
                             int32_t nz, int32_t cpad, void* stream);
/* One-channel heads (generator output conv, PatchGAN final conv) run with Cout padded to cpad:
 * y[r] = src[r, 0] (fp32), and the matching gradient packer dz[r, 0] = dy[r] * (tanh_out ? 1 - y[r]^2 : 1), rest 0
 * (bf16) -- the Tanh of bmgan_model.py:69 is applied in the conv epilogue, its derivative here. */
int32_t petsyn_take_channel0(const float* src, float* y, int64_t rows, int32_t cpad, void* stream);
int32_t petsyn_put_channel0_grad(const float* y, const float* dy, void* dz, int64_t rows, int32_t cpad,
                                 int32_t tanh_out, void* stream);

/* Last layer, Upsample x2 -> Conv3d(C -> 1, k3 p1) -> Tanh (unet_model.py:59-64).  The 27-tap conv on the upsampled
 * grid is computed as a per-SOURCE-voxel projection proj[s][k] = <x[s,:], W[k,:]> (a k=1 conv with cout=32, fp32
 * output, done by petsyn_conv_fprop) followed by this 27-term gather: y[o] = tanh(sum_k proj[(o+k-1)>>1][k]).
 * proj fp32 [n*d*h*w, 32]; y fp32 [n,1,2d,2h,2w]. */
int32_t petsyn_head_gather_tanh(const float* proj, float* y, int32_t n, int32_t d, int32_t h, int32_t w, void* stream);
/* Backward of the gather: dproj[s][k] = sum_{o: (o+k-1)>>1 == s} dy[o]*(1 - y[o]^2), written as bf16 rows of 64
 * (columns 27..63 zero) so that it feeds petsyn_conv_fprop / petsyn_conv_wgrad directly. */
int32_t petsyn_head_scatter_bwd(const float* y, const float* dy, void* dproj, int32_t n, int32_t d, int32_t h,
                                int32_t w, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Normalisation + activation + skip-concat (bandwidth-bound, vectorised, warp-shuffle reductions)
 * ---------------------------------------------------------------------------------------------------------- */

/* statistics granularity */
#define PETSYN_NORM_BATCH 0     /* nn.BatchNorm3d: per channel over N*D*H*W        (unet_model.py:50,52)        */
#define PETSYN_NORM_INSTANCE 1  /* nn.InstanceNorm3d: per (n, channel)             (bmgan_model.py via MONAI ADN) */
#define PETSYN_NORM_GROUP 2     /* nn.GroupNorm: per (n, group)                    (atten_unet_model.py:593-612) */

/* Per-channel partial sums of a raw conv output z (bf16 [rows, c] contiguous, rows = n*d*h*w):
 * sums[2*c] (fp32, caller-zeroed): sum z, sum z^2 per channel (BATCH) */
int32_t petsyn_bn_stats(const void* z, double* sums, int64_t rows, int32_t c, void* stream);
/* mean/rstd -> fused affine (scale = gamma*rstd, shift = beta - mean*scale) and the running-stat update
 * (momentum, unbiased variance) of nn.BatchNorm3d in training mode.  In eval mode (training == 0) scale/shift are
 * derived from the running statistics and `sums` is ignored.  save_mean/save_rstd are kept for backward. */
int32_t petsyn_bn_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                           float* running_var, float* scale, float* shift, float* save_mean, float* save_rstd,
                           int64_t rows, int32_t c, float eps, float momentum, int32_t training, void* stream);
/* dst1 = act1(z*scale + shift) [, dst2 = act2(same)] written into channel slices of wider NDHWC buffers -- the
 * copy-free form of `torch.cat([model(x), x], 1)` (unet_model.py:99) including the reference's in-place-activation
 * aliasing (the skip half is LeakyReLU(x), then ReLU'd by the parent).  scale/shift may be NULL (no norm). */
int32_t petsyn_norm_act_fwd(const void* z, const float* scale, const float* shift, void* dst1, int32_t dst1_cstride,
                            int32_t dst1_coff, int32_t act1, void* dst2, int32_t dst2_cstride, int32_t dst2_coff,
                            int32_t act2, float slope, int64_t rows, int32_t c, void* stream);
/* Backward, pass 1: with b = z*scale+shift and g = g1*act1'(b) [+ g2*act2'(b)], accumulate per channel
 * sums[c] += sum g, sums[c + C] += sum g*zhat (zhat = (z-mean)*rstd).  sums caller-zeroed. */
int32_t petsyn_norm_act_bwd_reduce(const void* z, const float* scale, const float* shift, const float* mean,
                                   const float* rstd, const void* g1, int32_t g1_cstride, int32_t g1_coff,
                                   int32_t act1, const void* g2, int32_t g2_cstride, int32_t g2_coff, int32_t act2,
                                   float slope, double* sums, int64_t rows, int32_t c, void* stream);
/* Backward, pass 2: dz = gamma*rstd*(g - sum_g/rows - zhat*sum_gz/rows) (BatchNorm training backward), or
 * dz = g when mean == NULL (no norm).  Also emits dgamma = sum g*zhat, dbeta = sum g when dgamma != NULL. */
int32_t petsyn_norm_act_bwd_apply(const void* z, const float* scale, const float* shift, const float* mean,
                                  const float* rstd, const float* gamma, const void* g1, int32_t g1_cstride,
                                  int32_t g1_coff, int32_t act1, const void* g2, int32_t g2_cstride, int32_t g2_coff,
                                  int32_t act2, float slope, const double* sums, void* dz, float* dgamma,
                                  float* dbeta, int64_t rows, int32_t c, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Generalised normalisation + activation (descriptor API): batch / instance / group statistics, residual add,
 * gradient fan-out.  Conv -> InstanceNorm3d -> LeakyReLU, ResidualUnit sums and dense concatenation of BMGAN
 * (bmgan_model.py:12-70 through MONAI's Convolution/ADN/ResidualUnit/ConvDenseBlock), PatchGAN BatchNorm3d.
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct petsyn_normact_desc {
  const void* z;             /* raw conv output, bf16 [nsamples*rows, c], contiguous */
  int64_t rows;              /* voxels per statistics group (per sample for instance/group norm, all for batch norm) */
  int32_t c;
  int32_t nsamples;          /* 1 for batch statistics */
  int32_t per_sample_stats;  /* 1: scale/shift/mean/rstd/sums are [nsamples][c]; 0: [c] */
  const float* scale;        /* fused affine from petsyn_norm_finalize; NULL = no normalisation */
  const float* shift;
  const float* mean;         /* saved statistics (backward only); NULL = no normalisation */
  const float* rstd;
  const float* gamma;        /* [c] or NULL */
  void* t1;                  /* fwd: destination 1 (bf16 channel slice); bwd: gradient w.r.t. destination 1 */
  int32_t t1_cstride, t1_coff, act1;
  void* t2;                  /* optional second destination / gradient source */
  int32_t t2_cstride, t2_coff, act2;
  float slope;               /* LeakyReLU slope */
  void* res;                 /* fwd: residual input, out = act(norm(z)) + res; bwd: gradient w.r.t. res (output) */
  int32_t res_cstride, res_coff;
  int32_t res_accumulate;    /* bwd: add into res instead of overwriting */
  double* sums;              /* bwd workspace [nsamples|1][2c], DOUBLE precision: the CTAs' fp32 partial sums are added with
                              * 64-bit atomics, which is exact -- so reproducible from run to run -- unless a partial cancels
                              * below 2^-28 of the total (csrc/det_reduce.cuh) */
  void* dz;                  /* bwd: gradient w.r.t. z, bf16 contiguous */
  float* dgamma;             /* optional outputs (batch statistics with affine) */
  float* dbeta;
  int32_t group_size;        /* channels sharing statistics (GroupNorm); 0/1 = per channel.  With group_size > 1 or a
                              * per-sample affine the `sums` workspace must hold 4*nsamples*c floats */
  int32_t dz_accumulate;     /* bwd: add into dz (the normalised tensor has other consumers, e.g. a ResnetBlock skip) */
  int32_t affine_accumulate; /* bwd: add into dgamma/dbeta (group/per-sample-affine path) */
  const float* slope_dev;    /* PETSYN_ACT_PRELU: device scalar holding the slope (MONAI ResidualUnit act="PRELU") */
  double* dslope;            /* bwd: d(loss)/d(slope) accumulated into this device scalar (double, caller-zeroed) */
  double* dz_colsum;         /* bwd, optional: [c] += column sums of the FINAL dz (after `extra` and dz_accumulate) over all rows
                              * and samples (caller-zeroed) -- the bias gradient of the convolution whose output gradient
                              * this is, as a by-product of the apply pass */
  double* t1_stats;          /* fwd, optional: statistics of what is written to destination 1, for the normalisation that
                              * consumes it: [nsamples][2][t1_stats_c] doubles (sum, sum of squares), this op's channels at
                              * offset t1_stats_coff; caller-zeroed; needs nsamples = per-sample launch */
  int32_t t1_stats_c, t1_stats_coff;
  double* t2_stats;          /* a second statistics target for the same values (another consuming normalisation); allowed
                              * without a second destination */
  int32_t t2_stats_c, t2_stats_coff;
  const double* fin_sums;    /* fwd, optional: fuse petsyn_norm_finalize (Instance / Group normalisation) into this launch.
                              * [nsamples][2][c] statistics sums (petsyn_norm_stats layout); scale / shift / mean / rstd then
                              * are OUTPUTS ([nsamples][c], written for the backward pass); needs per_sample_stats */
  const float* fin_gamma;    /* [c] affine weight or NULL */
  const float* fin_beta;     /* [c] affine bias or NULL */
  int32_t fin_group_size;    /* channels sharing their statistics (GroupNorm); <= 1: per channel */
  float fin_eps;
  int32_t separate_group_combine; /* bwd: 1 = run the GroupNorm / per-sample-affine constant combine as its own launch (the
                              * default folds it into the apply kernel's prologue) */
  int32_t sums_prezeroed;    /* bwd: the caller has already cleared the first nsamples * 2 * c floats of `sums` (one fill for
                              * all the normalisations of a step instead of a memset per call) */
  int32_t z_cstride, z_coff; /* z as a channel slice [z_coff, z_coff + c) of a buffer with pitch z_cstride; 0 = contiguous */
  int32_t dz_cstride, dz_coff; /* the same for dz */
  const void* extra;         /* bwd, optional: bf16 channel slice ADDED to dz -- the gradient that reaches the normalised tensor
                              * through an identity skip connection (ResnetBlock: out = f(x) + x, atten_unet_model.py:662), so
                              * the residual sum needs no backward pass of its own */
  int32_t extra_cstride, extra_coff;
  int32_t dz_colsum_coff, dz_colsum_c; /* dz_colsum covers channels [dz_colsum_coff, +dz_colsum_c) of dz only (the convolution
                              * wrote a channel slice of a wider concat buffer); dz_colsum_c == 0: all c channels */
  int32_t sums_precomputed;  /* bwd: `sums` already holds (sum g, sum g * zhat) per (sample, channel) -- filled by the fused
                              * epilogue of the data-gradient convolution that produced the incoming gradient
                              * (petsyn_conv_dgrad_epi) -- so the reduction pass is skipped and only the apply pass runs */
} petsyn_normact_desc;

/* sums[sample][0:c] += sum z, sums[sample][c:2c] += sum z^2 (double accumulators, caller-zeroed). */
int32_t petsyn_norm_stats(const void* z, double* sums, int64_t rows, int32_t c, int32_t nsamples, void* stream);
/* The same for z given as a channel slice [coff, coff + c) of a buffer with channel pitch cstride. */
int32_t petsyn_norm_stats_slice(const void* z, int32_t cstride, int32_t coff, double* sums, int64_t rows, int32_t c,
                                int32_t nsamples, void* stream);
/* Statistics -> fused affine scale/shift per (sample, channel).  group_size channels share statistics (GroupNorm,
 * atten_unet_model.py:593-612); group_size 1 with nsamples == N is InstanceNorm3d, with nsamples == 1 BatchNorm3d
 * (running statistics updated in training mode, used in eval mode). */
int32_t petsyn_norm_finalize(const double* sums, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, float* scale, float* shift, float* save_mean, float* save_rstd,
                             int64_t rows, int32_t c, int32_t nsamples, int32_t group_size, float eps, float momentum,
                             int32_t training, void* stream);
int32_t petsyn_normact_fwd(const petsyn_normact_desc* desc, void* stream);
/* Reduction pass (only when normalised) + apply pass: dz, optional d(res), dgamma, dbeta. */
int32_t petsyn_normact_bwd(const petsyn_normact_desc* desc, void* stream);
/* dst[:, dst_coff:+c] (+)= src[:, src_coff:+c] on bf16 channel slices. */
int32_t petsyn_add_slice(const void* src, int32_t src_cstride, int32_t src_coff, void* dst, int32_t dst_cstride,
                         int32_t dst_coff, int64_t rows, int32_t c, int32_t accumulate, void* stream);
/* out[c] = sum over rows of x[:, coff + c] (bias gradient); out double, overwritten. */
int32_t petsyn_colsum(const void* x, int32_t cstride, int32_t coff, double* out, int64_t rows, int32_t c, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Covariate-conditioned generator (AttenUNet, unet/utils/atten_unet_model.py): resampling inside the up/down
 * ResnetBlocks and the token-stream ops of the level-3 SpatialTransformer.  The linear layers around them are k=1
 * convolutions of the conv family.
 * ---------------------------------------------------------------------------------------------------------- */

/* up == 0: dst[o] (+)= scale * sum of the 2x2x2 src block (AvgPool3d(2,2) with scale 1/8, atten_unet_model.py:498,
 * 652-654; also the backward of nearest upsampling with scale 1).  up == 1: dst[o] (+)= scale * src[o/2] (nearest x2,
 * :535, 646-651; also the backward of average pooling with scale 1/8).  (od, oh, ow) are dst's dims. */
int32_t petsyn_resample2(const void* src, int32_t src_cstride, int32_t src_coff, void* dst, int32_t dst_cstride,
                         int32_t dst_coff, int32_t n, int32_t od, int32_t oh, int32_t ow, int32_t c, int32_t up,
                         float scale, int32_t accumulate, void* stream);
/* nn.LayerNorm(c) over the channel dim of a token stream [rows, c] (bf16), atten_unet_model.py:221-223. */
int32_t petsyn_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                             int64_t rows, int32_t c, float eps, void* stream);
int32_t petsyn_layernorm_bwd(const void* x, const void* dy, const float* gamma, const float* mean, const float* rstd,
                             void* dx, float* dgamma, float* dbeta, int64_t rows, int32_t c, int32_t accumulate_dx,
                             void* stream);
/* MONAI MLPBlock(act="GEGLU") gate: h [rows, 2f] = (x | gate) -> x * gelu(gate) (exact erf GELU), :211. */
int32_t petsyn_geglu_fwd(const void* h, void* out, int64_t rows, int32_t f, void* stream);
int32_t petsyn_geglu_bwd(const void* h, const void* dout, void* dh, int64_t rows, int32_t f, void* stream);
/* softmax(scale * Q K^T) V per (sample, head) without materialising the L x L scores (the reference's baddbmm ->
 * softmax -> bmm, :137-154).  qkv bf16 [n*l, 3*heads*head_dim] = (q | k | v); out bf16 [n*l, heads*head_dim];
 * lse fp32 [n, heads, l] saved for backward.  head_dim must be 32 (num_head_channels = 32 in training.json). */
int32_t petsyn_attention_fwd(const void* qkv, void* out, float* lse, int32_t n, int32_t l, int32_t heads,
                             int32_t head_dim, float scale, void* stream);
int32_t petsyn_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta,
                             void* dqkv, int32_t n, int32_t l, int32_t heads, int32_t head_dim, float scale,
                             void* stream);
/* Covariate injection: cross-attention over a length-1 context collapses to a per-sample bias
 * to_out(to_v(context)) broadcast over all tokens (:156-175).  ctx fp32 [n, cctx]; wv [c, cctx];
 * wo [c, c]; bo [c]; vbuf/bias fp32 [n, c] scratch (kept for backward); tokens bf16 [n*rows_per_sample, c] updated
 * in place. */
int32_t petsyn_covariate_bias_fwd(const float* ctx, const float* wv, const float* wo, const float* bo, float* vbuf,
                                  float* bias, void* tokens, int32_t n, int32_t cctx, int32_t c,
                                  int64_t rows_per_sample, void* stream);
/* dbias = per-sample column sums of dtokens; dwv, dwo, dbo = gradients of to_v.weight, to_out.weight, to_out.bias
 * (to_q / to_k receive exactly zero gradient). */
int32_t petsyn_covariate_bias_bwd(const float* ctx, const float* wo, const float* vbuf, const void* dtokens,
                                  float* dbias, float* dwv, float* dwo, float* dbo, int32_t n, int32_t cctx, int32_t c,
                                  int64_t rows_per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Losses and optimiser
 * ---------------------------------------------------------------------------------------------------------- */

/* nn.L1Loss() forward + backward seed in one pass (train_unet.py:106,149): loss[0] += mean|y - t| (caller-zeroed),
 * dy = grad_scale * sign(y - t) / numel. */
int32_t petsyn_l1_loss_fwd_bwd(const float* y, const float* t, float* loss, float* dy, int64_t numel, float grad_scale,
                               void* stream);
/* Single-scale SSIM with a separable 5-tap Gaussian window on fp32 [N,1,D,H,W] volumes: the published torchmetrics
 * definition behind the reference's evaluation (unet/scripts/output_predict.py:73,126: gaussian window, kernel_size 5,
 * sigma 0.5, data_range 1), averaged over the voxels whose window lies inside the volume, and the gradient of the
 * reconstruction loss 1 - mean(SSIM) (BASELINE north star: L1/SSIM loss).
 *   ssim_sum[2*i + 0] += sum of SSIM over the (d-4)*(h-4)*(w-4) valid voxels of sample i (caller-zeroed, 2*n floats)
 *   ssim_sum[2*i + 1] += sum of the contrast-structure term (2 s_xy + C2)/(s_x^2 + s_y^2 + C2) (MS-SSIM's per-scale factor)
 *   dx          (+)= grad_scale * d(1 - mean SSIM)/dx (added when `accumulate`, e.g. on top of the L1 gradient), or NULL
 *               for evaluation only
 *   workspace   petsyn_ssim_workspace_bytes() bytes (three derivative maps), needed when dx != NULL */
size_t petsyn_ssim_workspace_bytes(int32_t n, int32_t d, int32_t h, int32_t w);
int32_t petsyn_ssim_fwd_bwd(const float* x, const float* y, float* ssim_sum, float* dx, void* workspace, int32_t n,
                            int32_t d, int32_t h, int32_t w, float data_range, float sigma, float grad_scale, int32_t accumulate,
                            void* stream);
/* F.avg_pool3d(kernel_size=2) on fp32 [n,1,d,h,w]: the down-sampling between MS-SSIM scales. */
int32_t petsyn_avgpool2_f32(const float* src, float* dst, int32_t n, int32_t d, int32_t h, int32_t w, void* stream);
/* out[0] += sum |x - y|, out[1] += sum (x - y)^2 (caller-zeroed): MAE and MSE -> PSNR = 10 log10(range^2 / MSE)
 * (output_predict.py:121-133). */
int32_t petsyn_abs_sq_err(const float* x, const float* y, float* out, int64_t numel, void* stream);
/* LSGAN PatchAdversarialLoss(criterion="least_squares"): mean (x - target)^2 and its gradient. */
int32_t petsyn_mse_const_fwd_bwd(const float* x, float target, float* loss, float* dx, int64_t numel, float grad_scale,
                                 void* stream);
/* kl_divergence of train_bmgan.py:33-40 averaged over the batch (:176): loss[0] += mean_n( -0.5 * sum_j(1 + lv - mu^2 -
 * exp(lv)) ) (caller-zeroed); dmu / dlv = gradient * grad_scale.  mu, lv fp32 with row pitch `pitch` (>= dim). */
int32_t petsyn_kl_fwd_bwd(const float* mu, const float* logvar, float* loss, float* dmu, float* dlogvar, int32_t n,
                          int32_t dim, int32_t pitch, float grad_scale, void* stream);
/* torch.optim.Adam step (no amsgrad, weight_decay 0) over one flat fp32 parameter arena.  The 1-based step number is
 * `step`, or -- when step_dev != NULL -- read from device memory (so a captured CUDA graph of the step stays valid). */
int32_t petsyn_adam_step(float* p, const float* g, float* m, float* v, int64_t numel, float lr, float beta1,
                         float beta2, float eps, int32_t step, const int32_t* step_dev, void* stream);
/* sum of squares of a flat fp32 array, accumulated into out[0] (caller-zeroed); used for gradient norms. */
int32_t petsyn_sumsq(const float* g, float* out, int64_t numel, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Input data format
 * ---------------------------------------------------------------------------------------------------------- */

/* One raw volume as the reference's loader holds it after sitk.GetArrayFromImage: fp32 [d, h, w], contiguous, in device
 * memory (the H2D copy of the raw array is the caller's). */
typedef struct petsyn_volume_src {
  const float* data;
  int32_t d, h, w;
} petsyn_volume_src;

/* pair_PET_T1dataset._preprocess_img (unet/utils/dataset.py:70-105) for a batch of n <= 16 raw volumes of individual
 * extents: SpatialPad(crop) -> CenterSpatialCrop(crop) -> img / torch.max(img), written as the network input
 * dst fp32 [n, d, h, w] (= [n, 1, d, h, w]).  vmax [n] fp32 receives each volume's maximum over the window (zero padding
 * included when the volume is smaller than the crop).  `srcs` is a HOST array; the division is IEEE fp32, so the result is
 * bit-identical to the torch expression (an all-zero volume gives NaN there and here). */
int32_t petsyn_volume_prepare(const petsyn_volume_src* srcs, int32_t n, float* dst, int32_t d, int32_t h, int32_t w,
                              float* vmax, void* stream);
/* The window of the two MONAI transforms along one axis: output voxel o reads raw voxel o + offset (outside = 0). */
int32_t petsyn_volume_window_offset(int32_t raw, int32_t roi);

#ifdef __cplusplus
}
#endif
#endif /* PETSYN_H_ */
