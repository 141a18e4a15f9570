// Structural-similarity loss and evaluation metrics on fp32 single-channel volumes [N, 1, D, H, W].
//
// The reference evaluates synthesized PET with MAE, PSNR and (MS-)SSIM (unet/scripts/output_predict.py:73,121-133: torchmetrics
// with a Gaussian window, kernel_size = 5, sigma = 0.5, data_range = 1) and BASELINE's north star asks for an L1/SSIM
// reconstruction loss.  torchmetrics is not part of the reference checkout; what is implemented here is its published
// single-scale definition: a separable 5-tap Gaussian window w (sigma 0.5), local moments mu_x = w*x, s_xx = w*x^2,
// s_xy = w*xy, and
//     SSIM = ((2 mu_x mu_y + C1)(2 sigma_xy + C2)) / ((mu_x^2 + mu_y^2 + C1)(sigma_x^2 + sigma_y^2 + C2)),
// C1 = (0.01 L)^2, C2 = (0.03 L)^2, averaged over the voxels whose window lies inside the volume (torchmetrics reflect-pads
// and then crops exactly that border away).  Checked in tests/ against a float64 CPU restatement (F.conv3d + autograd);
// parity is "unpinned" against torchmetrics itself.
//
//  ssim_fwd_kernel : one thread per valid output voxel, 125 taps from a shared-memory tile (36 x 12 x 12 voxels of x and y);
//                    writes the three derivative maps dS/dmu_x, dS/ds_xx, dS/ds_xy and block-reduces sum(SSIM).
//  ssim_bwd_kernel : dx[u] = -scale * (sum_v w(v-u) A[v] + 2 x[u] sum_v w(v-u) B[v] + y[u] sum_v w(v-u) C[v]), the
//                    adjoint of the window applied to the maps (zero outside the valid region).
#include <algorithm>
#include <cmath>

#include "common.h"

namespace petsyn {
namespace ssim {

constexpr int kR = 2;                         // window radius (kernel_size 5)
constexpr int kTW = 32, kTH = 8, kTD = 8;     // output tile
constexpr int kSW = kTW + 2 * kR, kSH = kTH + 2 * kR, kSD = kTD + 2 * kR;
constexpr int kBTD = 4, kBSD = kBTD + 2 * kR;   // backward: three map tiles in (static) shared memory -> shallower tile

struct Win { float w[5]; };

__device__ __forceinline__ float block_sum256(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 8) r = red[threadIdx.x];
  if (threadIdx.x < 32)
    for (int o = 4; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;   // valid in thread 0
}

// grid: (tiles_w * tiles_h * tiles_d, N); 256 threads, each computing kTW*kTH*kTD / 256 = 8 output voxels
__global__ void __launch_bounds__(256) ssim_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                       float* __restrict__ maps, float* __restrict__ ssim_sum, int D,
                                                       int H, int W, Win g, float c1, float c2) {
  __shared__ float sx[kSD][kSH][kSW + 1], sy[kSD][kSH][kSW + 1];
  __shared__ float red[8];
  const int od = D - 2 * kR, oh = H - 2 * kR, ow = W - 2 * kR;     // valid output extent
  const int tiles_w = (ow + kTW - 1) / kTW, tiles_h = (oh + kTH - 1) / kTH;
  int t = blockIdx.x;
  const int tw = t % tiles_w; t /= tiles_w;
  const int th = t % tiles_h; t /= tiles_h;
  const int td = t;
  const int n = blockIdx.y;
  const int64_t vol = (int64_t)D * H * W;
  const float* xn = x + n * vol;
  const float* yn = y + n * vol;
  const int w0 = tw * kTW, h0 = th * kTH, d0 = td * kTD;          // output origin == input origin of the tile (valid conv)
  for (int i = threadIdx.x; i < kSD * kSH * kSW; i += 256) {
    const int lw = i % kSW, lh = (i / kSW) % kSH, ld = i / (kSW * kSH);
    const int gw = w0 + lw, gh = h0 + lh, gd = d0 + ld;
    const bool ok = gw < W && gh < H && gd < D;
    const int64_t off = ((int64_t)gd * H + gh) * W + gw;
    sx[ld][lh][lw] = ok ? __ldg(xn + off) : 0.f;
    sy[ld][lh][lw] = ok ? __ldg(yn + off) : 0.f;
  }
  __syncthreads();
  float acc = 0.f, acc_cs = 0.f;
  for (int o = threadIdx.x; o < kTW * kTH * kTD; o += 256) {
    const int lw = o % kTW, lh = (o / kTW) % kTH, ld = o / (kTW * kTH);
    const int gw = w0 + lw, gh = h0 + lh, gd = d0 + ld;           // output voxel (valid-region coordinates)
    if (gw >= ow || gh >= oh || gd >= od) continue;
    float mx = 0.f, my = 0.f, sxx = 0.f, syy = 0.f, sxy = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const float wab = g.w[a] * g.w[b];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float wt = wab * g.w[c];
          const float xv = sx[ld + a][lh + b][lw + c], yv = sy[ld + a][lh + b][lw + c];
          mx += wt * xv; my += wt * yv;
          sxx += wt * xv * xv; syy += wt * yv * yv; sxy += wt * xv * yv;
        }
      }
    const float vx = sxx - mx * mx, vy = syy - my * my, cxy = sxy - mx * my;
    const float a1 = 2.f * mx * my + c1, a2 = 2.f * cxy + c2;
    const float b1 = mx * mx + my * my + c1, b2 = vx + vy + c2;
    const float inv = 1.f / (b1 * b2);
    const float s = a1 * a2 * inv;
    acc += s;
    acc_cs += a2 / b2;                    // contrast-structure term (what MS-SSIM multiplies across scales)
    if (maps != nullptr) {
      const int64_t ovol = (int64_t)od * oh * ow;
      const int64_t oo = n * 3 * ovol + ((int64_t)gd * oh + gh) * ow + gw;
      maps[oo] = 2.f * my * (a2 - a1) * inv - 2.f * mx * s * (b2 - b1) * inv;   // dS/dmu_x (total)
      maps[oo + ovol] = -s / b2;                                                // dS/ds_xx
      maps[oo + 2 * ovol] = 2.f * a1 * inv;                                     // dS/ds_xy
    }
  }
  const float tot = block_sum256(acc, red);
  __syncthreads();
  const float tot_cs = block_sum256(acc_cs, red);
  if (threadIdx.x == 0) {
    atomicAdd(ssim_sum + 2 * n, tot);
    atomicAdd(ssim_sum + 2 * n + 1, tot_cs);
  }
}

// grid: (tiles over the FULL volume, N)
__global__ void __launch_bounds__(256) ssim_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                       const float* __restrict__ maps, float* __restrict__ dx, int D,
                                                       int H, int W, Win g, float scale, int accumulate) {
  __shared__ float sa[kBSD][kSH][kSW + 1], sb[kBSD][kSH][kSW + 1], sc[kBSD][kSH][kSW + 1];
  const int od = D - 2 * kR, oh = H - 2 * kR, ow = W - 2 * kR;
  const int tiles_w = (W + kTW - 1) / kTW, tiles_h = (H + kTH - 1) / kTH;
  int t = blockIdx.x;
  const int tw = t % tiles_w; t /= tiles_w;
  const int th = t % tiles_h; t /= tiles_h;
  const int td = t;
  const int n = blockIdx.y;
  const int64_t vol = (int64_t)D * H * W, ovol = (int64_t)od * oh * ow;
  const float* mp = maps + n * 3 * ovol;
  const int w0 = tw * kTW, h0 = th * kTH, d0 = td * kBTD;         // input voxel origin of the tile
  // input voxel u receives from output voxels v = u - 2R .. u (valid-region coordinates): tile of maps starts at u0 - 2R
  for (int i = threadIdx.x; i < kBSD * kSH * kSW; i += 256) {
    const int lw = i % kSW, lh = (i / kSW) % kSH, ld = i / (kSW * kSH);
    const int vw = w0 + lw - 2 * kR, vh = h0 + lh - 2 * kR, vd = d0 + ld - 2 * kR;
    const bool ok = vw >= 0 && vh >= 0 && vd >= 0 && vw < ow && vh < oh && vd < od;
    const int64_t off = ((int64_t)vd * oh + vh) * ow + vw;
    sa[ld][lh][lw] = ok ? __ldg(mp + off) : 0.f;
    sb[ld][lh][lw] = ok ? __ldg(mp + ovol + off) : 0.f;
    sc[ld][lh][lw] = ok ? __ldg(mp + 2 * ovol + off) : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < kTW * kTH * kBTD; o += 256) {
    const int lw = o % kTW, lh = (o / kTW) % kTH, ld = o / (kTW * kTH);
    const int gw = w0 + lw, gh = h0 + lh, gd = d0 + ld;
    if (gw >= W || gh >= H || gd >= D) continue;
    float fa = 0.f, fb = 0.f, fc = 0.f;
    // output voxel v = u - k (k = tap index 0..4 per axis) sits at smem index (l + 2R - k) = l + 4 - k
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
      for (int b = 0; b < 5; ++b) {
        const float wab = g.w[a] * g.w[b];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
          const float wt = wab * g.w[c];
          fa += wt * sa[ld + 4 - a][lh + 4 - b][lw + 4 - c];
          fb += wt * sb[ld + 4 - a][lh + 4 - b][lw + 4 - c];
          fc += wt * sc[ld + 4 - a][lh + 4 - b][lw + 4 - c];
        }
      }
    const int64_t off = n * vol + ((int64_t)gd * H + gh) * W + gw;
    const float gval = scale * (fa + 2.f * __ldg(x + off) * fb + __ldg(y + off) * fc);
    dx[off] = accumulate ? dx[off] + gval : gval;
  }
}

// dst[n, d/2, h/2, w/2] = mean of the 2x2x2 block (F.avg_pool3d(kernel 2) between MS-SSIM scales; odd tails are dropped)
__global__ void __launch_bounds__(256) avgpool2_kernel(const float* __restrict__ src, float* __restrict__ dst, int N, int D,
                                                       int H, int W) {
  const int od = D / 2, oh = H / 2, ow = W / 2;
  const int64_t total = (int64_t)N * od * oh * ow;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % ow), h = (int)((i / ow) % oh), d = (int)((i / ((int64_t)ow * oh)) % od);
    const int n = (int)(i / ((int64_t)ow * oh * od));
    const float* p = src + (((int64_t)n * D + 2 * d) * H + 2 * h) * W + 2 * w;
    float s = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) s += p[((int64_t)a * H + b) * W] + p[((int64_t)a * H + b) * W + 1];
    dst[i] = 0.125f * s;
  }
}

// out[0] += sum |x - y|, out[1] += sum (x - y)^2   (MAE / MSE -> PSNR on the host; output_predict.py:121-133)
__global__ void __launch_bounds__(256) abs_sq_err_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                         float* __restrict__ out, int64_t n) {
  __shared__ float red[8];
  float a = 0.f, b = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = x[i] - y[i];
    a += fabsf(d);
    b += d * d;
  }
  const float sa = block_sum256(a, red);
  __syncthreads();
  const float sb = block_sum256(b, red);
  if (threadIdx.x == 0) {
    atomicAdd(out, sa);
    atomicAdd(out + 1, sb);
  }
}

static Win gaussian(float sigma) {
  Win g;
  float s = 0.f;
  for (int i = 0; i < 5; ++i) {
    const float d = (float)(i - 2) / sigma;
    g.w[i] = std::exp(-0.5f * d * d);
    s += g.w[i];
  }
  for (int i = 0; i < 5; ++i) g.w[i] /= s;
  return g;
}

}  // namespace ssim
}  // namespace petsyn

using namespace petsyn;

extern "C" {

size_t petsyn_ssim_workspace_bytes(int32_t n, int32_t d, int32_t h, int32_t w) {
  if (n <= 0 || d < 5 || h < 5 || w < 5) return 0;
  return (size_t)n * 3 * (d - 4) * (h - 4) * (w - 4) * sizeof(float);
}

int32_t petsyn_ssim_fwd_bwd(const float* x, const float* y, float* ssim_sum, float* dx, void* workspace, int32_t n,
                            int32_t d, int32_t h, int32_t w, float data_range, float sigma, float grad_scale, int32_t accumulate,
                            void* stream) {
  PETSYN_REQUIRE(x && y && ssim_sum, "null argument");
  PETSYN_REQUIRE(n > 0 && d >= 5 && h >= 5 && w >= 5, "SSIM needs volumes of at least 5 voxels per axis");
  PETSYN_REQUIRE(dx == nullptr || workspace != nullptr, "the gradient needs the workspace (petsyn_ssim_workspace_bytes)");
  PETSYN_REQUIRE(data_range > 0.f && sigma > 0.f, "data_range and sigma must be positive");
  cudaStream_t st = as_stream(stream);
  const ssim::Win g = ssim::gaussian(sigma);
  const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
  const int od = d - 4, oh = h - 4, ow = w - 4;
  const double nvalid = (double)n * od * oh * ow;
  {
    dim3 grid((unsigned)(((ow + ssim::kTW - 1) / ssim::kTW) * ((oh + ssim::kTH - 1) / ssim::kTH) * ((od + ssim::kTD - 1) / ssim::kTD)),
              (unsigned)n);
    ssim::ssim_fwd_kernel<<<grid, 256, 0, st>>>(x, y, dx ? reinterpret_cast<float*>(workspace) : nullptr, ssim_sum, d, h, w,
                                                g, c1, c2);
    int32_t rc = check_launch("ssim_fwd_kernel");
    if (rc) return rc;
  }
  if (dx) {
    dim3 grid((unsigned)(((w + ssim::kTW - 1) / ssim::kTW) * ((h + ssim::kTH - 1) / ssim::kTH) * ((d + ssim::kBTD - 1) / ssim::kBTD)),
              (unsigned)n);
    // loss = 1 - mean(SSIM): d loss / dx = -(1 / nvalid) * adjoint terms
    ssim::ssim_bwd_kernel<<<grid, 256, 0, st>>>(x, y, reinterpret_cast<const float*>(workspace), dx, d, h, w, g,
                                                (float)(-grad_scale / nvalid), accumulate);
    int32_t rc = check_launch("ssim_bwd_kernel");
    if (rc) return rc;
  }
  return PETSYN_OK;
}

int32_t petsyn_avgpool2_f32(const float* src, float* dst, int32_t n, int32_t d, int32_t h, int32_t w, void* stream) {
  PETSYN_REQUIRE(src && dst && n > 0 && d >= 2 && h >= 2 && w >= 2, "bad argument");
  const int64_t total = (int64_t)n * (d / 2) * (h / 2) * (w / 2);
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, 148 * 8));
  ssim::avgpool2_kernel<<<blocks, 256, 0, as_stream(stream)>>>(src, dst, n, d, h, w);
  return check_launch("avgpool2_kernel");
}

int32_t petsyn_abs_sq_err(const float* x, const float* y, float* out, int64_t numel, void* stream) {
  PETSYN_REQUIRE(x && y && out && numel > 0, "bad argument");
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((numel + 255) / 256, 148 * 8));
  ssim::abs_sq_err_kernel<<<blocks, 256, 0, as_stream(stream)>>>(x, y, out, numel);
  return check_launch("abs_sq_err_kernel");
}

}  // extern "C"
