"""Thin Python objects over the C ABI: one class per kernel family, tensors in, tensors out.

Nothing here computes; every method forwards raw device pointers and the current CUDA stream to
``libpetsyn.so``.  Activations are channels-last ``(N, D, H, W, C)`` bf16 tensors (possibly channel slices
of wider buffers, described by ``cstride``/``coff``); weights are fp32 in PyTorch layout.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import os
from collections import OrderedDict

import torch

from . import _cabi
from ._cabi import (ACT_LRELU, ACT_NONE, ACT_PRELU, ACT_RELU, ACT_SILU, ACT_TANH, OP_CONV, OP_CONVT, OP_UPCONV, check, lib, ptr,
                    stream_ptr)

__all__ = ["ConvPlan", "OP_CONV", "OP_UPCONV", "OP_CONVT", "ACT_NONE", "ACT_RELU", "ACT_LRELU", "ACT_SILU",
           "ACT_TANH", "ACT_PRELU", "kl_fwd_bwd", "bn_stats", "bn_finalize", "norm_act_fwd", "norm_act_bwd", "l1_loss_fwd_bwd",
           "mse_const_fwd_bwd", "adam_step", "sumsq", "StemConv", "HeadConv"]


class ConvPlan:
    """Conv3d / Upsample+Conv3d / ConvTranspose3d as tcgen05 implicit GEMMs (fprop, dgrad, wgrad).

    Replaces ``F.conv3d`` (unet_model.py:47,60), ``nn.Upsample`` + ``F.conv3d`` (unet_model.py:59-60) and
    ``F.conv_transpose3d`` (bmgan_model.py:57-64) together with their cuDNN backward kernels.
    """

    def __init__(self, op: int, n: int, d: int, h: int, w: int, cin: int, cout: int, ksize: int, stride: int, pad: int,
                 x_cstride: Optional[int] = None, x_coff: int = 0, y_cstride: Optional[int] = None, y_coff: int = 0,
                 dy_cstride: Optional[int] = None, dy_coff: int = 0, dx_cstride: Optional[int] = None,
                 dx_coff: int = 0, act: int = ACT_NONE, slope: float = 0.2, y_fp32: bool = False):
        _cabi.require_cuda()
        self.desc = _cabi.ConvDesc(op, n, d, h, w, cin, cout, ksize, stride, pad,
                                   x_cstride or cin, x_coff, y_cstride or cout, y_coff,
                                   dy_cstride or cout, dy_coff, dx_cstride or cin, dx_coff, act, slope, int(y_fp32))
        handle = C.c_void_p()
        check(lib.petsyn_conv_plan_create(C.byref(self.desc), C.byref(handle)), "conv_plan_create")
        self._h = handle
        od, oh, ow = C.c_int32(), C.c_int32(), C.c_int32()
        check(lib.petsyn_conv_out_dims(self._h, C.byref(od), C.byref(oh), C.byref(ow)))
        self.out_dims: Tuple[int, int, int] = (od.value, oh.value, ow.value)
        fa, fe = C.c_double(), C.c_double()
        check(lib.petsyn_conv_flops(self._h, C.byref(fa), C.byref(fe)))
        self.flops_algorithmic, self.flops_executed = fa.value, fe.value
        # kernel family per pass (0 gather-form igemm, 1 slab, 2 small-channel wgrad): reporting only
        self.kernel_path = tuple(lib.petsyn_conv_kernel_path(self._h, i) for i in range(3))
        # fused epilogues (petsyn_conv_{fprop,dgrad}_epi): available per pass (fprop, dgrad)
        sup = [int(lib.petsyn_conv_epilogue_supported(self._h, i)) for i in range(2)]
        self.epi_ok = tuple(v == 1 for v in sup)          # any fused epilogue (residual, statistics, norm-backward reduce)
        self.epi_stats_ok = sup[0] in (1, 2)               # forward pass: at least the statistics epilogue
        self.packed_fprop_bytes = lib.petsyn_conv_packed_fprop_bytes(self._h)
        self.packed_dgrad_bytes = lib.petsyn_conv_packed_dgrad_bytes(self._h)
        self.wgrad_scratch_bytes = lib.petsyn_conv_wgrad_scratch_bytes(self._h)
        ws = lib.petsyn_conv_workspace_bytes(self._h)
        self.workspace: Optional[torch.Tensor] = None
        if ws:
            self.workspace = torch.empty(ws, dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
            check(lib.petsyn_conv_set_workspace(self._h, ptr(self.workspace), ws), "conv_set_workspace")
        self.w_fprop: Optional[torch.Tensor] = None
        self.w_dgrad: Optional[torch.Tensor] = None
        self._scratch: Optional[torch.Tensor] = None

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.petsyn_conv_plan_destroy(h)
            self._h = None

    # weights -------------------------------------------------------------------------------------------
    def pack(self, w: torch.Tensor, need_dgrad: bool = True) -> None:
        """fp32 PyTorch-layout weights -> packed bf16 GEMM operands held by the plan."""
        assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
        if self.w_fprop is None:
            self.w_fprop = torch.empty(self.packed_fprop_bytes, dtype=torch.uint8, device=w.device)
        if need_dgrad and self.w_dgrad is None:
            self.w_dgrad = torch.empty(self.packed_dgrad_bytes, dtype=torch.uint8, device=w.device)
        check(lib.petsyn_conv_pack_weights(self._h, ptr(w), ptr(self.w_fprop),
                                           ptr(self.w_dgrad) if need_dgrad else None, stream_ptr()), "conv_pack_weights")

    # kernels -------------------------------------------------------------------------------------------
    def fprop(self, x: torch.Tensor, y: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
        check(lib.petsyn_conv_fprop(self._h, ptr(x), ptr(self.w_fprop), ptr(bias), ptr(y), stream_ptr()), "conv_fprop")
        return y

    def dgrad(self, dy: torch.Tensor, dx: torch.Tensor) -> torch.Tensor:
        check(lib.petsyn_conv_dgrad(self._h, ptr(dy), ptr(self.w_dgrad), ptr(dx), stream_ptr()), "conv_dgrad")
        return dx

    def fprop_epi(self, x: torch.Tensor, y: torch.Tensor, bias: Optional[torch.Tensor], epi: "_cabi.ConvEpilogue") -> torch.Tensor:
        """fprop with a fused epilogue (residual add and / or the statistics of the consuming GroupNorm)."""
        check(lib.petsyn_conv_fprop_epi(self._h, ptr(x), ptr(self.w_fprop), ptr(bias), ptr(y), C.byref(epi), stream_ptr()),
              "conv_fprop_epi")
        return y

    def dgrad_epi(self, dy: torch.Tensor, dx: torch.Tensor, epi: "_cabi.ConvEpilogue") -> torch.Tensor:
        """dgrad whose epilogue also does the reduction pass of the backward of the normalisation in front of this conv."""
        check(lib.petsyn_conv_dgrad_epi(self._h, ptr(dy), ptr(self.w_dgrad), ptr(dx), C.byref(epi), stream_ptr()),
              "conv_dgrad_epi")
        return dx

    def wgrad(self, x: torch.Tensor, dy: torch.Tensor, dw: torch.Tensor, accumulate: bool = False,
              scratch: Optional[torch.Tensor] = None, dbias_acc: Optional[torch.Tensor] = None,
              dbias: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``dbias_acc`` (float64 column sums of dy, filled earlier on this stream) / ``dbias`` (fp32 gradient slot): the bias
        gradient is converted into its slot by the launch that lays out dw (no launch of its own)."""
        if scratch is None:
            if self._scratch is None:
                self._scratch = torch.empty(self.wgrad_scratch_bytes, dtype=torch.uint8, device=x.device)
            scratch = self._scratch
        assert scratch.numel() * scratch.element_size() >= self.wgrad_scratch_bytes
        if dbias is not None:
            assert dbias_acc.dtype == torch.float64 and dbias.dtype == torch.float32 and dbias.is_contiguous()
            check(lib.petsyn_conv_wgrad_bias(self._h, ptr(x), ptr(dy), ptr(scratch), ptr(dw), int(accumulate), ptr(dbias_acc),
                                             ptr(dbias), dbias.numel(), stream_ptr()), "conv_wgrad_bias")
        else:
            check(lib.petsyn_conv_wgrad(self._h, ptr(x), ptr(dy), ptr(scratch), ptr(dw), int(accumulate), stream_ptr()),
                  "conv_wgrad")
        return dw


# ------------------------------------------------------------------------------------------------ edge layers
class StemConv:
    """Conv3d(1 -> C, k4 s2 p1, no bias) on the fp32 NCDHW network input (unet_model.py:47,62).

    im2col of the single-channel volume (``petsyn_stem_im2col_k4s2``) turns it into a K = 64 GEMM that runs on the
    tcgen05 conv kernel (k = 1); the patches are kept for the weight gradient.
    """

    def __init__(self, n: int, d: int, h: int, w: int, cout: int, device):
        self.n, self.d, self.h, self.w, self.cout = n, d, h, w, cout
        self.rows = n * (d // 2) * (h // 2) * (w // 2)
        self.patches = torch.empty(self.rows, 64, dtype=torch.bfloat16, device=device)
        self.plan = ConvPlan(OP_CONV, n, d // 2, h // 2, w // 2, 64, cout, 1, 1, 0)

    def pack(self, w: torch.Tensor) -> None:
        self.plan.pack(w.view(self.cout, 64, 1, 1, 1), need_dgrad=False)

    def fprop(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        check(lib.petsyn_stem_im2col_k4s2(ptr(x), ptr(self.patches), self.n, self.d, self.h, self.w, stream_ptr()),
              "stem_im2col")
        return self.plan.fprop(self.patches, y)

    def wgrad(self, dy: torch.Tensor, dw: torch.Tensor) -> torch.Tensor:
        """dw [C,1,4,4,4] from the patches of the last fprop."""
        return self.plan.wgrad(self.patches, dy, dw)


class HeadConv:
    """ReLU'd skip tensor -> Upsample x2 -> Conv3d(C -> 1, k3 p1, no bias) -> Tanh (unet_model.py:59-64).

    fwd: proj[s, k] = <x[s, :], W[k, :]> (k = 1 conv, cout = 32, fp32 out, tensor cores) then a 27-term gather + tanh.
    bwd: scatter of dy*(1 - y^2) into dproj (bf16, 64 wide) then dx = dproj @ W (k = 1 conv) and dW = wgrad of the proj.
    """

    def __init__(self, n: int, d: int, h: int, w: int, cin: int, device):
        self.n, self.d, self.h, self.w, self.cin = n, d, h, w, cin
        self.rows = n * d * h * w
        self.proj = torch.empty(self.rows, 32, dtype=torch.float32, device=device)
        self.dproj = torch.zeros(self.rows, 64, dtype=torch.bfloat16, device=device)
        self.plan_proj = ConvPlan(OP_CONV, n, d, h, w, cin, 32, 1, 1, 0, dy_cstride=64, y_fp32=True)
        self.plan_dx = ConvPlan(OP_CONV, n, d, h, w, 64, cin, 1, 1, 0)
        self.w_proj = torch.zeros(32, cin, 1, 1, 1, dtype=torch.float32, device=device)
        self.w_dx = torch.zeros(cin, 64, 1, 1, 1, dtype=torch.float32, device=device)
        self.dw_proj = torch.empty(32, cin, 1, 1, 1, dtype=torch.float32, device=device)

    def pack(self, w: torch.Tensor, need_bwd: bool = True) -> None:
        """w: [1, C, 3, 3, 3] fp32."""
        wk = w.view(self.cin, 27)
        self.w_proj.view(32, self.cin)[:27].copy_(wk.t())
        self.plan_proj.pack(self.w_proj, need_dgrad=False)
        if need_bwd:
            self.w_dx.view(self.cin, 64)[:, :27].copy_(wk)
            self.plan_dx.pack(self.w_dx, need_dgrad=False)

    def fprop(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        self.plan_proj.fprop(x, self.proj)
        check(lib.petsyn_head_gather_tanh(ptr(self.proj), ptr(y), self.n, self.d, self.h, self.w, stream_ptr()),
              "head_gather_tanh")
        return y

    def backward(self, x: torch.Tensor, y: torch.Tensor, dy: torch.Tensor, dx: torch.Tensor, dw: torch.Tensor) -> None:
        check(lib.petsyn_head_scatter_bwd(ptr(y), ptr(dy), ptr(self.dproj), self.n, self.d, self.h, self.w,
                                          stream_ptr()), "head_scatter_bwd")
        self.plan_dx.fprop(self.dproj, dx)
        self.plan_proj.wgrad(x, self.dproj, self.dw_proj)
        dw.view(self.cin, 27).copy_(self.dw_proj.view(32, self.cin)[:27].t())


# ------------------------------------------------------------------------------------------------ norm / act
def bn_stats(z: torch.Tensor, sums: torch.Tensor, rows: int, c: int) -> None:
    check(lib.petsyn_bn_stats(ptr(z), ptr(sums), rows, c, stream_ptr()), "bn_stats")


def bn_finalize(sums, gamma, beta, running_mean, running_var, scale, shift, save_mean, save_rstd, rows: int, c: int,
                eps: float, momentum: float, training: bool) -> None:
    check(lib.petsyn_bn_finalize(ptr(sums), ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), ptr(scale),
                                 ptr(shift), ptr(save_mean), ptr(save_rstd), rows, c, eps, momentum, int(training),
                                 stream_ptr()), "bn_finalize")


def norm_act_fwd(z, scale, shift, dst1, cs1: int, co1: int, act1: int, dst2, cs2: int, co2: int, act2: int,
                 slope: float, rows: int, c: int) -> None:
    check(lib.petsyn_norm_act_fwd(ptr(z), ptr(scale), ptr(shift), ptr(dst1), cs1, co1, act1, ptr(dst2), cs2, co2, act2,
                                  slope, rows, c, stream_ptr()), "norm_act_fwd")


def norm_act_bwd(z, scale, shift, mean, rstd, gamma, g1, cs1: int, co1: int, act1: int, g2, cs2: int, co2: int,
                 act2: int, slope: float, sums, dz, dgamma, dbeta, rows: int, c: int) -> None:
    """Backward of norm_act_fwd: reduction pass (only when normalised) then the apply pass."""
    if mean is not None:
        sums.zero_()
        check(lib.petsyn_norm_act_bwd_reduce(ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(g1), cs1, co1,
                                             act1, ptr(g2), cs2, co2, act2, slope, ptr(sums), rows, c, stream_ptr()),
              "norm_act_bwd_reduce")
    check(lib.petsyn_norm_act_bwd_apply(ptr(z), ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(gamma), ptr(g1), cs1,
                                        co1, act1, ptr(g2), cs2, co2, act2, slope, ptr(sums), ptr(dz), ptr(dgamma),
                                        ptr(dbeta), rows, c, stream_ptr()), "norm_act_bwd_apply")


# ------------------------------------------------------------------------------------------------ losses / optimiser
def l1_loss_fwd_bwd(y: torch.Tensor, t: torch.Tensor, loss: torch.Tensor, dy: Optional[torch.Tensor],
                    grad_scale: float = 1.0) -> None:
    """nn.L1Loss() value (accumulated into the zeroed scalar ``loss``) and its gradient (train_unet.py:106,149)."""
    check(lib.petsyn_l1_loss_fwd_bwd(ptr(y), ptr(t), ptr(loss), ptr(dy), y.numel(), grad_scale, stream_ptr()),
          "l1_loss")


class SsimLoss:
    """1 - mean(SSIM) on fp32 [N,1,D,H,W] volumes (Gaussian window, kernel 5, sigma 0.5, data_range 1 -- the parameters of
    ``unet/scripts/output_predict.py:73``), value and gradient w.r.t. the prediction.  Holds the derivative-map workspace."""

    def __init__(self, shape, device, data_range: float = 1.0, sigma: float = 0.5):
        n, c, d, h, w = shape
        if c != 1:
            raise ValueError("SSIM kernels take single-channel volumes")
        self.n, self.d, self.h, self.w = n, d, h, w
        self.data_range, self.sigma = data_range, sigma
        self.count = n * (d - 4) * (h - 4) * (w - 4)
        self.per_sample = (d - 4) * (h - 4) * (w - 4)
        self.workspace = torch.empty(lib.petsyn_ssim_workspace_bytes(n, d, h, w), dtype=torch.uint8, device=device)
        self.sum = torch.zeros(n, 2, dtype=torch.float32, device=device)     # per sample: sum SSIM, sum contrast-structure

    def __call__(self, x: torch.Tensor, y: torch.Tensor, dx: Optional[torch.Tensor] = None, grad_scale: float = 1.0,
                 accumulate: bool = False) -> torch.Tensor:
        """Returns the device scalar mean SSIM; ``dx`` (if given) receives grad_scale * d(1 - mean SSIM)/dx (added to its
        contents when ``accumulate``)."""
        self.sum.zero_()
        check(lib.petsyn_ssim_fwd_bwd(ptr(x), ptr(y), ptr(self.sum), ptr(dx), ptr(self.workspace) if dx is not None else None,
                                      self.n, self.d, self.h, self.w, self.data_range, self.sigma, grad_scale, int(accumulate),
                                      stream_ptr()), "ssim")
        return self.sum[:, 0].sum() / self.count


MS_SSIM_BETAS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0, sigma: float = 0.5) -> torch.Tensor:
    """Multi-scale SSIM as the reference evaluates it (``MultiScaleStructuralSimilarityIndexMeasure(kernel_size=5,
    sigma=0.5, data_range=1)``, unet/scripts/output_predict.py:73,126): five scales, 2x average pooling in between, per
    image prod_i relu(cs_i)^beta_i * relu(ssim_last)^beta_last, averaged over the batch.  Returns a device scalar."""
    n, c, d, h, w = x.shape
    if min(d, h, w) // 2 ** (len(MS_SSIM_BETAS) - 1) < 5:
        raise ValueError("MS-SSIM with a 5-tap window needs at least 80 voxels per axis (5 at the coarsest of 5 scales)")
    x, y = x.contiguous().float(), y.contiguous().float()
    terms = []
    for i, beta in enumerate(MS_SSIM_BETAS):
        crit = SsimLoss(tuple(x.shape), x.device, data_range, sigma)
        crit(x, y)
        last = i == len(MS_SSIM_BETAS) - 1
        val = crit.sum[:, 0 if last else 1] / crit.per_sample
        terms.append(torch.relu(val) ** beta)
        if not last:
            nd, nh, nw = x.shape[2] // 2, x.shape[3] // 2, x.shape[4] // 2
            px, py = (torch.empty(n, 1, nd, nh, nw, dtype=torch.float32, device=x.device) for _ in range(2))
            for src, dst in ((x, px), (y, py)):
                check(lib.petsyn_avgpool2_f32(ptr(src), ptr(dst), n, x.shape[2], x.shape[3], x.shape[4], stream_ptr()),
                      "avgpool2")
            x, y = px, py
    return torch.stack(terms).prod(0).mean()


def eval_metrics(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0):
    """MAE and PSNR of a synthesized volume (output_predict.py:121-133); returns device scalars."""
    out = torch.zeros(2, dtype=torch.float32, device=x.device)
    check(lib.petsyn_abs_sq_err(ptr(x), ptr(y), ptr(out), x.numel(), stream_ptr()), "abs_sq_err")
    mae = out[0] / x.numel()
    psnr = 10.0 * torch.log10(data_range ** 2 / (out[1] / x.numel()))
    return mae, psnr


def mse_const_fwd_bwd(x, target: float, loss, dx, grad_scale: float = 1.0) -> None:
    check(lib.petsyn_mse_const_fwd_bwd(ptr(x), target, ptr(loss), ptr(dx), x.numel(), grad_scale, stream_ptr()),
          "mse_const")


def kl_fwd_bwd(mu, logvar, loss, dmu, dlogvar, n: int, dim: int, pitch: int, grad_scale: float = 1.0) -> None:
    """kl_divergence(...).mean() of train_bmgan.py:33-40,174-176 and its gradient."""
    check(lib.petsyn_kl_fwd_bwd(ptr(mu), ptr(logvar), ptr(loss), ptr(dmu), ptr(dlogvar), n, dim, pitch, grad_scale,
                                stream_ptr()), "kl_fwd_bwd")


def adam_step(p, g, m, v, lr: float, beta1: float, beta2: float, eps: float, step: int,
              step_dev: Optional[torch.Tensor] = None) -> None:
    """torch.optim.Adam update; ``step_dev`` (int32 device scalar) overrides ``step`` for graph-captured loops."""
    check(lib.petsyn_adam_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, step, ptr(step_dev),
                               stream_ptr()), "adam_step")


def launch_count() -> int:
    """Kernels launched by libpetsyn so far in this process."""
    return int(lib.petsyn_launch_count())


def sumsq(g: torch.Tensor, out: torch.Tensor) -> None:
    check(lib.petsyn_sumsq(ptr(g), ptr(out), g.numel(), stream_ptr()), "sumsq")


class EngineCache(OrderedDict):
    """Per-module cache of engines keyed by (input shape, device): least recently used entries are dropped beyond
    ``PETSYN_MAX_ENGINES`` (default 4) so that inference over many volume shapes does not keep every shape's activation
    buffers alive.  An evicted engine still referenced by a pending autograd node or a trainer lives until they let go."""

    def get(self, key, default=None):
        if key in self:
            self.move_to_end(key)
            return self[key]
        return default

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        self.move_to_end(key)
        limit = max(1, int(os.environ.get("PETSYN_MAX_ENGINES", "4")))
        while len(self) > limit:
            self.popitem(last=False)
