"""The two decoders of the causal_synthesis model -- LABELLED RESTATEMENTS, PARITY UNPINNED.

``causal_synthesis/scripts/train_unify_causal_gen.py:5-7`` imports ``DiffusionModelEncoder``, ``Decoder`` and
``DiffusionModelDecoder`` from the authors' un-vendored ``monai_diffusion`` fork; none of their source is in the reference
checkout, so nothing here can be checked against the authors' arithmetic.  What IS fixed by the reference:

  * the constructor keywords (``causal_synthesis/configs/training_causal.json:40-74``: ``decoder`` = in 3, out 1, channels
    [32, 64, 64, 64], 2 ResBlocks per level, 32 groups, no attention; ``pet_decoder_def`` = in 3, out 1, channels [64, 64, 32],
    2 ResnetBlocks per level, 32 groups, attention levels [True, False, False], ``with_conditioning``, plus
    ``cross_attention_dim = len(need_values)`` injected at :114-116);
  * the call sites (:213-224): ``latent = t1_encoder(t1)``; ``z_mu, z_sigma = latent[:, :3], latent[:, 3:]``;
    ``t1_rec = t1_decoder(z_mu + eps * z_sigma)``; ``rec_pet = pet_decoder(z_mu + eps' * z_sigma, info)`` -- a 3-channel latent
    in, one full-resolution volume out, i.e. a x8 up-sampling for the 96x128x96 crop and its 12x16x12 latent;
  * the losses (:57-73, 226-247): L1 on both reconstructions and ``kl_divergence(z_mu, z_sigma)`` (sigma passed where the
    formula expects a log-variance, SURVEY 9 Q8 -- reproduced as written by ``kl_divergence`` below).

What is RESTATED (choices, not facts):
  * ``Decoder`` follows upstream MONAI-GenerativeModels ``autoencoderkl.Decoder`` (module list ``blocks``: conv, per level
    [ResBlock x2 (+ Upsample between levels)], GroupNorm, conv -- no activation before the last conv, as upstream).
  * ``DiffusionModelDecoder`` is not an upstream class; it is built from the VENDORED blocks of
    ``unet/utils/atten_unet_model.py`` (ResnetBlock :565-662, SpatialTransformer :238-343 with the covariate cross-attention,
    Upsample :510-562) the way ``AttenUNet``'s up path is, without skip connections: conv_in, per level [ResnetBlock
    (+ SpatialTransformer) x num_res_blocks, Upsample(use_conv=True)] (EVERY level up-samples: three levels <-> x8), then
    GroupNorm -> SiLU -> conv.  Attention heads have the vendored default of 8 channels.
The encoder stays out: ``training_causal.json``'s ``atten_encoder`` read with the vendored block semantics would put
self-attention on a 48x64x48 grid (147 456 tokens), which nobody trained at batch 2 -- the fork's class evidently differs.

Both modules run on libpetsyn's kernels through the same op tape as ``AttenUNet`` and are differentiable w.r.t. their
parameters AND the latent input, so they compose with any encoder through autograd.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from ._cabi import check, lib, ptr, stream_ptr
from .atten_unet_model import (ResnetBlock, SpatialTransformer, Upsample, _AttenEngine, _Container, _Convolution, _rep,
                               _ZeroGrad)
from .bmgan_model import _EngineBase
from .graph import Buf, NormActOp, Sl


def kl_divergence(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """``train_unify_causal_gen.py:57-73`` as written: ``-0.5 * sum(1 + logvar - mu^2 - exp(logvar)) / N``.  The script calls it
    as ``kl_divergence(z_mu, z_sigma)`` (:228): sigma where a log-variance is expected."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / mu.shape[0]


def reparameterize(z_mu: torch.Tensor, z_sigma: torch.Tensor, eps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``z_mu + eps * z_sigma`` with ``eps = randn_like(z_sigma)`` (:217-224)."""
    return z_mu + (torch.randn_like(z_sigma) if eps is None else eps) * z_sigma


class _ResBlock(_Container):
    """upstream ``autoencoderkl.ResBlock``: norm1 -> SiLU -> conv1 -> norm2 -> SiLU -> conv2, + nin_shortcut(x)."""

    def __init__(self, cin: int, cout: int, groups: int, eps: float):
        super().__init__()
        self.channels, self.out_channels = cin, cout
        self.up = self.down = False
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps, affine=True)
        self.conv1 = _Convolution(cin, cout, 3)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps, affine=True)
        self.conv2 = _Convolution(cout, cout, 3)
        self.nin_shortcut = nn.Identity() if cin == cout else _Convolution(cin, cout, 1)

    @property
    def skip_connection(self):            # the name the engine's ResnetBlock builder reads
        return self.nin_shortcut


class _LatentFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx_, eng, x, context, *params):
        ctx_.need_dx = x.requires_grad
        y = eng.forward(x, context).clone()
        eng.stamp(ctx_, (x,) if context is None else (x, context))
        ctx_.has_context = context is not None
        return y

    @staticmethod
    def backward(ctx_, dy):
        ctx_.eng.restore(ctx_)
        dx, grads = ctx_.eng.backward(dy.contiguous().float(), need_dx=ctx_.need_dx)
        return (None, dx, None, *grads)


class _DecoderBase(nn.Module):
    _engines: Dict[Tuple, "_DecoderEngine"]

    def _check(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError(f"petsyn {type(self).__name__} runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected a latent of shape [N, {self.in_channels}, D, H, W], got {tuple(x.shape)}")
        return x.contiguous().float()

    def engine_for(self, x: torch.Tensor) -> "_DecoderEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _DecoderEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def _run(self, x: torch.Tensor, context: Optional[torch.Tensor]) -> torch.Tensor:
        eng = self.engine_for(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in eng.params)):
            return _LatentFn.apply(eng, x, context, *eng.params)
        return eng.forward(x, context).clone()


class Decoder(_DecoderBase):
    """T1 decoder: ``Decoder(**training_causal.json['t1_autoencoder_def']['decoder'])`` (train_unify_causal_gen.py:110,218).
    RESTATEMENT of upstream ``generative.networks.nets.autoencoderkl.Decoder`` -- parity unpinned (module docstring)."""

    def __init__(self, spatial_dims: int, num_channels: Sequence[int], in_channels: int, out_channels: int,
                 num_res_blocks: Sequence[int] | int, norm_num_groups: int, norm_eps: float, attention_levels: Sequence[bool],
                 with_nonlocal_attn: bool = True, with_encoder_nonlocal_attn: bool = False,
                 with_decoder_nonlocal_attn: Optional[bool] = None, use_flash_attention: bool = False,
                 use_convtranspose: bool = False) -> None:
        super().__init__()
        n = len(num_channels)
        num_res_blocks = _rep(num_res_blocks, n)
        nonlocal_attn = with_nonlocal_attn if with_decoder_nonlocal_attn is None else with_decoder_nonlocal_attn
        if spatial_dims != 3 or out_channels != 1 or any(attention_levels) or nonlocal_attn or use_convtranspose:
            raise NotImplementedError("petsyn Decoder implements training_causal.json's T1 decoder: 3-D, one output channel, no "
                                      "attention levels, no non-local attention, nearest + conv up-sampling")
        if any(c % norm_num_groups for c in num_channels):
            raise ValueError("Decoder expects all num_channels being multiple of norm_num_groups")
        self.in_channels, self.out_channels = in_channels, out_channels
        rch, rres = list(reversed(list(num_channels))), list(reversed(num_res_blocks))
        blocks: List[nn.Module] = [_Convolution(in_channels, rch[0], 3)]
        cout = rch[0]
        for i in range(n):
            cin, cout = cout, rch[i]
            for _ in range(rres[i]):
                blocks.append(_ResBlock(cin, cout, norm_num_groups, norm_eps))
                cin = cout
            if i != n - 1:
                blocks.append(Upsample(cin))
        blocks.append(nn.GroupNorm(norm_num_groups, cin, eps=norm_eps, affine=True))
        blocks.append(_Convolution(cin, out_channels, 3))
        self.blocks = nn.ModuleList(blocks)
        self._engines = ops.EngineCache()

    def plan(self):
        """(conv_in, [stage, ...], final norm, SiLU before the last conv?, out conv); a stage is ('res', block, None) or
        ('up', Upsample)."""
        b = list(self.blocks)
        stages = [("res", m, None) if isinstance(m, _ResBlock) else ("up", m) for m in b[1:-2]]
        return b[0].conv, stages, b[-2], False, b[-1].conv

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self._run(self._check(x), None)


class _UpStage(_Container):
    def __init__(self, prev: int, c: int, nres: int, groups: int, eps: float, attn: Optional[dict]):
        super().__init__()
        if attn is not None:
            self.attentions = nn.ModuleList([SpatialTransformer(c, **attn) for _ in range(nres)])
        self.resnets = nn.ModuleList([ResnetBlock(prev if j == 0 else c, c, norm_num_groups=groups, norm_eps=eps)
                                      for j in range(nres)])
        self.upsampler = Upsample(c)


class DiffusionModelDecoder(_DecoderBase):
    """PET decoder with covariate cross-attention: ``DiffusionModelDecoder(**training_causal.json['pet_decoder_def'],
    cross_attention_dim=len(need_values))`` (train_unify_causal_gen.py:114-116,224).  NOT an upstream class: a RESTATEMENT built
    from the vendored AttenUNet blocks -- parity unpinned (module docstring)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, num_channels: Sequence[int] = (64, 64, 32),
                 num_res_blocks: Sequence[int] | int = 2, norm_num_groups: int = 32, norm_eps: float = 1e-6,
                 attention_levels: Sequence[bool] = (True, False, False), with_conditioning: bool = False,
                 cross_attention_dim: Optional[int] = None, num_head_channels: int | Sequence[int] = 8,
                 transformer_num_layers: int = 1, upcast_attention: bool = False, use_flash_attention: bool = False) -> None:
        super().__init__()
        if with_conditioning is True and cross_attention_dim is None:
            raise ValueError("DiffusionModelDecoder expects dimension of the cross-attention conditioning (cross_attention_dim) "
                             "when using with_conditioning.")
        if cross_attention_dim is not None and with_conditioning is False:
            raise ValueError("DiffusionModelDecoder expects with_conditioning=True when specifying the cross_attention_dim.")
        if any((c % norm_num_groups) != 0 for c in num_channels):
            raise ValueError("DiffusionModelDecoder expects all num_channels being multiple of norm_num_groups")
        if len(num_channels) != len(attention_levels):
            raise ValueError("DiffusionModelDecoder expects num_channels being same size of attention_levels")
        n = len(num_channels)
        num_head_channels, num_res_blocks = _rep(num_head_channels, n), _rep(num_res_blocks, n)
        if spatial_dims != 3 or out_channels != 1 or not with_conditioning or transformer_num_layers != 1:
            raise NotImplementedError("petsyn DiffusionModelDecoder implements training_causal.json's use: 3-D, one output "
                                      "channel, with_conditioning=True, one transformer layer")
        for lvl, a in enumerate(attention_levels):
            if a and (num_head_channels[lvl] not in (8, 16, 32) or num_channels[lvl] % num_head_channels[lvl]):
                raise NotImplementedError("attention levels need num_head_channels in {8, 16, 32} dividing the channel count")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.cfg = dict(cross_attention_dim=cross_attention_dim)
        ch, g, e = list(num_channels), norm_num_groups, norm_eps
        self.conv_in = _Convolution(in_channels, ch[0], 3)
        self.up_blocks = nn.ModuleList([])
        prev = ch[0]
        for i in range(n):
            attn = None
            if attention_levels[i]:
                attn = dict(heads=ch[i] // num_head_channels[i], head_channels=num_head_channels[i], num_layers=1,
                            norm_num_groups=g, norm_eps=e, cross_attention_dim=cross_attention_dim)
            self.up_blocks.append(_UpStage(prev, ch[i], num_res_blocks[i], g, e, attn))
            prev = ch[i]
        self.out = nn.Sequential(nn.GroupNorm(g, ch[-1], eps=e, affine=True), nn.SiLU(), _Convolution(ch[-1], out_channels, 3))
        self._engines = ops.EngineCache()

    def plan(self):
        stages = []
        for blk in self.up_blocks:
            for j, rb in enumerate(blk.resnets):
                stages.append(("res", rb, blk.attentions[j] if hasattr(blk, "attentions") else None))
            stages.append(("up", blk.upsampler))
        return self.conv_in.conv, stages, self.out[0], True, self.out[2].conv

    def forward(self, x: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = self._check(x)
        if context is None:
            raise ValueError("DiffusionModelDecoder(with_conditioning=True) needs the covariate context tensor")
        ctx = context.reshape(x.shape[0], -1).contiguous().float()
        if ctx.shape[1] != self.cfg["cross_attention_dim"]:
            raise ValueError(f"context has {ctx.shape[1]} covariates, expected {self.cfg['cross_attention_dim']}")
        return self._run(x, ctx)


class _DecoderEngine(_AttenEngine):
    """Op tape of a decoder for one latent shape: the ResnetBlock / SpatialTransformer builders of the AttenUNet engine in a
    straight line (no skip connections), conv-form up-sampling through the phase-decomposed gather kernel."""

    def __init__(self, net: _DecoderBase, shape, dev):
        _EngineBase.__init__(self, net, dev)
        n, cin, D, H, W = shape
        self.shape, self.n = shape, n
        self.context: Optional[torch.Tensor] = None
        self.zero = _ZeroGrad()
        self._zero_params: List[nn.Parameter] = []
        t = self.tape
        conv_in, stages, norm_out, silu_out, conv_out = net.plan()
        self.inp = Buf(n, D, H, W, self.CPAD, dev, "latent")
        self.cin = cin
        h: Sl = self._conv(self.inp.sl(), conv_in, ksize=3, stride=1, pad=1, need_dx=True, name="conv_in").z.sl()
        for k, st in enumerate(stages):
            if st[0] == "res":
                h = self._resnet(st[1], h, 0, None, f"s{k}.res")
                if st[2] is not None:
                    h = self._transformer(st[2], h, 0, None, f"s{k}.attn")
            else:                               # Upsample(use_conv=True): nearest x2 + Conv3d k3, never materialised
                h = self._conv(h, st[1].conv.conv, ksize=3, stride=1, pad=1, op=ops.OP_UPCONV, name=f"s{k}.up").z.sl()
        a = Buf(n, h.buf.d, h.buf.h, h.buf.w, h.c, dev, "out.a")
        gn = NormActOp(h, "group", ops.ACT_SILU if silu_out else ops.ACT_NONE, [a.sl()], gn=norm_out)
        t.add(gn)
        self._bind += [(gn, "grad_gamma", norm_out.weight), (gn, "grad_beta", norm_out.bias)]
        self.head = self._conv(a.sl(), conv_out, ksize=3, stride=1, pad=1, y_fp32=True, name="out.conv")
        t.add(self.zero)
        self.out_dims = (h.buf.d, h.buf.h, h.buf.w)
        self.y = torch.zeros(n, 1, *self.out_dims, dtype=torch.float32, device=dev)
        self._finish()
        for p in self._zero_params:
            if all(p is not q for q in self.params):
                self.params.append(p)

    def forward(self, x: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        n, cin, D, H, W = self.shape
        self.generation += 1
        self.context = context
        # layout conversion of the (12x16x12-sized) latent: NCDHW fp32 -> channels-last bf16, zero-padded to CPAD channels
        self.inp.t.view(n, D, H, W, self.CPAD)[..., :cin].copy_(x.permute(0, 2, 3, 4, 1))
        self.tape.forward(self.training())
        check(lib.petsyn_take_channel0(ptr(self.head.zf), ptr(self.y), self.y.numel(), self.head.cout, stream_ptr()),
              "take_channel0")
        return self.y

    def backward(self, dy: torch.Tensor, need_dx: bool = False, out: Optional[Dict[int, torch.Tensor]] = None, on_ready=None):
        n, cin, D, H, W = self.shape
        grads = self.grad_slots(out)
        check(lib.petsyn_put_channel0_grad(None, ptr(dy), ptr(self.head.zg), dy.numel(), self.head.cout, 0,
                                           stream_ptr()), "put_channel0_grad")
        self.run_backward(on_ready)
        dx = None
        if need_dx:
            dx = self.inp.g.view(n, D, H, W, self.CPAD)[..., :cin].permute(0, 4, 1, 2, 3).float().contiguous()
        return dx, ([g.clone() for g in grads] if out is None else [])
