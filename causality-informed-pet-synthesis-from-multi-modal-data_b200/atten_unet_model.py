"""Drop-in replacement for the reference's covariate-conditioned generator ``AttenUNet``
(``unet/utils/atten_unet_model.py:1575-1860``) running on libpetsyn's sm_100a kernels.

Kept identical to the reference: the constructor signature and its ``ValueError`` checks, the module tree (``conv_in``,
``down_blocks.{i}.resnets/attentions/downsampler``, ``middle_block.resnet_1/attention/resnet_2``,
``up_blocks.{i}.resnets/attentions/upsampler``, ``out``; MONAI's ``Convolution(conv_only=True)`` child name ``conv``,
``MLPBlock``'s ``linear1/linear2``) and therefore every ``state_dict`` key and shape (416 tensors for
``unet/config/training.json``), the ``zero_module`` initialisation, ``forward(x, context)`` on fp32 NCDHW tensors, autograd.

Implemented configuration families: ``resblock_updown`` True (``training.json``) or False (conv-form resampling, the
reference's own smoke block); ``with_conditioning=True`` (attention levels use cross-attention transformer blocks with one
or more transformer layers; heads of 8, 16 or 32 channels) or False (``AttentionBlock`` with 32-channel heads, no context); any
number of levels / ResnetBlocks per level.  Class embeddings (broken in the reference itself) and
attention dropout raise ``NotImplementedError``.

Execution: a static op tape (``graph.py``).  GroupNorm+SiLU, residual sums and skip concatenation are fused
bandwidth kernels over channels-last bf16 buffers; every Conv3d (3^3, 1^3, the nearest-x2 + 3^3 of the up-sampling
ResnetBlocks) and every Linear of the transformer runs on the tcgen05 implicit-GEMM kernel; self-attention is a
flash-style kernel (no L x L score tensor); cross-attention over the length-1 covariate context is computed as what it
is mathematically -- a per-sample bias ``to_out(to_v(context))`` added to every token -- with exact zero
gradients for ``attn2.to_q``, ``attn2.to_k`` and ``norm2``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import os

import torch
import torch.nn as nn

from . import graph, ops
from ._cabi import check, lib, ptr, stream_ptr
from .bmgan_model import _EngineBase
from .graph import (AttentionOp, Buf, ConvOp, CovariateBiasOp, DropoutOp, GegluOp, LayerNormOp, NormActOp, ResampleOp, Sl, Tape)


def zero_module(module: nn.Module) -> nn.Module:
    for p in module.parameters():
        p.detach().zero_()
    return module


class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("petsyn AttenUNet blocks are parameter containers; call AttenUNet.forward")


class _Convolution(nn.Sequential):
    """MONAI ``Convolution(conv_only=True)``: one child named ``conv``."""

    def __init__(self, cin: int, cout: int, k: int, stride: int = 1):
        super().__init__()
        self.add_module("conv", nn.Conv3d(cin, cout, k, stride=stride, padding=(k - 1) // 2, bias=True))


class Downsample(_Container):
    """``Downsample(use_conv=True)`` (atten_unet_model.py:464-507): ``op`` = Conv3d(k3, stride 2, padding 1)."""

    def __init__(self, channels: int):
        super().__init__()
        self.num_channels = self.out_channels = channels
        self.op = _Convolution(channels, channels, 3, stride=2)


class Upsample(_Container):
    """``Upsample(use_conv=True)`` (atten_unet_model.py:510-562): nearest x2, then ``conv`` = Conv3d(k3, padding 1)."""

    def __init__(self, channels: int, out_channels: Optional[int] = None):
        super().__init__()
        self.num_channels, self.out_channels = channels, out_channels or channels
        self.conv = _Convolution(channels, self.out_channels, 3)


class ResnetBlock(_Container):
    def __init__(self, in_channels: int, out_channels: Optional[int] = None, up: bool = False, down: bool = False,
                 norm_num_groups: int = 32, norm_eps: float = 1e-6):
        super().__init__()
        self.channels = in_channels
        self.out_channels = out_channels or in_channels
        self.up, self.down = up, down
        self.norm1 = nn.GroupNorm(norm_num_groups, in_channels, eps=norm_eps, affine=True)
        self.nonlinearity = nn.SiLU()
        self.conv1 = _Convolution(in_channels, self.out_channels, 3)
        self.upsample = self.downsample = None
        if up:
            self.upsample = nn.Upsample(scale_factor=2.0, mode="nearest")
        elif down:
            self.downsample = nn.AvgPool3d(kernel_size=2, stride=2)
        self.norm2 = nn.GroupNorm(norm_num_groups, self.out_channels, eps=norm_eps, affine=True)
        self.conv2 = zero_module(_Convolution(self.out_channels, self.out_channels, 3))
        self.skip_connection = (nn.Identity() if self.out_channels == in_channels
                                else _Convolution(in_channels, self.out_channels, 1))


class CrossAttention(_Container):
    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int, head_channels: int):
        super().__init__()
        inner = heads * head_channels
        cross_attention_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.num_heads, self.head_channels = heads, head_channels
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))


class _MLPBlock(_Container):
    """MONAI ``MLPBlock(act="GEGLU")``."""

    def __init__(self, hidden: int, mlp_dim: int):
        super().__init__()
        self.linear1 = nn.Linear(hidden, mlp_dim * 2)
        self.linear2 = nn.Linear(mlp_dim, hidden)
        self.drop1, self.drop2 = nn.Dropout(0.0), nn.Dropout(0.0)


class BasicTransformerBlock(_Container):
    def __init__(self, channels: int, heads: int, head_channels: int, cross_attention_dim: Optional[int]):
        super().__init__()
        self.attn1 = CrossAttention(channels, None, heads, head_channels)
        self.ff = _MLPBlock(channels, channels * 4)
        self.attn2 = CrossAttention(channels, cross_attention_dim, heads, head_channels)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(channels), nn.LayerNorm(channels), nn.LayerNorm(channels)


class SpatialTransformer(_Container):
    def __init__(self, in_channels: int, heads: int, head_channels: int, num_layers: int, norm_num_groups: int,
                 norm_eps: float, cross_attention_dim: Optional[int]):
        super().__init__()
        inner = heads * head_channels
        self.norm = nn.GroupNorm(norm_num_groups, in_channels, eps=norm_eps, affine=True)
        self.proj_in = _Convolution(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, head_channels, cross_attention_dim) for _ in range(num_layers)])
        self.proj_out = zero_module(_Convolution(inner, in_channels, 1))


class AttentionBlock(_Container):
    """``AttentionBlock`` (atten_unet_model.py:346-461), the attention of the ``with_conditioning=False`` family: GroupNorm,
    ``to_q`` / ``to_k`` / ``to_v`` with bias, multi-head self-attention, residual.  ``proj_attn`` is constructed (:383) but the
    reference's forward never applies it: it keeps its place in the state dict and receives zero gradients."""

    def __init__(self, channels: int, head_channels: Optional[int], norm_num_groups: int, norm_eps: float):
        super().__init__()
        self.num_channels = channels
        self.num_heads = channels // head_channels if head_channels is not None else 1
        self.norm = nn.GroupNorm(norm_num_groups, channels, eps=norm_eps, affine=True)
        self.to_q, self.to_k, self.to_v = nn.Linear(channels, channels), nn.Linear(channels, channels), nn.Linear(channels, channels)
        self.proj_attn = nn.Linear(channels, channels)


def _attention_module(c: int, attn: dict) -> nn.Module:
    if attn.get("block"):                                 # with_conditioning=False
        return AttentionBlock(c, attn["head_channels"], attn["norm_num_groups"], attn["norm_eps"])
    return SpatialTransformer(c, **attn)


class _DownBlock(_Container):
    def __init__(self, cin, cout, nres, groups, eps, add_downsample, attn: Optional[dict], resblock_updown: bool = True):
        super().__init__()
        # construction order = the reference's (a resnet, then its attention: :790-810, :893-926), so a seeded default
        # initialisation draws the same numbers; registration order: `attentions` before `resnets` (:812-813, :929-930)
        resnets, attentions = [], []
        for i in range(nres):
            resnets.append(ResnetBlock(cin if i == 0 else cout, cout, norm_num_groups=groups, norm_eps=eps))
            if attn is not None:
                attentions.append(_attention_module(cout, attn))
        if attn is not None:
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        self.downsampler = None
        if add_downsample:        # atten_unet_model.py:712-730: a down-sampling ResnetBlock, or Downsample(use_conv=True)
            self.downsampler = (ResnetBlock(cout, cout, down=True, norm_num_groups=groups, norm_eps=eps)
                                if resblock_updown else Downsample(cout))


class _MidBlock(_Container):
    def __init__(self, c, groups, eps, attn: dict):
        super().__init__()
        self.resnet_1 = ResnetBlock(c, c, norm_num_groups=groups, norm_eps=eps)
        self.attention = _attention_module(c, attn)
        self.resnet_2 = ResnetBlock(c, c, norm_num_groups=groups, norm_eps=eps)


class _UpBlock(_Container):
    def __init__(self, cin, prev, cout, nres, groups, eps, add_upsample, attn: Optional[dict], resblock_updown: bool = True):
        super().__init__()
        resnets, attentions = [], []
        for i in range(nres):
            skip_c = cin if i == nres - 1 else cout
            in_c = prev if i == 0 else cout
            resnets.append(ResnetBlock(in_c + skip_c, cout, norm_num_groups=groups, norm_eps=eps))
            if attn is not None:
                attentions.append(_attention_module(cout, attn))
        if attn is not None and not attn.get("block"):    # CrossAttnUpBlock: attentions first (:1371-1372)
            self.attentions = nn.ModuleList(attentions)
        self.resnets = nn.ModuleList(resnets)
        if attn is not None and attn.get("block"):        # AttnUpBlock: resnets first (:1253-1254)
            self.attentions = nn.ModuleList(attentions)
        self.upsampler = None
        if add_upsample:          # atten_unet_model.py:1148-1165
            self.upsampler = (ResnetBlock(cout, cout, up=True, norm_num_groups=groups, norm_eps=eps)
                              if resblock_updown else Upsample(cout))


def _rep(v, n):
    return list(v) if isinstance(v, (list, tuple)) else [v] * n


class AttenUNet(nn.Module):
    """B200-native ``AttenUNet`` (reference atten_unet_model.py:1575-1860): same ctor, keys and forward contract."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int,
                 num_res_blocks: Sequence[int] | int = (2, 2, 2, 2), num_channels: Sequence[int] = (32, 64, 64, 64),
                 attention_levels: Sequence[bool] = (False, False, True, True), norm_num_groups: int = 32,
                 norm_eps: float = 1e-6, resblock_updown: bool = False, num_head_channels: int | Sequence[int] = 8,
                 with_conditioning: bool = False, transformer_num_layers: int = 1,
                 cross_attention_dim: int | None = None, num_class_embeds: int | None = None,
                 upcast_attention: bool = False, use_flash_attention: bool = False,
                 dropout_cattn: float = 0.0) -> None:
        super().__init__()
        # the reference's own argument checks (atten_unet_model.py:1623-1666)
        if with_conditioning is True and cross_attention_dim is None:
            raise ValueError("AttenUNet expects dimension of the cross-attention conditioning (cross_attention_dim) "
                             "when using with_conditioning.")
        if cross_attention_dim is not None and with_conditioning is False:
            raise ValueError("AttenUNet expects with_conditioning=True when specifying the cross_attention_dim.")
        if dropout_cattn > 1.0 or dropout_cattn < 0.0:
            raise ValueError("Dropout cannot be negative or >1.0!")
        if any((c % norm_num_groups) != 0 for c in num_channels):
            raise ValueError("AttenUNet expects all num_channels being multiple of norm_num_groups")
        if len(num_channels) != len(attention_levels):
            raise ValueError("AttenUNet expects num_channels being same size of attention_levels")
        n = len(num_channels)
        num_head_channels = _rep(num_head_channels, n)
        if len(num_head_channels) != n:
            raise ValueError("num_head_channels should have the same length as attention_levels.")
        num_res_blocks = _rep(num_res_blocks, n)
        if len(num_res_blocks) != n:
            raise ValueError("`num_res_blocks` should be a single integer or a tuple of integers with the same length "
                             "as `num_channels`.")
        if spatial_dims != 3 or in_channels != 1 or out_channels != 1:
            raise NotImplementedError("petsyn AttenUNet implements the reference use: 3-D, one channel in, one out")
        if transformer_num_layers < 1 or num_class_embeds is not None or dropout_cattn != 0.0:
            raise NotImplementedError("petsyn AttenUNet implements no class embeddings / attention dropout")
        for lvl, a in enumerate(list(attention_levels) + [True]):          # the middle block always attends
            hc = num_head_channels[min(lvl, n - 1)]
            if a and hc not in (8, 16, 32):
                raise NotImplementedError("attention needs num_head_channels in {8, 16, 32} (heads narrower than the attention "
                                          "kernel's 32 channels are zero-padded)")
            if a and num_channels[min(lvl, n - 1)] % hc:
                raise ValueError("num_channels must be divisible by num_head_channels at attention levels")
        self.cfg = dict(num_channels=list(num_channels), num_res_blocks=num_res_blocks,
                        attention_levels=list(attention_levels), num_head_channels=num_head_channels,
                        norm_num_groups=norm_num_groups, norm_eps=norm_eps, cross_attention_dim=cross_attention_dim,
                        resblock_updown=bool(resblock_updown))
        self.in_channels, self.out_channels = in_channels, out_channels
        self.block_out_channels = num_channels
        self.with_conditioning = with_conditioning
        ch, g, e = list(num_channels), norm_num_groups, norm_eps

        def attn_kw(lvl):
            if not attention_levels[lvl]:
                return None
            if not with_conditioning:                     # Attn{Down,Mid,Up}Block: AttentionBlock, no context
                return dict(block=True, head_channels=num_head_channels[lvl], norm_num_groups=g, norm_eps=e)
            return dict(heads=ch[lvl] // num_head_channels[lvl], head_channels=num_head_channels[lvl],
                        num_layers=transformer_num_layers, norm_num_groups=g, norm_eps=e,
                        cross_attention_dim=cross_attention_dim)

        self.conv_in = _Convolution(in_channels, ch[0], 3)
        self.down_blocks = nn.ModuleList([])
        out_c = ch[0]
        for i in range(n):
            in_c, out_c = out_c, ch[i]
            self.down_blocks.append(_DownBlock(in_c, out_c, num_res_blocks[i], g, e, i != n - 1, attn_kw(i), resblock_updown))
        if with_conditioning:
            mid_attn = dict(heads=ch[-1] // num_head_channels[-1], head_channels=num_head_channels[-1],
                            num_layers=transformer_num_layers, norm_num_groups=g, norm_eps=e,
                            cross_attention_dim=cross_attention_dim)
        else:
            mid_attn = dict(block=True, head_channels=num_head_channels[-1], norm_num_groups=g, norm_eps=e)
        self.middle_block = _MidBlock(ch[-1], g, e, mid_attn)
        self.up_blocks = nn.ModuleList([])
        rch = list(reversed(ch))
        out_c = rch[0]
        for i in range(n):
            prev, out_c = out_c, rch[i]
            in_c = rch[min(i + 1, n - 1)]
            lvl = n - 1 - i
            self.up_blocks.append(_UpBlock(in_c, prev, out_c, num_res_blocks[lvl] + 1, g, e, i != n - 1, attn_kw(lvl),
                                           resblock_updown))
        self.out = nn.Sequential(nn.GroupNorm(g, ch[0], eps=e, affine=True), nn.SiLU(),
                                 zero_module(_Convolution(ch[0], out_channels, 3)))
        self._engines: Dict[Tuple, "_AttenEngine"] = ops.EngineCache()

    def engine_for(self, x: torch.Tensor) -> "_AttenEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _AttenEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def forward(self, x: torch.Tensor, context: torch.Tensor | None = None, class_labels: torch.Tensor | None = None,
                down_block_additional_residuals=None, mid_block_additional_residual=None) -> torch.Tensor:
        if class_labels is not None or down_block_additional_residuals is not None \
                or mid_block_additional_residual is not None:
            raise NotImplementedError("class labels / additional residuals are not used by the reference scripts")
        if context is not None and self.with_conditioning is False:       # atten_unet_model.py:1822-1823
            raise ValueError("model should have with_conditioning = True if context is provided")
        if context is None and self.with_conditioning:
            raise ValueError("AttenUNet(with_conditioning=True) needs the covariate context tensor")
        if not x.is_cuda:
            raise RuntimeError("petsyn AttenUNet runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        if self.with_conditioning:
            ctx = context.reshape(x.shape[0], -1).contiguous().float()    # [N, 1, C] or [N, C] (:110-112) -> [N, C]
            if ctx.shape[1] != self.cfg["cross_attention_dim"]:
                raise ValueError(f"context has {ctx.shape[1]} covariates, expected {self.cfg['cross_attention_dim']} "
                                 "(context length must be 1)")
        else:
            ctx = x.new_zeros(x.shape[0], 1)                              # placeholder: no op of the tape reads it
        eng = self.engine_for(x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in eng.params):
            return _AttenFn.apply(eng, x, ctx, *eng.params)
        return eng.forward(x, ctx).clone()


class _AttenFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx_, eng, x, context, *params):
        y = eng.forward(x, context).clone()
        eng.stamp(ctx_, (x, context))
        return y

    @staticmethod
    def backward(ctx_, dy):
        ctx_.eng.restore(ctx_)
        grads = ctx_.eng.backward(dy.contiguous().float())
        return (None, None, None, *grads)


class _ZeroGrad(graph.Op):
    """Parameters the forward never uses mathematically (attn2.to_q / to_k / norm2 with a length-1 context): zero grads."""

    def __init__(self):
        self.slots: List[torch.Tensor] = []
        self._zeroed: tuple = ()

    def bwd(self) -> None:
        # nothing on the path ever writes these slots, so they are cleared once per binding (24 fill launches a step otherwise)
        key = tuple(s.data_ptr() for s in self.slots)
        if key != self._zeroed:
            for s in self.slots:
                s.zero_()
            self._zeroed = key


_EPI_PROBE: Dict[Tuple, bool] = {}


def _residual_epilogue_ok(n: int, d: int, h: int, w: int, c: int) -> bool:
    """Can a Conv3d(c -> c, k3 s1 p1) on this grid add its residual in the kernel epilogue (petsyn_conv_fprop_epi)?"""
    if os.environ.get("PETSYN_NO_EPI_RES"):
        return False
    key = (n, d, h, w, c)
    if key not in _EPI_PROBE:
        _EPI_PROBE[key] = ops.ConvPlan(ops.OP_CONV, n, d, h, w, c, c, 3, 1, 1).epi_ok[0]
    return _EPI_PROBE[key]


class _AttenEngine(_EngineBase):
    CPAD = 16

    def __init__(self, net: AttenUNet, shape, dev):
        super().__init__(net, dev)
        n, _, D, H, W = shape
        cfg = net.cfg
        ch = cfg["num_channels"]
        nl = len(ch)
        if D % (1 << (nl - 1)) or H % (1 << (nl - 1)) or W % (1 << (nl - 1)):
            raise ValueError(f"spatial dims {D}x{H}x{W} must be divisible by {1 << (nl - 1)}")
        for c in ch:
            if c % 8:
                raise ValueError(f"channel widths must be multiples of 8 (got {c})")
        self.shape, self.n = shape, n
        self.context: Optional[torch.Tensor] = None
        self.zero = _ZeroGrad()
        self._zero_params: List[nn.Parameter] = []
        t = self.tape
        dims = lambda lvl: (D >> lvl, H >> lvl, W >> lvl)
        B = lambda lvl, c, name: Buf(n, *dims(lvl), c, dev, name)
        nres = cfg["num_res_blocks"]
        # ---- plan the skip list and the up path's concat buffers (cat([h, skip]) is never executed) ----
        skip_c, skip_lvl = [ch[0]], [0]
        for i in range(nl):
            for _ in range(nres[i]):
                skip_c.append(ch[i]); skip_lvl.append(i)
            if i != nl - 1:
                skip_c.append(ch[i]); skip_lvl.append(i + 1)
        cats: List[List[Buf]] = []
        slot: Dict[int, Sl] = {}
        k = len(skip_c) - 1
        h_c = ch[-1]
        for i in range(nl):
            lvl = nl - 1 - i
            row = []
            for j in range(nres[lvl] + 1):
                assert skip_lvl[k] == lvl
                cb = B(lvl, h_c + skip_c[k], f"up{i}.{j}.cat")
                slot[k] = cb.sl(h_c, skip_c[k])
                row.append(cb)
                h_c = ch[lvl]
                k -= 1
            cats.append(row)
        assert k == -1
        # ---- stem: conv_in writes straight into its skip slot; every later consumer reads that channel slice ----
        self.inp = B(0, self.CPAD, "input")
        self._conv(self.inp.sl(), net.conv_in.conv, ksize=3, stride=1, pad=1, need_dx=False, name="conv_in", out=slot[0])
        h: Sl = slot[0]
        sk = 1
        # ---- down path: a block's output lives ONLY in its skip slot (a channel slice of the up path's concat buffer) ----
        for i, blk in enumerate(net.down_blocks):
            for j, rb in enumerate(blk.resnets):
                if cfg["attention_levels"][i]:
                    h = self._resnet(rb, h, i, None, f"down{i}.{j}")
                    h = self._transformer(blk.attentions[j], h, i, slot[sk], f"down{i}.{j}.attn")
                else:
                    h = self._resnet(rb, h, i, slot[sk], f"down{i}.{j}")
                sk += 1
            if blk.downsampler is not None:
                if isinstance(blk.downsampler, ResnetBlock):
                    h = self._resnet(blk.downsampler, h, i, slot[sk], f"down{i}.ds", down=True)
                else:                                   # Downsample(use_conv=True): Conv3d k3 s2 p1 straight into the slot
                    self._conv(h, blk.downsampler.op.conv, ksize=3, stride=2, pad=1, name=f"down{i}.ds", out=slot[sk])
                    h = slot[sk]
                sk += 1
        # ---- middle ----
        mb = net.middle_block
        h = self._resnet(mb.resnet_1, h, nl - 1, None, "mid.r1")
        h = self._transformer(mb.attention, h, nl - 1, None, "mid.attn")
        self._resnet(mb.resnet_2, h, nl - 1, cats[0][0].sl(0, ch[-1]), "mid.r2")
        # ---- up path ----
        for i, blk in enumerate(net.up_blocks):
            lvl = nl - 1 - i
            for j, rb in enumerate(blk.resnets):
                last_in_block = j == len(blk.resnets) - 1
                nxt: Optional[Sl] = None if last_in_block else cats[i][j + 1].sl(0, ch[lvl])
                if cfg["attention_levels"][lvl]:
                    h = self._resnet(rb, cats[i][j].sl(), lvl, None, f"up{i}.{j}")
                    h = self._transformer(blk.attentions[j], h, lvl, nxt, f"up{i}.{j}.attn")
                else:
                    h = self._resnet(rb, cats[i][j].sl(), lvl, nxt, f"up{i}.{j}")
            if blk.upsampler is not None:
                nxt_h = cats[i + 1][0].sl(0, ch[lvl])
                if isinstance(blk.upsampler, ResnetBlock):
                    h = self._resnet(blk.upsampler, h, lvl, nxt_h, f"up{i}.us", up=True)
                else:                                   # Upsample(use_conv=True): nearest x2 + Conv3d k3 (never materialised)
                    self._conv(h, blk.upsampler.conv.conv, ksize=3, stride=1, pad=1, op=ops.OP_UPCONV, name=f"up{i}.us",
                               out=nxt_h)
                    h = nxt_h
        # ---- head ----
        a = B(0, ch[0], "out.a")
        gn = NormActOp(h, "group", ops.ACT_SILU, [a.sl()], gn=net.out[0])
        t.add(gn)
        self._bind += [(gn, "grad_gamma", net.out[0].weight), (gn, "grad_beta", net.out[0].bias)]
        self.head = self._conv(a.sl(), net.out[2].conv, ksize=3, stride=1, pad=1, y_fp32=True, name="out.conv")
        t.add(self.zero)
        self.y = torch.zeros(n, 1, D, H, W, dtype=torch.float32, device=dev)
        self._finish()
        for p in self._zero_params:
            if all(p is not q for q in self.params):
                self.params.append(p)

    # ------------------------------------------------------------------------------------------------ builders
    def _gn_act(self, z, gn: nn.GroupNorm, act: int, dst: Buf, extra: Optional[Sl] = None) -> NormActOp:
        op = NormActOp(z, "group", act, [dst.sl()], gn=gn, extra=extra)
        self.tape.add(op)
        self._bind += [(op, "grad_gamma", gn.weight), (op, "grad_beta", gn.bias)]
        return op

    def _resnet(self, rb: ResnetBlock, x: Sl, lvl: int, dst: Optional[Sl], name: str, up: bool = False,
                down: bool = False) -> Sl:
        """ResnetBlock.forward (atten_unet_model.py:641-662).  ``x``: the input as a channel slice; ``dst``: where the block
        output goes (a slot of a concat buffer) or None for a fresh buffer.  Returns the slice holding the output.

        The residual sum ``out = conv2(...) + skip(x)`` has NO backward pass: d out / d conv2 = d out / d skip = identity, so
        conv2 (and a 1x1 skip convolution, or the resampling of an up/down block) read their output gradient straight from
        out's gradient slice, and with an identity skip that slice is added to norm1's dz inside its apply pass (``extra``)."""
        t, dev, n = self.tape, self.dev, self.n
        cin, cout = rb.channels, rb.out_channels
        assert x.c == cin, (name, x.c, cin)
        xb = x.buf
        d, h, w = xb.d, xb.h, xb.w
        od, oh, ow = (2 * d, 2 * h, 2 * w) if up else ((d // 2, h // 2, w // 2) if down else (d, h, w))
        out = dst if dst is not None else Buf(n, od, oh, ow, cout, dev, name + ".out").sl()
        assert (out.buf.d, out.buf.h, out.buf.w, out.c) == (od, oh, ow, cout), name
        identity = isinstance(rb.skip_connection, nn.Identity)
        fused = _residual_epilogue_ok(n, od, oh, ow, cout)
        if fused and not identity:
            # the 1x1 skip convolution writes its result into the output slot, where conv2's epilogue picks it up and overwrites
            # it with the sum.  It is the FIRST op of the block: in backward its data gradient then runs last and adds (TMA
            # reduction) into the input gradient that norm1's backward has just written, instead of norm1's apply pass
            # read-modify-writing it (measured: 14.66 -> 14.38 ms per step; running it on a side stream beside
            # norm1 -> conv1 -> norm2 in forward was slower, 14.62 ms)
            sk = self._conv(x, rb.skip_connection.conv, ksize=1, stride=1, pad=0, name=name + ".skip", out=out)
            sk.absorbed = True
        a1 = Buf(n, d, h, w, cin, dev, name + ".a1")
        self._gn_act(x, rb.norm1, ops.ACT_SILU, a1, extra=out if (identity and not up and not down) else None)
        if down:
            assert identity
            a1p = Buf(n, od, oh, ow, cin, dev, name + ".a1p")
            xs = Buf(n, od, oh, ow, cin, dev, name + ".xs").sl()
            t.add(ResampleOp(a1.sl(), a1p.sl(), up=False))
            t.add(ResampleOp(x, xs, up=False, dst_grad=out))
            c1 = self._conv(a1p.sl(), rb.conv1.conv, ksize=3, stride=1, pad=1, name=name + ".conv1")
        elif up:
            assert identity
            xs = Buf(n, od, oh, ow, cin, dev, name + ".xs").sl()
            t.add(ResampleOp(x, xs, up=True, dst_grad=out))
            if cin <= 32 and cout <= 32:
                # few channels (the full-resolution end of the up path): materialising the up-sampled tensor (a bandwidth-bound copy) and running the
                # slab kernels on it beats the phase-decomposed gather-form kernel, whose 64-channel K chunks are half empty
                a1u = Buf(n, od, oh, ow, cin, dev, name + ".a1u")
                t.add(ResampleOp(a1.sl(), a1u.sl(), up=True))
                c1 = self._conv(a1u.sl(), rb.conv1.conv, ksize=3, stride=1, pad=1, name=name + ".conv1")
            else:
                c1 = self._conv(a1.sl(), rb.conv1.conv, ksize=3, stride=1, pad=1, op=ops.OP_UPCONV, name=name + ".conv1")
        else:
            xs = x
            c1 = self._conv(a1.sl(), rb.conv1.conv, ksize=3, stride=1, pad=1, name=name + ".conv1")
        a2 = Buf(n, od, oh, ow, cout, dev, name + ".a2")
        self._gn_act(c1.z, rb.norm2, ops.ACT_SILU, a2)
        if fused:
            # the residual sum happens in conv2's epilogue (the tile is added to the skip tile before it is stored, and summed
            # for the GroupNorm that reads the block output): no pass of its own
            self._conv(a2.sl(), rb.conv2.conv, ksize=3, stride=1, pad=1, name=name + ".conv2", out=out,
                       res=xs if identity else out)
            return out
        c2 = self._conv(a2.sl(), rb.conv2.conv, ksize=3, stride=1, pad=1, name=name + ".conv2", dy_from=out)
        if identity:
            res = xs
        else:
            res = self._conv(xs, rb.skip_connection.conv, ksize=1, stride=1, pad=0, name=name + ".skip", dy_from=out).z.sl()
        t.add(NormActOp(c2.z, "none", ops.ACT_NONE, [out], res=res, no_bwd=True))
        return out

    def _linear(self, x: Sl, lin: nn.Linear, name: str, out: Optional[Sl] = None, **kw) -> ConvOp:
        return self._conv(x, lin, ksize=1, stride=1, pad=0, name=name, out=out, **kw)

    def _attention_block(self, ab: AttentionBlock, x: Sl, dst: Optional[Sl], name: str) -> Sl:
        """AttentionBlock.forward (atten_unet_model.py:421-461): out = attention(to_q/k/v(GroupNorm(x))) + x."""
        t, dev, n = self.tape, self.dev, self.n
        c, xb = x.c, x.buf
        L = xb.d * xb.h * xb.w
        heads = ab.num_heads
        hd = c // heads
        if hd != 32:
            raise NotImplementedError("AttentionBlock heads must be 32 channels wide, the attention kernel's head size (there is "
                                      "no output projection that could absorb zero-padded heads; num_head_channels=None means "
                                      f"one head of {c} channels)")
        T = lambda ch_, nm: Buf(n, xb.d, xb.h, xb.w, ch_, dev, f"{name}.{nm}")
        g = T(c, "gn")
        self._gn_act(x, ab.norm, ops.ACT_NONE, g)
        qkv = T(3 * c, "qkv")
        for idx, lin in enumerate((ab.to_q, ab.to_k, ab.to_v)):
            self._linear(g.sl(), lin, f"{name}.qkv{idx}", out=qkv.sl(idx * c, c))
        o = T(c, "o")
        t.add(AttentionOp(qkv, o, heads, L))
        out = dst if dst is not None else T(c, "out").sl()
        t.add(NormActOp(o, "none", ops.ACT_NONE, [out], res=x))
        self._zero_params += [ab.proj_attn.weight, ab.proj_attn.bias]       # constructed, never applied (:383, :421-461)
        return out

    def _transformer(self, st, x: Sl, lvl: int, dst: Optional[Sl], name: str) -> Sl:
        """SpatialTransformer.forward (atten_unet_model.py:315-343): GroupNorm, proj_in, the BasicTransformerBlocks, proj_out, + x."""
        if isinstance(st, AttentionBlock):
            return self._attention_block(st, x, dst, name)
        t, dev, n = self.tape, self.dev, self.n
        c = x.c
        xb = x.buf
        L = xb.d * xb.h * xb.w
        heads = st.transformer_blocks[0].attn1.num_heads
        T = lambda ch_, nm: Buf(n, xb.d, xb.h, xb.w, ch_, dev, f"{name}.{nm}")
        g = T(c, "gn")
        self._gn_act(x, st.norm, ops.ACT_NONE, g)
        t0 = self._conv(g.sl(), st.proj_in.conv, ksize=1, stride=1, pad=0, name=name + ".proj_in").z
        for li, blk in enumerate(st.transformer_blocks):           # in sequence (:336-337)
            t0 = self._transformer_layer(blk, t0, heads, L, T, name if li == 0 else f"{name}.l{li}", "" if li == 0 else f"l{li}.")
        t3 = t0
        po = self._conv(t3.sl(), st.proj_out.conv, ksize=1, stride=1, pad=0, name=name + ".proj_out").z
        out = dst if dst is not None else T(c, "out").sl()
        t.add(NormActOp(po, "none", ops.ACT_NONE, [out], res=x))
        return out

    def _transformer_layer(self, blk: BasicTransformerBlock, t0: Buf, heads: int, L: int, T0, name: str, tag: str) -> Buf:
        """One BasicTransformerBlock (:225-235) on the token buffer ``t0``; returns the buffer holding its output."""
        t = self.tape
        T = lambda ch_, nm: T0(ch_, tag + nm)
        inner = t0.c
        n1 = T(inner, "n1")
        ln1 = LayerNormOp(t0, n1, blk.norm1)
        t.add(ln1)
        self._bind += [(ln1, "grad_gamma", blk.norm1.weight), (ln1, "grad_beta", blk.norm1.bias)]
        hd = blk.attn1.head_channels
        if hd == 32:
            qkv = T(3 * inner, "qkv")
            for idx, lin in enumerate((blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v)):
                self._linear(n1.sl(), lin, f"{name}.qkv{idx}", out=qkv.sl(idx * inner, inner))
            o = T(inner, "o")
            t.add(AttentionOp(qkv, o, heads, L))
            p = self._linear(o.sl(), blk.attn1.to_out[0], name + ".to_out").z
        else:
            # heads of 8 / 16 channels (the reference's default num_head_channels = 8): every head is zero-padded to the
            # attention kernel's 32 channels through the weight staging of to_q / to_k / to_v (zero output rows) and to_out
            # (zero input columns); zero channels change neither q k^T nor p v, the softmax scale stays 1 / sqrt(hd)
            wide = heads * 32
            at = [h_ * 32 + j for h_ in range(heads) for j in range(hd)]
            qkv = T(3 * wide, "qkv")
            for idx, lin in enumerate((blk.attn1.to_q, blk.attn1.to_k, blk.attn1.to_v)):
                self._linear(n1.sl(), lin, f"{name}.qkv{idx}", out=qkv.sl(idx * wide, wide), cout_pad=wide, out_index=at)
            o = T(wide, "o")
            t.add(AttentionOp(qkv, o, heads, L, true_head_dim=hd))
            p = self._conv(o.sl(), blk.attn1.to_out[0], ksize=1, stride=1, pad=0, name=name + ".to_out", in_index=at).z
        t1 = T(inner, "t1")
        t.add(NormActOp(p, "none", ops.ACT_NONE, [t1.sl()], res=t0.sl()))
        cb = CovariateBiasOp(t1, self, blk.attn2.to_v, blk.attn2.to_out[0], L)
        t.add(cb)
        self._bind += [(cb, "grad_wv", blk.attn2.to_v.weight), (cb, "grad_wo", blk.attn2.to_out[0].weight),
                       (cb, "grad_bo", blk.attn2.to_out[0].bias)]
        self._zero_params += [blk.attn2.to_q.weight, blk.attn2.to_k.weight, blk.norm2.weight, blk.norm2.bias]
        n3 = T(inner, "n3")
        ln3 = LayerNormOp(t1, n3, blk.norm3)
        t.add(ln3)
        self._bind += [(ln3, "grad_gamma", blk.norm3.weight), (ln3, "grad_beta", blk.norm3.bias)]
        hh = self._linear(n3.sl(), blk.ff.linear1, name + ".ff1").z
        gg = T(hh.c // 2, "geglu")
        t.add(GegluOp(hh, gg))
        f2 = self._linear(gg.sl(), blk.ff.linear2, name + ".ff2").z
        t3 = T(inner, "t3")
        t.add(NormActOp(f2, "none", ops.ACT_NONE, [t3.sl()], res=t1.sl()))
        return t3

    # ------------------------------------------------------------------------------------------------ run
    def grad_slots(self, out=None):
        grads = super().grad_slots(out)
        by_id = {id(p): g for p, g in zip(self.params, grads)}
        self.zero.slots = [by_id[id(p)] for p in self._zero_params]
        return grads

    def forward(self, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        n, _, D, H, W = self.shape
        self.generation += 1
        self.context = context
        check(lib.petsyn_concat_latent(ptr(x), ptr(x), ptr(self.inp.t), D * H * W, n, 0, self.CPAD, stream_ptr()),
              "concat_latent")
        self.tape.forward(self.training())
        check(lib.petsyn_take_channel0(ptr(self.head.zf), ptr(self.y), self.y.numel(), self.head.cout, stream_ptr()),
              "take_channel0")
        return self.y

    def backward(self, dy: torch.Tensor, out: Optional[Dict[int, torch.Tensor]] = None, on_ready=None):
        grads = self.grad_slots(out)
        check(lib.petsyn_put_channel0_grad(None, ptr(dy), ptr(self.head.zg), dy.numel(), self.head.cout, 0,
                                           stream_ptr()), "put_channel0_grad")
        self.run_backward(on_ready)
        if on_ready is not None:
            for p in self._zero_params:
                on_ready(p)
        return [g.clone() for g in grads] if out is None else []


# ======================================================================================================================
# pMCI / sMCI classifier for synthesize -> classify (BASELINE configs[4]) -- a LABELLED RESTATEMENT, parity unpinned.
# ======================================================================================================================
class DiffusionModelEncoder(nn.Module):
    """Encoder + fully connected head in the shape of ``DiffusionModelEncoder`` (vendored copy:
    ``unet/utils/atten_unet_model.py:1863-2032``; used as ``model(imgs, zeros_timesteps, info)`` at
    ``pet_for_classification/train_atten_encoder_MCI.py:87,169`` with ``config/training_atten.json``).

    **Restatement, parity unpinned.**  The class the reference scripts import lives in the authors'
    un-vendored fork; the vendored copy cannot run (``get_timestep_embedding`` is undefined, its ResnetBlocks take no time
    embedding, and ``Linear(4096, 512)`` does not match the 4 608 features a 96x128x96 input produces).  This class keeps
    what the vendored source does define -- constructor signature, argument checks, module tree and state-dict keys
    (``conv_in``, ``time_embed``, ``down_blocks.{i}.{attentions,resnets,downsampler}``, ``out.{0,3}``), every level ends in a
    down-sampling ResnetBlock because ``is_final_block = i == len(num_channels)`` is never true (:1966), flatten in NCDHW
    order, ``Linear -> ReLU -> Dropout(0.1) -> Linear`` -- and documents what it decides: ``timesteps`` is accepted and
    ignored (the scripts always pass zeros), ``time_embed`` holds parameters that nothing reads, and the head's input width is
    the constructor argument ``head_in_features`` (default 4096 as written; 4608 for the reference crop).
    Forward and backward: ``eval()`` + ``no_grad`` is what configs[4] runs; with gradients enabled the module is differentiable
    w.r.t. its parameters (the training step of train_atten_encoder_MCI.py:169-175: logits -> weighted cross-entropy ->
    backward -> Adam; ``time_embed`` receives exact zero gradients)."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int,
                 num_res_blocks: Sequence[int] | int = (2, 2, 2, 2), num_channels: Sequence[int] = (32, 64, 64, 64),
                 attention_levels: Sequence[bool] = (False, False, True, True), norm_num_groups: int = 32,
                 norm_eps: float = 1e-6, resblock_updown: bool = False, num_head_channels: int | Sequence[int] = 8,
                 with_conditioning: bool = False, transformer_num_layers: int = 1,
                 cross_attention_dim: int | None = None, num_class_embeds: int | None = None,
                 upcast_attention: bool = False, head_in_features: int = 4096) -> None:
        super().__init__()
        if with_conditioning is True and cross_attention_dim is None:
            raise ValueError("DiffusionModelEncoder expects dimension of the cross-attention conditioning "
                             "(cross_attention_dim) when using with_conditioning.")
        if cross_attention_dim is not None and with_conditioning is False:
            raise ValueError("DiffusionModelEncoder expects with_conditioning=True when specifying the cross_attention_dim.")
        if any((c % norm_num_groups) != 0 for c in num_channels):
            raise ValueError("DiffusionModelEncoder expects all num_channels being multiple of norm_num_groups")
        if len(num_channels) != len(attention_levels):
            raise ValueError("DiffusionModelEncoder expects num_channels being same size of attention_levels")
        n = len(num_channels)
        num_head_channels = _rep(num_head_channels, n)
        if len(num_head_channels) != n:
            raise ValueError("num_head_channels should have the same length as attention_levels.")
        num_res_blocks = _rep(num_res_blocks, n)
        if spatial_dims != 3 or in_channels != 1:
            raise NotImplementedError("petsyn DiffusionModelEncoder: 3-D, one input channel (PET only)")
        if not (resblock_updown and with_conditioning) or transformer_num_layers != 1 or num_class_embeds is not None:
            raise NotImplementedError("petsyn DiffusionModelEncoder implements the training_atten.json family")
        for lvl, a in enumerate(attention_levels):
            if a and num_head_channels[lvl] != 32:
                raise NotImplementedError("attention levels need num_head_channels == 32 (the attention kernel's head size)")
        self.cfg = dict(num_channels=list(num_channels), num_res_blocks=num_res_blocks,
                        attention_levels=list(attention_levels), num_head_channels=num_head_channels,
                        norm_num_groups=norm_num_groups, norm_eps=norm_eps, cross_attention_dim=cross_attention_dim)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.block_out_channels = num_channels
        self.with_conditioning = with_conditioning
        ch, g, e = list(num_channels), norm_num_groups, norm_eps
        self.conv_in = _Convolution(in_channels, ch[0], 3)
        ted = ch[0] * 4
        self.time_embed = nn.Sequential(nn.Linear(ch[0], ted), nn.SiLU(), nn.Linear(ted, ted))
        self.down_blocks = nn.ModuleList([])
        out_c = ch[0]
        for i in range(n):
            in_c, out_c = out_c, ch[i]
            attn = None
            if attention_levels[i]:
                attn = dict(heads=ch[i] // num_head_channels[i], head_channels=num_head_channels[i],
                            num_layers=transformer_num_layers, norm_num_groups=g, norm_eps=e,
                            cross_attention_dim=cross_attention_dim)
            self.down_blocks.append(_DownBlock(in_c, out_c, num_res_blocks[i], g, e, True, attn))
        self.out = nn.Sequential(nn.Linear(head_in_features, 512), nn.ReLU(), nn.Dropout(0.1),
                                 nn.Linear(512, out_channels))
        self._engines: Dict[Tuple, "_ClsEngine"] = ops.EngineCache()

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor | None = None, context: torch.Tensor | None = None,
                class_labels: torch.Tensor | None = None) -> torch.Tensor:
        if class_labels is not None:
            raise NotImplementedError("class labels are not used by the reference scripts")
        if context is None:
            raise ValueError("DiffusionModelEncoder(with_conditioning=True) needs the tabular context tensor")
        if not x.is_cuda:
            raise RuntimeError("petsyn DiffusionModelEncoder runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        ctx = context.reshape(x.shape[0], -1).contiguous().float()
        if ctx.shape[1] != self.cfg["cross_attention_dim"]:
            raise ValueError(f"context has {ctx.shape[1]} values, expected {self.cfg['cross_attention_dim']}")
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _ClsEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        if torch.is_grad_enabled() and any(p.requires_grad for p in eng.params):
            return _AttenFn.apply(eng, x, ctx, *eng.params)      # train_atten_encoder_MCI.py:169-175: predict -> CE -> backward
        return eng.forward(x, ctx).clone()


class _ClsEngine(_AttenEngine):
    """Op tape of the classifier for one input shape: the AttenUNet down path + two linears (inference only)."""

    def __init__(self, net: DiffusionModelEncoder, shape, dev):
        _EngineBase.__init__(self, net, dev)
        n, _, D, H, W = shape
        cfg = net.cfg
        ch = cfg["num_channels"]
        nl = len(ch)
        if D % (1 << nl) or H % (1 << nl) or W % (1 << nl):
            raise ValueError(f"spatial dims {D}x{H}x{W} must be divisible by {1 << nl} (every level down-samples)")
        self.shape, self.n = shape, n
        self.context = None
        self.zero = _ZeroGrad()
        self._zero_params = []
        t = self.tape
        self.inp = Buf(n, D, H, W, self.CPAD, dev, "cls.input")
        c_in = self._conv(self.inp.sl(), net.conv_in.conv, ksize=3, stride=1, pad=1, need_dx=False, name="cls.conv_in")
        hs: Sl = c_in.z.sl()
        for i, blk in enumerate(net.down_blocks):
            for j, rb in enumerate(blk.resnets):
                hs = self._resnet(rb, hs, i, None, f"cls.down{i}.{j}")
                if cfg["attention_levels"][i]:
                    hs = self._transformer(blk.attentions[j], hs, i, None, f"cls.down{i}.{j}.attn")
            hs = self._resnet(blk.downsampler, hs, i, None, f"cls.down{i}.ds", down=True)
        h = hs.buf
        assert hs.off == 0 and hs.c == h.c
        feats = (h.rows // n) * h.c
        lin1, lin2 = net.out[0], net.out[3]
        if feats != lin1.in_features:
            raise ValueError(f"the encoder produces {h.c} x {h.d}x{h.h}x{h.w} = {feats} features for this input but the head "
                             f"was built with head_in_features={lin1.in_features} (the vendored class hard-codes 4096, "
                             "atten_unet_model.py:1987; a 96x128x96 input needs 4608)")
        vox, fc = h.rows // n, h.c
        flat = h.alias(n, 1, 1, 1, feats, "cls.flat")
        # head (atten_unet_model.py:1987): Linear -> ReLU -> Dropout(0.1) -> Linear on nn.Flatten() of the NCDHW map.  Flatten
        # order is (c, voxel), the channels-last buffer is (voxel, c): input column c * vox + v of linear1 reads position
        # v * C + c -- a channel index map of the weight staging, like the padded attention heads.
        perm = [v * fc + c for c in range(fc) for v in range(vox)]
        l1 = self._conv(flat.sl(), lin1, ksize=1, stride=1, pad=0, name="cls.linear1", in_index=perm)
        hid = Buf(n, 1, 1, 1, 512, dev, "cls.hid")
        t.add(NormActOp(l1.z, "none", ops.ACT_RELU, [hid.sl()]))
        t.add(DropoutOp(hid, net.out[2]))
        self.head = self._conv(hid.sl(), lin2, ksize=1, stride=1, pad=0, y_fp32=True, name="cls.linear2")
        self.nout = lin2.out_features
        self._zero_params += list(net.time_embed.parameters())          # parameters nothing reads: exact zero gradients
        t.add(self.zero)
        self._finish()
        for p in self._zero_params:
            if all(p is not q for q in self.params):
                self.params.append(p)

    def forward(self, x: torch.Tensor, context: torch.Tensor) -> torch.Tensor:
        n, _, D, H, W = self.shape
        self.generation += 1
        self.context = context
        check(lib.petsyn_concat_latent(ptr(x), ptr(x), ptr(self.inp.t), D * H * W, n, 0, self.CPAD, stream_ptr()),
              "concat_latent")
        self.tape.forward(self.training())
        return self.head.zf[:, :self.nout]

    def backward(self, dlogits: torch.Tensor, out: Optional[Dict[int, torch.Tensor]] = None, on_ready=None):
        grads = self.grad_slots(out)
        self.head.zg.zero_()
        self.head.zg[:, :self.nout].copy_(dlogits)
        self.run_backward(on_ready)
        if on_ready is not None:
            for p in self._zero_params:
                on_ready(p)
        return [g.clone() for g in grads] if out is None else []
