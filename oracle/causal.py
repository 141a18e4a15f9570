"""Oracle (CPU, fp32) for the two decoders of the causal_synthesis model (TEST INFRASTRUCTURE ONLY) -- PARITY UNPINNED.

``DiffusionModelEncoder`` / ``Decoder`` / ``DiffusionModelDecoder`` of ``causal_synthesis/scripts/train_unify_causal_gen.py:5-7``
live in the authors' un-vendored ``monai_diffusion`` fork: there is no reference source to pin this file to.
It restates, in plain PyTorch, the SAME labelled restatement the product makes (``causal_model.py``): the T1 decoder after
upstream MONAI-GenerativeModels ``autoencoderkl.Decoder``; the PET decoder from the vendored blocks of
``unet/utils/atten_unet_model.py`` (ResnetBlock :565-662, SpatialTransformer :238-343, Upsample :510-562), driven by
``causal_synthesis/configs/training_causal.json:40-74``.  What the GPU tests prove with it is that the CUDA kernels compute
that graph -- not that the graph is the authors'.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from . import atten_unet as OA

T1_DECODER = dict(spatial_dims=3, in_channels=3, out_channels=1, num_channels=[32, 64, 64, 64], num_res_blocks=2,
                  norm_num_groups=32, norm_eps=1e-6, attention_levels=[False, False, False, False],
                  with_encoder_nonlocal_attn=False, with_decoder_nonlocal_attn=False)      # training_causal.json:40-60
PET_DECODER = dict(spatial_dims=3, in_channels=3, out_channels=1, num_channels=[64, 64, 32], num_res_blocks=2,
                   norm_num_groups=32, norm_eps=1e-6, attention_levels=[True, False, False], with_conditioning=True,
                   cross_attention_dim=5)                                                  # :62-74 + :114-116


def _resblock(sd, pre, x, groups, eps, skip="skip_connection."):
    h = F.silu(F.group_norm(x, groups, sd[pre + "norm1.weight"], sd[pre + "norm1.bias"], eps))
    h = OA._conv(sd, pre + "conv1.", h, 1)
    h = F.silu(F.group_norm(h, groups, sd[pre + "norm2.weight"], sd[pre + "norm2.bias"], eps))
    h = OA._conv(sd, pre + "conv2.", h, 1)
    if pre + skip + "conv.weight" in sd:
        x = OA._conv(sd, pre + skip, x, 0)
    return x + h


def _upsample(sd, pre, x):
    return OA._conv(sd, pre + "conv.", F.interpolate(x, scale_factor=2.0, mode="nearest"), 1)


def t1_decoder_forward(z: torch.Tensor, sd: Dict[str, torch.Tensor], cfg=T1_DECODER) -> torch.Tensor:
    """upstream autoencoderkl.Decoder.forward: ``for block in self.blocks: x = block(x)``."""
    ch: Sequence[int] = list(reversed(cfg["num_channels"]))
    nres = cfg["num_res_blocks"]
    nres = list(reversed([nres] * len(ch) if isinstance(nres, int) else list(nres)))
    g, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    x = OA._conv(sd, "blocks.0.", z, 1)
    k = 1
    for i in range(len(ch)):
        for _ in range(nres[i]):
            x = _resblock(sd, f"blocks.{k}.", x, g, eps, skip="nin_shortcut.")
            k += 1
        if i != len(ch) - 1:
            x = _upsample(sd, f"blocks.{k}.", x)
            k += 1
    x = F.group_norm(x, g, sd[f"blocks.{k}.weight"], sd[f"blocks.{k}.bias"], eps)          # no activation here (as upstream)
    return OA._conv(sd, f"blocks.{k + 1}.", x, 1)


def pet_decoder_forward(z: torch.Tensor, context: torch.Tensor, sd: Dict[str, torch.Tensor], cfg=PET_DECODER,
                        num_head_channels: int = 8) -> torch.Tensor:
    ch: Sequence[int] = cfg["num_channels"]
    nres = cfg["num_res_blocks"]
    nres = [nres] * len(ch) if isinstance(nres, int) else list(nres)
    g, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    if context.dim() < 3:
        context = context.unsqueeze(1)
    x = OA._conv(sd, "conv_in.", z, 1)
    for i in range(len(ch)):
        for j in range(nres[i]):
            x = _resblock(sd, f"up_blocks.{i}.resnets.{j}.", x, g, eps)
            if cfg["attention_levels"][i]:
                x = OA.transformer(sd, f"up_blocks.{i}.attentions.{j}.", x, context, g, eps, ch[i] // num_head_channels)
        x = _upsample(sd, f"up_blocks.{i}.upsampler.", x)
    x = F.silu(F.group_norm(x, g, sd["out.0.weight"], sd["out.0.bias"], eps))
    return OA._conv(sd, "out.2.", x, 1)


def kl_divergence(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """train_unify_causal_gen.py:57-73 (called with z_sigma as logvar at :228)."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp()) / mu.shape[0]
