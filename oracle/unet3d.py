"""Oracle (CPU, fp32) for the pix2pix-style 3D U-Net generator.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, as plain functional PyTorch on CPU, what the reference's
``UnetGenerator3d`` / ``UnetSkipConnectionBlock3d`` compute
(reference: ``unet/utils/unet_model.py:5-99``).  Parameters are addressed by the
reference's own state-dict keys, so a reference checkpoint drives the oracle
unchanged.  Pinned by ``tests/test_oracle_cpu.py::test_oracle_matches_live_reference`` (live reference) and
``tests/golden/unet3d_*.npz``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm3d default (unet_model.py:7 norm_layer default)
BN_MOMENTUM = 0.1
LRELU_SLOPE = 0.2   # unet_model.py:49


@dataclass
class Level:
    """One UnetSkipConnectionBlock3d, outermost first (unet_model.py:37-99)."""
    outer_nc: int
    inner_nc: int
    outermost: bool
    innermost: bool
    prefix: str          # state-dict prefix of the block's nn.Sequential ("model.model." ...)


def level_specs(input_nc: int, output_nc: int, num_downs: int, ngf: int = 64) -> List[Level]:
    """Widths of the nested blocks, outermost first.

    Follows the constructor at unet_model.py:14-25: innermost (8ngf, 8ngf), then
    ``num_downs-5`` blocks of (8ngf, 8ngf), then (4ngf, 8ngf), (2ngf, 4ngf) and either
    [(ngf, 2ngf), outermost(output_nc, ngf)] for num_downs >= 5 or
    outermost(output_nc, 2ngf) otherwise (the ``else`` branch at :23-24).
    """
    assert input_nc == output_nc  # unet_model.py:12
    inner_to_outer = [(ngf * 8, ngf * 8)]
    for _ in range(num_downs - 5):
        inner_to_outer.append((ngf * 8, ngf * 8))
    inner_to_outer.append((ngf * 4, ngf * 8))
    inner_to_outer.append((ngf * 2, ngf * 4))
    if num_downs >= 5:
        inner_to_outer.append((ngf, ngf * 2))
        inner_to_outer.append((output_nc, ngf))
    else:
        inner_to_outer.append((output_nc, ngf * 2))
    widths = inner_to_outer[::-1]
    levels: List[Level] = []
    prefix = "model.model."
    for i, (outer, inner) in enumerate(widths):
        outermost = i == 0
        innermost = i == len(widths) - 1
        levels.append(Level(outer, inner, outermost, innermost, prefix))
        # position of the submodule inside this block's Sequential:
        # outermost [downconv, sub, ...] -> 1; middle [lrelu, conv, norm, sub, ...] -> 3
        prefix = prefix + ("1.model." if outermost else "3.model.")
    return levels


def _slots(level: Level) -> Dict[str, Optional[int]]:
    """Index of each layer inside the block's nn.Sequential (unet_model.py:62-90)."""
    if level.outermost:
        return dict(downconv=0, downnorm=None, upconv=4, upnorm=None)
    if level.innermost:
        return dict(downconv=1, downnorm=None, upconv=4, upnorm=5)
    return dict(downconv=1, downnorm=2, upconv=6, upnorm=7)


def init_state_dict(input_nc: int = 1, output_nc: int = 1, num_downs: int = 4, ngf: int = 64,
                    seed: Optional[int] = 777) -> "OrderedDict[str, torch.Tensor]":
    """Seeded default initialisation, consuming the CPU RNG in the same order as the
    reference constructor (innermost block first; inside a block ``downconv`` then the up
    ``conv``, unet_model.py:47-60).  nn.Conv3d default = kaiming_uniform(a=sqrt(5)), no bias
    (BatchNorm => use_bias False, unet_model.py:42-45); BatchNorm3d = ones/zeros.
    """
    if seed is not None:
        torch.manual_seed(seed)
    levels = level_specs(input_nc, output_nc, num_downs, ngf)
    drawn: Dict[str, torch.Tensor] = {}
    for lv in reversed(levels):
        sl = _slots(lv)
        wd = torch.empty(lv.inner_nc, lv.outer_nc, 4, 4, 4)
        torch.nn.init.kaiming_uniform_(wd, a=math.sqrt(5))
        up_in = lv.inner_nc if lv.innermost else lv.inner_nc * 2
        wu = torch.empty(lv.outer_nc, up_in, 3, 3, 3)
        torch.nn.init.kaiming_uniform_(wu, a=math.sqrt(5))
        drawn[f"{lv.prefix}{sl['downconv']}.weight"] = wd
        drawn[f"{lv.prefix}{sl['upconv']}.weight"] = wu
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for key in state_dict_keys(input_nc, output_nc, num_downs, ngf):
        if key in drawn:
            sd[key] = drawn[key]
        else:
            ch = _bn_channels(levels, key)
            if key.endswith("running_var") or key.endswith(".weight"):
                sd[key] = torch.ones(ch)
            elif key.endswith("num_batches_tracked"):
                sd[key] = torch.zeros((), dtype=torch.long)
            else:
                sd[key] = torch.zeros(ch)
    return sd


def _bn_channels(levels: List[Level], key: str) -> int:
    for lv in levels:
        sl = _slots(lv)
        if sl["downnorm"] is not None and key.startswith(f"{lv.prefix}{sl['downnorm']}."):
            return lv.inner_nc
        if sl["upnorm"] is not None and key.startswith(f"{lv.prefix}{sl['upnorm']}."):
            return lv.outer_nc
    raise KeyError(key)


def state_dict_keys(input_nc: int = 1, output_nc: int = 1, num_downs: int = 4, ngf: int = 64) -> List[str]:
    """Key order of ``UnetGenerator3d(...).state_dict()`` (module registration order)."""
    levels = level_specs(input_nc, output_nc, num_downs, ngf)
    bn = ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"]

    def rec(i: int) -> List[str]:
        lv = levels[i]
        sl = _slots(lv)
        keys = [f"{lv.prefix}{sl['downconv']}.weight"]
        if sl["downnorm"] is not None:
            keys += [f"{lv.prefix}{sl['downnorm']}.{s}" for s in bn]
        if not lv.innermost:
            keys += rec(i + 1)
        keys.append(f"{lv.prefix}{sl['upconv']}.weight")
        if sl["upnorm"] is not None:
            keys += [f"{lv.prefix}{sl['upnorm']}.{s}" for s in bn]
        return keys

    return rec(0)


def _batchnorm(x, sd, prefix, training, new_buffers):
    """nn.BatchNorm3d forward: batch statistics (biased variance) in training mode with a
    momentum-0.1 running update that uses the unbiased variance; running statistics in eval."""
    w, b = sd[prefix + "weight"], sd[prefix + "bias"]
    rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
    if training:
        dims = (0, 2, 3, 4)
        mean = x.mean(dims)
        var = x.var(dims, unbiased=False)
        n = x.numel() // x.shape[1]
        if new_buffers is not None:
            with torch.no_grad():
                new_buffers[prefix + "running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean
                new_buffers[prefix + "running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (n / max(n - 1, 1))
                new_buffers[prefix + "num_batches_tracked"] = sd[prefix + "num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    shape = (1, -1, 1, 1, 1)
    return (x - mean.view(shape)) * torch.rsqrt(var.view(shape) + BN_EPS) * w.view(shape) + b.view(shape)


def _instancenorm(x, sd, prefix, eps=1e-5):
    """nn.InstanceNorm3d forward (track_running_stats=False: the same in train and eval): per (sample, channel) statistics
    over the volume, biased variance; affine only when the state dict carries weight / bias for it."""
    mean = x.mean((2, 3, 4), keepdim=True)
    var = x.var((2, 3, 4), unbiased=False, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    if prefix + "weight" in sd:
        shape = (1, -1, 1, 1, 1)
        y = y * sd[prefix + "weight"].view(shape) + sd[prefix + "bias"].view(shape)
    return y


def forward(x: torch.Tensor, sd: Dict[str, torch.Tensor], input_nc: int = 1, output_nc: int = 1,
            num_downs: int = 4, ngf: int = 64, training: bool = True,
            new_buffers: Optional[Dict[str, torch.Tensor]] = None, norm: str = "batch",
            dropout_masks: Optional[Dict[int, torch.Tensor]] = None) -> torch.Tensor:
    """y = UnetGenerator3d(x) (unet_model.py:27-32, 93-99), out of place.

    The reference's in-place LeakyReLU/ReLU (unet_model.py:49,51) make the skip half of every
    concat ``LeakyReLU(x)`` rather than ``x``; that is reproduced here explicitly
    (``skip = a`` below), without mutating the caller's tensor.

    ``norm="instance"``: norm_layer=nn.InstanceNorm3d, whose convolutions carry biases (unet_model.py:42-45; read from the
    state dict when present).  ``dropout_masks``: {level index (0 = outermost): mask already scaled by 1 / (1 - p)} applied
    where ``use_dropout=True`` puts nn.Dropout(0.5), after the up normalisation (unet_model.py:87-88); None = eval mode.
    """
    levels = level_specs(input_nc, output_nc, num_downs, ngf)

    def _norm(t, prefix):
        if norm == "instance":
            return _instancenorm(t, sd, prefix)
        return _batchnorm(t, sd, prefix, training, new_buffers)

    def block(i: int, xin: torch.Tensor) -> torch.Tensor:
        lv = levels[i]
        sl = _slots(lv)
        if lv.outermost:
            a = xin                                                    # no downrelu (:62)
        else:
            a = F.leaky_relu(xin, LRELU_SLOPE)                         # downrelu (:49), in place in the reference
        d = F.conv3d(a, sd[f"{lv.prefix}{sl['downconv']}.weight"], sd.get(f"{lv.prefix}{sl['downconv']}.bias"), stride=2,
                     padding=1)                                                                 # :47
        if sl["downnorm"] is not None:
            d = _norm(d, f"{lv.prefix}{sl['downnorm']}.")                                      # :50
        inner = d if lv.innermost else block(i + 1, d)
        r = F.relu(inner)                                              # uprelu (:51)
        u = F.interpolate(r, scale_factor=2, mode="nearest")           # nn.Upsample(scale_factor=2) (:59)
        u = F.conv3d(u, sd[f"{lv.prefix}{sl['upconv']}.weight"], sd.get(f"{lv.prefix}{sl['upconv']}.bias"), stride=1,
                     padding=1)                                                                 # :60
        if lv.outermost:
            return torch.tanh(u)                                       # :64
        u = _norm(u, f"{lv.prefix}{sl['upnorm']}.")                                            # :52
        if dropout_masks is not None and i in dropout_masks:
            u = u * dropout_masks[i]                                   # nn.Dropout(0.5) (:88)
        return torch.cat([u, a], 1)                                    # :99 (skip is the activated input)

    return block(0, x)


def train_step(x: torch.Tensor, target: torch.Tensor, sd: Dict[str, torch.Tensor], **cfg):
    """zero_grad -> forward -> nn.L1Loss -> backward (train_unet.py:106,149,167 with the
    perceptual and adversarial terms dropped).  Returns (loss, output, {key: grad})."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point
              and not k.endswith("running_mean") and not k.endswith("running_var")}
    full = dict(sd)
    full.update(params)
    new_buffers: Dict[str, torch.Tensor] = {}
    y = forward(x, full, training=True, new_buffers=new_buffers, **cfg)
    loss = (y - target).abs().mean()                                   # nn.L1Loss()
    loss.backward()
    grads = {k: p.grad for k, p in params.items()}
    return loss.detach(), y.detach(), grads, new_buffers


def conv_flops(shape, num_downs: int = 4, ngf: int = 64, nc: int = 1) -> float:
    """Direct-convolution forward FLOPs 2*M*Cout*Cin*k^3 summed over the convs."""
    n, _, d, h, w = shape
    levels = level_specs(nc, nc, num_downs, ngf)
    total = 0.0
    vox = d * h * w
    for i, lv in enumerate(levels):
        vox_out = vox // 8 ** (i + 1)
        total += 2.0 * n * vox_out * lv.inner_nc * lv.outer_nc * 64           # k4 s2 down conv
        up_in = lv.inner_nc if lv.innermost else 2 * lv.inner_nc
        total += 2.0 * n * (vox // 8 ** i) * lv.outer_nc * up_in * 27         # k3 conv on the upsampled grid
    return total


def randomize_(sd: Dict[str, torch.Tensor], seed: int) -> Dict[str, torch.Tensor]:
    """Fill a state dict in place with values drawn BY KEY (one generator per key, seeded from the key's position), so
    that the live reference (tests/golden/make_golden.py) and the module under test get the same parameters whatever
    their constructors drew -- used for the constructor families whose default initialisation the oracle does not restate
    (biased convolutions of the InstanceNorm3d family).  Convolution weights ~ N(0, 1/fan_in), biases and affine shifts
    ~ 0.1 N(0, 1), affine scales ~ 1 + 0.1 N(0, 1); buffers are left alone."""
    for i, (k, v) in enumerate(sd.items()):
        if not v.dtype.is_floating_point or k.endswith("running_mean") or k.endswith("running_var"):
            continue
        g = torch.Generator().manual_seed(seed * 1000 + i)
        r = torch.randn(v.shape, generator=g)
        if v.dim() == 5:
            r = r / math.sqrt(v[0].numel())
        elif k.endswith(".weight"):
            r = 1.0 + 0.1 * r
        else:
            r = 0.1 * r
        with torch.no_grad():
            v.copy_(r)
    return sd
