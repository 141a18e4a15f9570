"""Restatement of the third-party blocks the reference imports but does not vendor (TEST INFRASTRUCTURE ONLY).

The reference pulls ``ConvDenseBlock`` / ``ResidualUnit`` from ``monai.networks.blocks``
(``bl_methods/BMGAN/bmgan_model.py:6``) and ``PatchDiscriminator`` from the authors' private
``monai_diffusion`` fork of MONAI-GenerativeModels (``bmgan_model.py:9``).  Neither package is installed here and no
version is pinned anywhere in the reference (no requirements file), so these classes restate the *published upstream*
behaviour (Project-MONAI/MONAI 1.x ``Convolution``/``ADN``/``ResidualUnit``/``ConvDenseBlock``;
Project-MONAI/GenerativeModels ``PatchDiscriminator``) with upstream's child-module names, so that state-dict keys come
out as upstream's.  PARITY UNPINNED behind this boundary: the reference holds no test or golden vector for it
(SURVEY 8c); what *is* pinned is that the reference's own ``bmgan_model.py`` imports and runs unmodified on top of it
and that ``oracle/bmgan.py`` agrees with that run bit for bit.
"""
from __future__ import annotations

import sys
import types
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn


def _act(act) -> nn.Module:
    name, kwargs = (act, {}) if isinstance(act, str) else (act[0], dict(act[1]))
    name = name.lower()
    if name == "prelu":
        return nn.PReLU(**kwargs)
    if name == "leakyrelu":
        return nn.LeakyReLU(**kwargs)
    if name == "relu":
        return nn.ReLU(**kwargs)
    raise ValueError(name)


def _norm(norm, channels: int) -> nn.Module:
    name = (norm if isinstance(norm, str) else norm[0]).lower()
    if name == "instance":
        return nn.InstanceNorm3d(channels)          # affine=False, track_running_stats=False (MONAI default)
    if name == "batch":
        return nn.BatchNorm3d(channels)
    raise ValueError(name)


class ADN(nn.Sequential):
    """Activation / Dropout / Normalisation in ``ordering`` (upstream default "NDA"); children named N, D, A."""

    def __init__(self, ordering="NDA", in_channels=None, act="RELU", norm=None, dropout=None):
        super().__init__()
        for item in ordering.upper():
            if item == "N" and norm is not None:
                self.add_module("N", _norm(norm, in_channels))
            elif item == "D" and dropout is not None:
                self.add_module("D", nn.Dropout3d(dropout))
            elif item == "A" and act is not None:
                self.add_module("A", _act(act))


class Convolution(nn.Sequential):
    """``conv`` (+ ``adn`` unless conv_only); padding defaults to "same" for odd kernels ((k-1)/2)."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dilation=1, bias=True, conv_only=False, padding=None):
        super().__init__()
        assert spatial_dims == 3
        if padding is None:
            padding = (kernel_size - 1) // 2 * dilation
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size, stride=strides, padding=padding,
                                          dilation=dilation, bias=bias))
        if conv_only:
            return
        if act is None and norm is None and dropout is None:
            return
        self.add_module("adn", ADN(adn_ordering, out_channels, act, norm, dropout))


class ResidualUnit(nn.Module):
    """``conv`` = Sequential(unit0..unit{s-1}) of Convolutions (first one strided); ``residual`` = Identity, or a conv
    (k = kernel_size when strided else 1) when the stride or the channel count changes; forward = conv(x) + residual(x)."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, subunits=2, adn_ordering="NDA",
                 act="PRELU", norm="INSTANCE", dropout=None, dilation=1, bias=True, last_conv_only=False, padding=None):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        if padding is None:
            padding = (kernel_size - 1) // 2 * dilation
        schannels, sstrides = in_channels, strides
        subunits = max(1, subunits)
        for su in range(subunits):
            conv_only = last_conv_only and su == subunits - 1
            self.conv.add_module(f"unit{su:d}", Convolution(spatial_dims, schannels, out_channels, strides=sstrides,
                                                            kernel_size=kernel_size, adn_ordering=adn_ordering, act=act,
                                                            norm=norm, dropout=dropout, dilation=dilation, bias=bias,
                                                            conv_only=conv_only, padding=padding))
            schannels, sstrides = out_channels, 1
        if strides != 1 or in_channels != out_channels:
            rk, rp = kernel_size, padding
            if strides == 1:
                rk, rp = 1, 0
            self.residual = nn.Conv3d(in_channels, out_channels, rk, strides, rp, bias=bias)

    def forward(self, x):
        return self.conv(x) + self.residual(x)


class ConvDenseBlock(nn.Sequential):
    """children ``layers{i}`` = ResidualUnit(l_in -> c, subunits=num_res_units); forward concatenates each layer's
    output to its input along channels."""

    def __init__(self, in_channels, channels: Sequence[int], spatial_dims=3, dilations=None, kernel_size=3,
                 num_res_units=0, adn_ordering="NDA", act="PRELU", norm="INSTANCE", dropout=None, bias=True):
        super().__init__()
        l_in = in_channels
        dilations = dilations if dilations is not None else [1] * len(channels)
        for i, (c, d) in enumerate(zip(channels, dilations)):
            if num_res_units > 0:
                layer = ResidualUnit(spatial_dims, l_in, c, strides=1, kernel_size=kernel_size, subunits=num_res_units,
                                     adn_ordering=adn_ordering, act=act, norm=norm, dropout=dropout, dilation=d, bias=bias)
            else:
                layer = Convolution(spatial_dims, l_in, c, strides=1, kernel_size=kernel_size, act=act, norm=norm,
                                    dropout=dropout, dilation=d, bias=bias)
            self.add_module(f"layers{i}", layer)
            l_in += c

    def forward(self, x):
        for layer in self.children():
            x = torch.cat([x, layer(x)], 1)
        return x


class PatchDiscriminator(nn.Sequential):
    """Pix2Pix PatchGAN of MONAI-GenerativeModels: ``initial_conv`` (k4 s2, bias, act, no norm), ``"0".."n-1"``
    (k4, stride 2 except the last, no bias, BatchNorm, LeakyReLU(0.2), channels doubling), ``final_conv`` (k4 s1,
    bias, conv only).  ``forward`` returns the list of all stage outputs.  Init: conv W ~ N(0, 0.02); BatchNorm
    gamma ~ N(1, 0.02), beta = 0."""

    def __init__(self, spatial_dims, num_channels, in_channels, out_channels=1, num_layers_d=3, kernel_size=4,
                 activation=("LEAKYRELU", {"negative_slope": 0.2}), norm="BATCH", bias=False, padding=1, dropout=0.0,
                 last_conv_kernel_size=None):
        super().__init__()
        if last_conv_kernel_size is None:
            last_conv_kernel_size = kernel_size
        self.num_layers_d = num_layers_d
        self.add_module("initial_conv", Convolution(spatial_dims, in_channels, num_channels, strides=2,
                                                    kernel_size=kernel_size, act=activation, norm=None, dropout=None,
                                                    bias=True, padding=padding))
        input_channels, output_channels = num_channels, num_channels * 2
        for l_ in range(num_layers_d):
            stride = 1 if l_ == num_layers_d - 1 else 2
            self.add_module(str(l_), Convolution(spatial_dims, input_channels, output_channels, strides=stride,
                                                 kernel_size=kernel_size, act=activation, norm=norm, dropout=None,
                                                 bias=bias, padding=padding))
            input_channels, output_channels = output_channels, output_channels * 2
        self.add_module("final_conv", Convolution(spatial_dims, input_channels, out_channels, strides=1,
                                                  kernel_size=last_conv_kernel_size, bias=True, conv_only=True,
                                                  padding=int((last_conv_kernel_size - 1) / 2)))
        self.apply(self.initialise_weights)

    def forward(self, x):
        out = [x]
        for submodel in self.children():
            out.append(submodel(out[-1]))
        return out[1:]

    @staticmethod
    def initialise_weights(m: nn.Module) -> None:
        cls = m.__class__.__name__
        if cls.find("Conv3d") != -1:
            nn.init.normal_(m.weight.data, 0.0, 0.02)
        elif cls.find("BatchNorm3d") != -1:
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0)


def install() -> None:
    """Register stub ``monai`` / ``monai_diffusion`` packages so that the reference's model files import unmodified."""
    def mod(name):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        return m

    monai = mod("monai")
    networks = mod("monai.networks")
    blocks = mod("monai.networks.blocks")
    monai.networks = networks
    networks.blocks = blocks
    blocks.ConvDenseBlock, blocks.ResidualUnit, blocks.Convolution, blocks.ADN = ConvDenseBlock, ResidualUnit, Convolution, ADN
    md = mod("monai_diffusion")
    gen = mod("monai_diffusion.generative")
    nets_pkg = mod("monai_diffusion.generative.networks")
    nets = mod("monai_diffusion.generative.networks.nets")
    md.generative, gen.networks, nets_pkg.nets = gen, nets_pkg, nets
    nets.PatchDiscriminator = PatchDiscriminator


# ------------------------------------------------------------------------------------------------ AttenUNet imports
class MLPBlock(nn.Module):
    """Upstream ``monai.networks.blocks.MLPBlock(hidden, mlp_dim, act="GEGLU")`` (atten_unet_model.py:40,211):
    linear1: hidden -> 2*mlp_dim; x, gate = chunk(2, -1); x * gelu(gate) (exact erf GELU); linear2: mlp_dim -> hidden."""

    def __init__(self, hidden_size, mlp_dim, dropout_rate=0.0, act="GELU", dropout_mode="vit"):
        super().__init__()
        assert act == "GEGLU"
        self.linear1 = nn.Linear(hidden_size, mlp_dim * 2)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.drop1, self.drop2 = nn.Dropout(dropout_rate), nn.Dropout(dropout_rate)

    def forward(self, x):
        x, gate = self.linear1(x).chunk(2, dim=-1)
        return self.drop2(self.linear2(self.drop1(x * nn.functional.gelu(gate))))


class _PoolFactory:
    """``monai.networks.layers.factories.Pool``: ``Pool[Pool.AVG, 3]`` is ``nn.AvgPool3d`` (atten_unet_model.py:41,498)."""
    AVG = "avg"

    def __getitem__(self, key):
        return nn.AvgPool3d


def install_atten() -> None:
    """Stub ``monai`` so that ``unet/utils/atten_unet_model.py`` imports unmodified (it uses Convolution with
    conv_only=True, MLPBlock(GEGLU), Pool and ensure_tuple_rep; :40-42)."""
    install()
    blocks = sys.modules["monai.networks.blocks"]
    blocks.MLPBlock = MLPBlock
    for name in ("monai.networks.layers", "monai.networks.layers.factories", "monai.utils"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["monai.networks.layers.factories"].Pool = _PoolFactory()
    sys.modules["monai.utils"].ensure_tuple_rep = lambda v, n: tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


# ------------------------------------------------------------------------------------------------ dataset imports
class SpatialPad:
    """Upstream ``monai.transforms.SpatialPad(spatial_size)`` (unet/utils/dataset.py:12,81), channel-first input, default
    ``method="symmetric"``, zero padding: per spatial axis width = max(target - size, 0), pad (width // 2, width - width // 2)."""

    def __init__(self, spatial_size):
        self.spatial_size = tuple(spatial_size)

    def __call__(self, img):
        pads = []
        for s, r in zip(img.shape[1:], self.spatial_size):
            width = max(r - s, 0)
            pads.append((width // 2, width - width // 2))
        flat = [p for pair in reversed(pads) for p in pair]
        return nn.functional.pad(img, flat)


class CenterSpatialCrop:
    """Upstream ``monai.transforms.CenterSpatialCrop(roi_size)`` (dataset.py:12,83): centre = size // 2,
    start = max(centre - roi // 2, 0), end = start + roi, per spatial axis of a channel-first input."""

    def __init__(self, roi_size):
        self.roi_size = tuple(roi_size)

    def __call__(self, img):
        sl = [slice(None)]
        for s, r in zip(img.shape[1:], self.roi_size):
            start = max(s // 2 - r // 2, 0)
            sl.append(slice(start, start + r))
        return img[tuple(sl)]


def install_dataset() -> None:
    """Stub ``monai.transforms`` and ``SimpleITK`` so that ``unet/utils/dataset.py`` imports unmodified (:9,12); only the
    two transforms ``_preprocess_img`` uses with crop=True are restated, the others raise if touched."""
    install()

    def _absent(*a, **k):
        raise NotImplementedError("not restated: only SpatialPad / CenterSpatialCrop are on the path")

    tr = types.ModuleType("monai.transforms")
    tr.SpatialPad, tr.CenterSpatialCrop, tr.Resize, tr.RandSpatialCrop = SpatialPad, CenterSpatialCrop, _absent, _absent
    sys.modules["monai.transforms"] = tr
    sys.modules["monai"].transforms = tr
    if "SimpleITK" not in sys.modules:
        sys.modules["SimpleITK"] = types.ModuleType("SimpleITK")
