"""Oracle (CPU, fp32) for the BMGAN baseline networks (TEST INFRASTRUCTURE ONLY).

Restates ``bl_methods/BMGAN/bmgan_model.py`` -- ``get_dense_block`` :12-23, ``dense_unet_generator`` :25-101,
``ResNet_encoder`` :103-130, ``patch_discriminator`` :133-144 -- on top of ``oracle/monai_stub.py`` (the un-vendored
MONAI blocks; parity unpinned behind that boundary), with the reference's module names so state-dict keys match.
Also the in-repo loss pieces of the BMGAN step (``train_bmgan.py:33-40,141-200``).
Pinned by ``tests/test_oracle_cpu.py::test_bmgan_oracle_matches_live_reference`` (the reference file imported
unmodified over the same stubs) and ``tests/golden/bmgan_*.npz``.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn as nn

from .monai_stub import ConvDenseBlock, PatchDiscriminator, ResidualUnit

LRELU = ("leakyrelu", {"negative_slope": 0.2})


def dense_block(cin: int, c: int) -> List[nn.Module]:
    """bmgan_model.py:12-23: [dense(cin->c) ; conv3(cin+c->c) IN LReLU] twice (second time cin = c)."""
    mods: List[nn.Module] = []
    for a in (cin, c):
        mods += [ConvDenseBlock(spatial_dims=3, in_channels=a, channels=[c], num_res_units=1, act=LRELU),
                 nn.Conv3d(a + c, c, 3, padding=1), nn.InstanceNorm3d(c), nn.LeakyReLU(0.2)]
    return mods


def conv_in_lrelu(cin: int, cout: int, stride: int = 1) -> List[nn.Module]:
    return [nn.Conv3d(cin, cout, 3, padding=1, stride=stride), nn.InstanceNorm3d(cout), nn.LeakyReLU(0.2)]


class DenseUnetGenerator(nn.Module):
    """bmgan_model.py:25-101."""

    def __init__(self, input_channel=9, input_conv_channel=64, output_conv_channel=64, down_layers=5,
                 down_channels=(128, 256, 256, 512), middle_layers=1, middle_channels=(512,), up_layers=6,
                 up_channels=(512, 256, 256, 256, 128)):
        super().__init__()
        ic = input_conv_channel
        self.input_layer = nn.Sequential(*(conv_in_lrelu(input_channel, ic) + conv_in_lrelu(ic, ic)
                                           + conv_in_lrelu(ic, ic, stride=2)))                      # :34-38
        self.down_layers = nn.ModuleList()
        cur = ic
        for c in down_channels:                                                                    # :43-51
            self.down_layers.append(nn.Sequential(*(dense_block(cur, c) + conv_in_lrelu(c, c, stride=2))))
            cur = c
        self.middle_layers = nn.Sequential(*dense_block(cur, middle_channels[-1]))                 # :53
        cur = middle_channels[-1]
        skips = [ic] + list(down_channels)
        self.up_layers = nn.ModuleList()
        for i, c in enumerate(up_channels):                                                        # :57-64
            self.up_layers.append(nn.Sequential(*(dense_block(cur + skips[-1 - i], c) + [
                nn.ConvTranspose3d(c, c, kernel_size=4, stride=2, padding=1), nn.InstanceNorm3d(c), nn.LeakyReLU(0.2)])))
            cur = c
        oc = output_conv_channel
        self.output_layer = nn.Sequential(*(conv_in_lrelu(cur, oc) + conv_in_lrelu(oc, oc)
                                            + [nn.Conv3d(oc, 1, 3, padding=1), nn.Tanh()]))       # :66-70

    def forward(self, x: torch.Tensor, sampled_latent_vector: torch.Tensor) -> torch.Tensor:
        n = x.shape[0]
        z = sampled_latent_vector.view(n, -1)[:, :, None, None, None].expand(-1, -1, *x.shape[2:])  # :76-77
        feat = self.input_layer(torch.cat([x, z], 1))                                               # :79-81
        skips = [feat]
        for blk in self.down_layers:                                                                # :83-86
            feat = blk(feat)
            skips.append(feat)
        feat = self.middle_layers(feat)                                                             # :91
        for i, blk in enumerate(self.up_layers):                                                    # :93-96
            feat = blk(torch.cat([feat, skips[-1 - i]], 1))
        return self.output_layer(feat)                                                              # :99


class ResNetEncoder(nn.Module):
    """bmgan_model.py:103-130."""

    def __init__(self, input_layer_channel=32, channels=(64, 128, 128, 128, 128, 128)):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv3d(1, input_layer_channel, 3, padding=1),
                                         nn.InstanceNorm3d(input_layer_channel), nn.ReLU())
        self.resblocks = nn.ModuleList()
        cur = input_layer_channel
        for c in channels:
            self.resblocks.append(ResidualUnit(3, cur, c, strides=2, padding=1))
            cur = c
        self.linear1 = nn.Linear(128 * 8, 8)
        self.linear2 = nn.Linear(128 * 8, 8)

    def forward(self, x):
        out = self.input_layer(x)
        for blk in self.resblocks:
            out = blk(out)
        out = out.flatten(1)
        return self.linear1(out), self.linear2(out)


class PatchDiscriminatorWrapper(nn.Module):
    """bmgan_model.py:133-144: PatchDiscriminator(3, 32, 1, num_layers_d=4); forward returns the last stage output."""

    def __init__(self):
        super().__init__()
        self.patch_d = PatchDiscriminator(3, 32, 1, num_layers_d=4)

    def forward(self, x):
        return self.patch_d(x)[-1]


def lsgan_loss(logits: torch.Tensor, target_is_real: bool) -> torch.Tensor:
    """PatchAdversarialLoss(criterion="least_squares"): MSE against a constant 1/0 target (train_bmgan.py:152)."""
    return ((logits - (1.0 if target_is_real else 0.0)) ** 2).mean()


def kl_divergence(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    """train_bmgan.py:33-40: -0.5 * sum(1 + logvar - mu^2 - exp(logvar)) over the last dim."""
    return -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=-1)


def generator_step(gen: nn.Module, disc: nn.Module, t1, pet, z, lamda_l1: float = 20.0):
    """G phase of train_bmgan.py:141-161 with the perceptual term dropped (needs downloaded LPIPS weights):
    loss = LSGAN(D(G(t1,z))[-1], real) + lamda_l1 * L1(G(t1,z), pet); D frozen.  Returns (loss, adv, l1, fake)."""
    for p in disc.parameters():
        p.requires_grad_(False)
    fake = gen(t1, z)
    logits = disc(fake.contiguous().float())[-1]        # the reference indexes the wrapper's output again
    adv = lsgan_loss(logits, True)
    l1 = (fake - pet).abs().mean()
    loss = adv + lamda_l1 * l1
    return loss, adv, l1, fake


def discriminator_step(disc: nn.Module, fake, real):
    """D phase of train_bmgan.py:183-200: two backward calls, gradients accumulate, no optimiser step."""
    for p in disc.parameters():
        p.requires_grad_(True)
    lf = lsgan_loss(disc(fake.contiguous().detach())[-1], False)
    lf.backward()
    lr = lsgan_loss(disc(real.contiguous().detach())[-1], True)
    lr.backward()
    return 0.5 * (lf + lr)
