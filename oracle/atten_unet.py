"""Oracle (CPU, fp32) for the covariate-conditioned generator ``AttenUNet`` (TEST INFRASTRUCTURE ONLY).

Functional restatement of ``unet/utils/atten_unet_model.py`` driven by the reference's state-dict keys:
``ResnetBlock`` :565-662, ``Downsample``/``Upsample`` :464-562 (``resblock_updown`` form), ``DownBlock`` /
``CrossAttnDownBlock`` :665-967, ``CrossAttnMidBlock`` :1037-1100, ``UpBlock`` / ``CrossAttnUpBlock`` :1103-1409,
``SpatialTransformer`` :238-343, ``BasicTransformerBlock`` :178-235, ``CrossAttention`` :65-175, ``AttenUNet.forward``
:1792-1860.  The two MONAI pieces the file imports are restated per upstream: ``Convolution(conv_only=True)`` =
``nn.Conv3d`` child named ``conv``; ``MLPBlock(act="GEGLU")`` = linear1 -> chunk -> x*gelu(gate) -> linear2.
Pinned by ``tests/test_oracle_cpu.py::test_atten_unet_oracle_matches_live_reference`` and ``::test_atten_unet_oracle_matches_golden`` (the reference class imported
unmodified over a stub ``monai``) and ``tests/golden/atten_unet_*.npz``.
"""
from __future__ import annotations

import zlib
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

# the reference's own smoke configuration (unet/utils/atten_unet_model.py:2034-2051): conv-form down / up sampling
# (resblock_updown=False), the default 8-channel attention heads, 8 GroupNorm groups, a 2-D context of 3 covariates
SMOKE_CFG = dict(spatial_dims=3, in_channels=1, out_channels=1, cross_attention_dim=3, with_conditioning=True,
                 num_res_blocks=(1, 1, 1), num_channels=(8, 16, 16), norm_num_groups=8, attention_levels=[False, False, True],
                 norm_eps=1e-6, resblock_updown=False, num_head_channels=8)

TRAINING_JSON = dict(  # unet/config/training.json:8-38 (atten_unet_def) + cross_attention_dim injected at train_unet.py:64-68
    spatial_dims=3, in_channels=1, out_channels=1, num_channels=[16, 32, 64, 128], num_res_blocks=2,
    attention_levels=[False, False, False, True], norm_num_groups=16, norm_eps=1e-6, resblock_updown=True,
    num_head_channels=[0, 0, 0, 32], with_conditioning=True, transformer_num_layers=1, upcast_attention=False,
    use_flash_attention=False, cross_attention_dim=5)


# with_conditioning = False: the Attn{Down,Mid,Up}Block family (:751-852, :970-1029, :1190-1293) -- AttentionBlock (:346-461) instead
# of the SpatialTransformer, no context
# transformer_num_layers = 2: two BasicTransformerBlocks chained inside every SpatialTransformer (:300-301, :336-337)
TWO_LAYER_CFG = dict(spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(32, 32, 64),
                     attention_levels=(False, True, True), norm_num_groups=16, norm_eps=1e-6, resblock_updown=True,
                     num_head_channels=(0, 32, 32), with_conditioning=True, transformer_num_layers=2, cross_attention_dim=5)
ATTN_ONLY_CFG = dict(spatial_dims=3, in_channels=1, out_channels=1, num_res_blocks=1, num_channels=(32, 32, 64),
                     attention_levels=(False, True, True), norm_num_groups=16, norm_eps=1e-6, resblock_updown=True,
                     num_head_channels=32, with_conditioning=False, cross_attention_dim=None)


def randomize_(named_tensors, seed: int = 0) -> None:
    """Deterministic, construction-order-independent re-draw of every parameter (by NAME), so that the reference, the
    oracle and the CUDA modules can be given identical weights -- and so that the reference's ``zero_module`` tensors
    (conv2 / proj_out / out conv, atten_unet_model.py:56-62,303,616,1779), which make the default-initialised network
    output exactly 0, are non-zero."""
    with torch.no_grad():
        for name, p in named_tensors:
            g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
            r = torch.randn(p.shape, generator=g)
            if p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(r / fan_in ** 0.5)
            elif "norm" in name and name.endswith("weight") or name.endswith("out.0.weight"):
                p.copy_(1.0 + 0.1 * r)
            else:
                p.copy_(0.1 * r)


def _gn(sd, pre, x, groups, eps):
    return F.group_norm(x, groups, sd[pre + "weight"], sd[pre + "bias"], eps)


def _conv(sd, pre, x, pad):
    return F.conv3d(x, sd[pre + "conv.weight"], sd[pre + "conv.bias"], padding=pad)


def resnet(sd, pre, x, groups, eps, up=False, down=False):
    """ResnetBlock.forward (:641-662)."""
    h = F.silu(_gn(sd, pre + "norm1.", x, groups, eps))
    if up:
        x, h = (F.interpolate(t, scale_factor=2.0, mode="nearest") for t in (x, h))
    elif down:
        x, h = (F.avg_pool3d(t, 2, 2) for t in (x, h))
    h = _conv(sd, pre + "conv1.", h, 1)
    h = F.silu(_gn(sd, pre + "norm2.", h, groups, eps))
    h = _conv(sd, pre + "conv2.", h, 1)
    if pre + "skip_connection.conv.weight" in sd:
        x = _conv(sd, pre + "skip_connection.", x, 0)
    return x + h


def cross_attention(sd, pre, x, ctx, heads):
    """CrossAttention.forward (:156-175) with the naive baddbmm/softmax/bmm attention (:137-154)."""
    q = F.linear(x, sd[pre + "to_q.weight"])
    k = F.linear(ctx, sd[pre + "to_k.weight"])
    v = F.linear(ctx, sd[pre + "to_v.weight"])

    def split(t):
        b, l, d = t.shape
        return t.reshape(b, l, heads, d // heads).permute(0, 2, 1, 3).reshape(b * heads, l, d // heads)

    q, k, v = split(q), split(k), split(v)
    scale = 1.0 / (q.shape[-1] ** 0.5)
    probs = (torch.bmm(q, k.transpose(1, 2)) * scale).softmax(-1)
    o = torch.bmm(probs, v)
    bh, l, d = o.shape
    o = o.reshape(bh // heads, heads, l, d).permute(0, 2, 1, 3).reshape(bh // heads, l, d * heads)
    return F.linear(o, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])


def attention_block(sd, pre, x, groups, eps, heads):
    """AttentionBlock.forward (:421-461): GroupNorm -> to_q / to_k / to_v (with bias) -> multi-head attention -> + x.
    ``proj_attn`` is constructed (:383) but never applied by the reference's forward: its parameters get no gradient."""
    n, c, d, h, w = x.shape
    t = _gn(sd, pre + "norm.", x, groups, eps).view(n, c, d * h * w).transpose(1, 2)
    q, k, v = (F.linear(t, sd[pre + nm + ".weight"], sd[pre + nm + ".bias"]) for nm in ("to_q", "to_k", "to_v"))

    def split(u):
        b, l, dim = u.shape
        return u.reshape(b, l, heads, dim // heads).permute(0, 2, 1, 3).reshape(b * heads, l, dim // heads)

    q, k, v = split(q), split(k), split(v)
    scale = 1.0 / ((c / heads) ** 0.5)
    o = torch.bmm((torch.bmm(q, k.transpose(1, 2)) * scale).softmax(-1), v)
    bh, l, dh = o.shape
    o = o.reshape(bh // heads, heads, l, dh).permute(0, 2, 1, 3).reshape(bh // heads, l, dh * heads)
    return o.transpose(-1, -2).reshape(n, c, d, h, w) + x


def transformer(sd, pre, x, ctx, groups, eps, heads):
    """SpatialTransformer.forward (:315-343): the BasicTransformerBlocks (:225-235) in sequence (:336-337), as many as the
    state dict holds (transformer_num_layers)."""
    n, c, d, h, w = x.shape
    res = x
    t = _conv(sd, pre + "proj_in.", _gn(sd, pre + "norm.", x, groups, eps), 0)
    t = t.permute(0, 2, 3, 4, 1).reshape(n, d * h * w, -1)
    layer = 0
    while pre + f"transformer_blocks.{layer}.norm1.weight" in sd:
        b = pre + f"transformer_blocks.{layer}."
        ln = lambda name, v: F.layer_norm(v, (v.shape[-1],), sd[b + name + ".weight"], sd[b + name + ".bias"])
        t = cross_attention(sd, b + "attn1.", ln("norm1", t), ln("norm1", t), heads) + t
        t = cross_attention(sd, b + "attn2.", ln("norm2", t), ctx, heads) + t
        y = F.linear(ln("norm3", t), sd[b + "ff.linear1.weight"], sd[b + "ff.linear1.bias"])
        a, gate = y.chunk(2, dim=-1)
        t = F.linear(a * F.gelu(gate), sd[b + "ff.linear2.weight"], sd[b + "ff.linear2.bias"]) + t
        layer += 1
    t = t.reshape(n, d, h, w, -1).permute(0, 4, 1, 2, 3).contiguous()
    return _conv(sd, pre + "proj_out.", t, 0) + res


def forward(x: torch.Tensor, context: torch.Tensor, sd: Dict[str, torch.Tensor], cfg=TRAINING_JSON) -> torch.Tensor:
    """AttenUNet.forward (:1792-1860) for the resblock_updown / with_conditioning configuration."""
    ch: Sequence[int] = cfg["num_channels"]
    nres = cfg["num_res_blocks"]
    nres = [nres] * len(ch) if isinstance(nres, int) else list(nres)
    att = cfg["attention_levels"]
    hc = cfg["num_head_channels"]
    hc = [hc] * len(ch) if isinstance(hc, int) else list(hc)
    groups, eps = cfg["norm_num_groups"], cfg.get("norm_eps", 1e-6)
    updown = cfg.get("resblock_updown", False)
    cond = cfg["with_conditioning"]
    if cond and context.dim() < 3:
        context = context.unsqueeze(1)                                            # :110-112
    assert cond or context is None                                                # :1822-1823
    heads = lambda lvl: ch[lvl] // hc[lvl]
    if not cond:                                                                  # AttentionBlock instead of SpatialTransformer
        transformer = lambda sd_, pre, x_, ctx_, g_, e_, nh: attention_block(sd_, pre, x_, g_, e_, nh)   # noqa: F811
    else:
        transformer = globals()["transformer"]
    h = _conv(sd, "conv_in.", x, 1)
    skips: List[torch.Tensor] = [h]
    for i in range(len(ch)):
        for j in range(nres[i]):
            h = resnet(sd, f"down_blocks.{i}.resnets.{j}.", h, groups, eps)
            if att[i]:
                h = transformer(sd, f"down_blocks.{i}.attentions.{j}.", h, context, groups, eps, heads(i))
            skips.append(h)
        if i != len(ch) - 1:
            if updown:
                h = resnet(sd, f"down_blocks.{i}.downsampler.", h, groups, eps, down=True)
            else:                                                                   # Downsample(use_conv=True) :464-507
                pre = f"down_blocks.{i}.downsampler.op."
                h = F.conv3d(h, sd[pre + "conv.weight"], sd[pre + "conv.bias"], stride=2, padding=1)
            skips.append(h)
    h = resnet(sd, "middle_block.resnet_1.", h, groups, eps)
    h = transformer(sd, "middle_block.attention.", h, context, groups, eps, heads(len(ch) - 1))
    h = resnet(sd, "middle_block.resnet_2.", h, groups, eps)
    for i in range(len(ch)):
        lvl = len(ch) - 1 - i
        for j in range(nres[lvl] + 1):
            h = torch.cat([h, skips.pop()], 1)
            h = resnet(sd, f"up_blocks.{i}.resnets.{j}.", h, groups, eps)
            if att[lvl]:
                h = transformer(sd, f"up_blocks.{i}.attentions.{j}.", h, context, groups, eps, heads(lvl))
        if i != len(ch) - 1:
            if updown:
                h = resnet(sd, f"up_blocks.{i}.upsampler.", h, groups, eps, up=True)
            else:                                                                   # Upsample(use_conv=True) :510-562
                h = _conv(sd, f"up_blocks.{i}.upsampler.conv.", F.interpolate(h, scale_factor=2.0, mode="nearest"), 1)
    h = F.silu(F.group_norm(h, groups, sd["out.0.weight"], sd["out.0.bias"], eps))
    return F.conv3d(h, sd["out.2.conv.weight"], sd["out.2.conv.bias"], padding=1)


def train_step(x, context, target, sd, cfg=TRAINING_JSON):
    """fwd -> nn.L1Loss -> bwd (train_unet.py:147-149,167 with the perceptual/adversarial terms dropped)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    y = forward(x, context, params, cfg)
    loss = (y - target).abs().mean()
    loss.backward()
    return loss.detach(), y.detach(), {k: p.grad if p.grad is not None else torch.zeros_like(p) for k, p in params.items()}


def param_shapes(cfg=TRAINING_JSON) -> Dict[str, tuple]:
    """State-dict keys and shapes of ``AttenUNet(**cfg)`` (resblock_updown / with_conditioning family) in the reference's
    registration order, derived from the constructor walk (atten_unet_model.py:1676-1790) -- lets the oracle run without
    any module object (``tests/test_oracle_cpu.py`` checks it against the live reference class)."""
    ch: Sequence[int] = cfg["num_channels"]
    n = len(ch)
    nres = cfg["num_res_blocks"]
    nres = [nres] * n if isinstance(nres, int) else list(nres)
    att, cdim = cfg["attention_levels"], cfg["cross_attention_dim"]
    updown = cfg.get("resblock_updown", False)
    out: Dict[str, tuple] = {}

    def conv(pre, cin, cout, k):
        out[pre + "conv.weight"], out[pre + "conv.bias"] = (cout, cin, k, k, k), (cout,)

    def norm(pre, c):
        out[pre + "weight"], out[pre + "bias"] = (c,), (c,)

    def resnet(pre, cin, cout):
        norm(pre + "norm1.", cin)
        conv(pre + "conv1.", cin, cout, 3)
        norm(pre + "norm2.", cout)
        conv(pre + "conv2.", cout, cout, 3)
        if cin != cout:
            conv(pre + "skip_connection.", cin, cout, 1)

    def transformer(pre, c):
        norm(pre + "norm.", c)
        conv(pre + "proj_in.", c, c, 1)
        for layer in range(cfg.get("transformer_num_layers", 1)):
            b = pre + f"transformer_blocks.{layer}."
            for a, kv in (("attn1.", c), ("attn2.", cdim)):
                out[b + a + "to_q.weight"], out[b + a + "to_k.weight"], out[b + a + "to_v.weight"] = (c, c), (c, kv), (c, kv)
                out[b + a + "to_out.0.weight"], out[b + a + "to_out.0.bias"] = (c, c), (c,)
                if a == "attn1.":
                    out[b + "ff.linear1.weight"], out[b + "ff.linear1.bias"] = (8 * c, c), (8 * c,)
                    out[b + "ff.linear2.weight"], out[b + "ff.linear2.bias"] = (c, 4 * c), (c,)
            for nm in ("norm1.", "norm2.", "norm3."):
                norm(b + nm, c)
        conv(pre + "proj_out.", c, c, 1)

    if not cfg["with_conditioning"]:
        def transformer(pre, c):                          # AttentionBlock (:377-383)  # noqa: F811
            norm(pre + "norm.", c)
            for nm in ("to_q.", "to_k.", "to_v.", "proj_attn."):
                out[pre + nm + "weight"], out[pre + nm + "bias"] = (c, c), (c,)

    conv("conv_in.", 1, ch[0], 3)
    oc = ch[0]
    for i in range(n):
        ic, oc = oc, ch[i]
        pre = f"down_blocks.{i}."
        if att[i]:                                        # CrossAttnDownBlock registers attentions before resnets
            for j in range(nres[i]):
                transformer(pre + f"attentions.{j}.", oc)
        for j in range(nres[i]):
            resnet(pre + f"resnets.{j}.", ic if j == 0 else oc, oc)
        if i != n - 1:
            if updown:
                resnet(pre + "downsampler.", oc, oc)
            else:
                conv(pre + "downsampler.op.", oc, oc, 3)
    resnet("middle_block.resnet_1.", ch[-1], ch[-1])
    transformer("middle_block.attention.", ch[-1])
    resnet("middle_block.resnet_2.", ch[-1], ch[-1])
    rch = list(reversed(ch))
    oc = rch[0]
    for i in range(n):
        prev, oc = oc, rch[i]
        ic = rch[min(i + 1, n - 1)]
        lvl = n - 1 - i
        pre = f"up_blocks.{i}."
        nr = nres[lvl] + 1
        if att[lvl] and cfg["with_conditioning"]:         # CrossAttnUpBlock: attentions first (:1371-1372)
            for j in range(nr):
                transformer(pre + f"attentions.{j}.", oc)
        for j in range(nr):
            skip_c = ic if j == nr - 1 else oc
            in_c = prev if j == 0 else oc
            resnet(pre + f"resnets.{j}.", in_c + skip_c, oc)
        if att[lvl] and not cfg["with_conditioning"]:     # AttnUpBlock: resnets first (:1253-1254)
            for j in range(nr):
                transformer(pre + f"attentions.{j}.", oc)
        if i != n - 1:
            if updown:
                resnet(pre + "upsampler.", oc, oc)
            else:
                conv(pre + "upsampler.conv.", oc, oc, 3)
    norm("out.0.", ch[0])
    conv("out.2.", ch[0], 1, 3)
    return out


def init_state_dict(cfg=TRAINING_JSON, seed: int = 0) -> Dict[str, torch.Tensor]:
    """A full set of seeded, non-zero parameters (``randomize_`` by name) for ``forward`` / ``train_step``."""
    sd = {k: torch.zeros(s) for k, s in param_shapes(cfg).items()}
    randomize_(sd.items(), seed=seed)
    return sd


def adversarial_step(x, context, target, sd, disc, adv_weight: float = 0.1, base_lr: float = 5e-4, disc_lr: float = 1e-4,
                     cfg=TRAINING_JSON):
    """One full step of ``unet/scripts/train_unet.py:136-193`` with the terms that exist offline (LPIPS has weight 0 in
    unet/config/training.json:55): G phase -- D frozen, ``g_loss = L1 + adv_weight * LSGAN(D(G(t1, c))[-1], real)``, Adam(base_lr)
    -- then the D phase -- generator forward again under no_grad with the UPDATED weights, ``LSGAN(D(fake), fake).backward()``,
    ``LSGAN(D(real), real).backward()``, Adam(disc_lr).  ``disc``: a ``monai_stub.PatchDiscriminator`` (updated in place).
    Returns the losses, the generator / discriminator gradients of the step and the updated generator parameters."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    g_opt = torch.optim.Adam(list(params.values()), lr=base_lr)
    d_opt = torch.optim.Adam(disc.parameters(), lr=disc_lr)
    mse = lambda t, v: ((t - v) ** 2).mean()
    for p in disc.parameters():                                                     # requires_grad(discriminator, False) :136
        p.requires_grad_(False)
    y = forward(x, context, params, cfg)
    rec = (y - target).abs().mean()                                                 # :149
    adv = mse(disc(y.contiguous().float())[-1], 1.0)                                # :153-155
    g_loss = rec + adv_weight * adv                                                 # :159
    g_opt.zero_grad()
    g_loss.backward()
    g_grads = {k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for k, p in params.items()}
    g_opt.step()                                                                    # :166-168
    for p in disc.parameters():                                                     # :172-173
        p.requires_grad_(True)
    d_opt.zero_grad()
    with torch.no_grad():
        y2 = forward(x, context, params, cfg)                                       # :175-176
    d_fake = mse(disc(y2.contiguous().detach())[-1], 0.0)                           # :179-181
    d_fake.backward()
    d_real = mse(disc(target.contiguous().detach())[-1], 1.0)                       # :182-184
    d_real.backward()
    d_grads = {k: p.grad.clone() for k, p in disc.named_parameters()}
    d_opt.step()                                                                    # :193
    return dict(rec=rec.detach(), adv=adv.detach(), d_fake=d_fake.detach(), d_real=d_real.detach(), g_grads=g_grads,
                d_grads=d_grads, params={k: p.detach() for k, p in params.items()}, output=y.detach())
