"""CPU tests: the oracle against the committed golden fixtures (generated from the live reference), and -- when the
reference checkout is mounted -- against the reference modules themselves."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import unet3d as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = os.environ.get("PETSYN_REFERENCE", "/root/reference")


def synth_pair(shape, seed):
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    return torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, d, h, w, generator=g)


@pytest.mark.parametrize("name", ["unet3d_ngf8_1x32x48x32", "unet3d_ngf32_2x32x48x32"])
def test_oracle_matches_golden(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    ngf, shape, seed = int(gold["ngf"]), tuple(int(v) for v in gold["shape"]), int(gold["seed"])
    sd = O.init_state_dict(1, 1, 4, ngf, seed=seed)
    for k, v in sd.items():                                  # seeded init == the reference constructor's
        if v.dtype.is_floating_point:
            assert abs(float(v.double().abs().sum()) - float(gold["wsum/" + k])) <= 1e-6 * max(1.0, float(gold["wsum/" + k])), k
    t1, pet = synth_pair(shape, seed)
    loss, y, grads, bufs = O.train_step(t1, pet, sd, num_downs=4, ngf=ngf)
    assert abs(loss.item() - float(gold["loss"])) < 1e-6
    assert np.abs(y.numpy() - gold["output"]).max() < 1e-5
    for k, g in grads.items():
        ref = float(gold["gradnorm/" + k])
        assert abs(g.double().norm().item() - ref) <= 1e-4 * ref + 1e-9, k
    for k in gold.files:
        if k.startswith("buffer/"):
            assert np.abs(bufs[k[len("buffer/"):]].numpy() - gold[k]).max() < 1e-5, k
    # inference path (eval mode, updated running statistics)
    sd2 = dict(sd)
    sd2.update(bufs)
    ye = O.forward(t1, sd2, num_downs=4, ngf=ngf, training=False)
    assert np.abs(ye.numpy() - gold["output_eval"]).max() < 1e-4


def test_oracle_key_order_and_flops():
    keys = O.state_dict_keys(1, 1, 4, 64)
    assert keys[0] == "model.model.0.weight" and keys[-1] == "model.model.4.weight"
    assert "model.model.1.model.3.model.3.model.5.running_var" in keys
    # 839.8 GF forward per 96x112x96 volume
    assert abs(O.conv_flops((1, 1, 96, 112, 96)) / 1e9 - 839.8) < 0.5
    # num_downs = 5 takes the other constructor branch (unet_model.py:20-22)
    lv = O.level_specs(1, 1, 5, 64)
    assert [(l.outer_nc, l.inner_nc) for l in lv] == [(1, 64), (64, 128), (128, 256), (256, 512), (512, 512)]


def test_in_place_activation_aliasing_is_reproduced():
    """the skip half of the innermost concat is LeakyReLU(x), i.e. the BN output passed through the
    reference's in-place downrelu."""
    sd = O.init_state_dict(1, 1, 4, 8, seed=1)
    x = torch.randn(1, 1, 16, 16, 16)
    levels = O.level_specs(1, 1, 4, 8)
    assert levels[-1].innermost and levels[0].outermost
    y = O.forward(x, sd, num_downs=4, ngf=8, training=True)
    assert y.shape == x.shape and float(y.abs().max()) <= 1.0


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "unet", "utils")), reason="reference checkout not mounted")
def test_oracle_matches_live_reference():
    sys.path.insert(0, REF)
    from unet.utils.unet_model import UnetGenerator3d
    torch.manual_seed(5)
    ref = UnetGenerator3d(1, 1, num_downs=4, ngf=8).train()
    sd = O.init_state_dict(1, 1, 4, 8, seed=5)
    rsd = ref.state_dict()
    assert list(rsd.keys()) == list(sd.keys())
    assert all(torch.equal(rsd[k], sd[k]) for k in sd)
    t1, pet = synth_pair((2, 16, 32, 16), 9)
    y = ref(t1.clone())
    loss = torch.nn.L1Loss()(y, pet)
    loss.backward()
    lo, yo, go, bo = O.train_step(t1, pet, sd, num_downs=4, ngf=8)
    assert torch.allclose(y, yo, atol=1e-6) and abs(float(loss) - float(lo)) < 1e-7
    for k, p in ref.named_parameters():
        assert torch.allclose(p.grad, go[k], atol=1e-6, rtol=1e-5), k
    for k, v in bo.items():
        assert torch.allclose(ref.state_dict()[k].float(), v.float(), atol=1e-6), k


# ---------------------------------------------------------------------------------------------------------------------
# AttenUNet / BMGAN oracles: pinned to the fixtures generated from the reference classes (make_golden_atten.py /
# make_golden_bmgan.py) here on the CPU, and -- marker `reference` -- to the live classes imported over the MONAI stub.
# ---------------------------------------------------------------------------------------------------------------------
def _atten_inputs(shape, seed, cdim):
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    return (torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, cdim, generator=g), torch.rand(n, 1, d, h, w, generator=g))


@pytest.mark.parametrize("name,cfg_name", [("atten_unet_2x32x48x32", "TRAINING_JSON"), ("atten_unet_smoke_1x44x64x44", "SMOKE_CFG"),
                                           ("atten_unet_attnonly_1x32x48x32", "ATTN_ONLY_CFG"),
                                           ("atten_unet_twolayer_1x32x48x32", "TWO_LAYER_CFG")])
def test_atten_unet_oracle_matches_golden(name, cfg_name):
    from oracle import atten_unet as OA
    cfg = getattr(OA, cfg_name)
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    shape, seed = tuple(int(v) for v in gold["shape"]), int(gold["seed"])
    st = int(gold["stride"]) if "stride" in gold.files else 1
    sd = OA.init_state_dict(cfg, seed=seed)                       # randomize_ by NAME == what the fixture script drew
    for k, v in sd.items():
        ref = float(gold["wsum/" + k])
        assert abs(float(v.double().abs().sum()) - ref) <= 1e-6 * max(1.0, ref), k
    x, ctx, tgt = _atten_inputs(shape, seed, cfg["cross_attention_dim"] or 1)
    if cfg_name == "SMOKE_CFG":
        ctx = ctx[:, 0]                                            # 2-D context (atten_unet_model.py:110-112)
    if not cfg["with_conditioning"]:
        ctx = None                                                 # AttentionBlock family: no context (:1822-1823)
    loss, y, grads = OA.train_step(x, ctx, tgt, sd, cfg)
    assert abs(loss.item() - float(gold["loss"])) < 1e-6
    assert np.abs(y.numpy()[:, :, ::st, ::st, ::st] - gold["output"]).max() < 1e-5
    tot = 0.0
    for k, g in grads.items():
        ref = float(gold["gradnorm/" + k])
        tot += g.double().norm().item() ** 2
        # (biases in front of a GroupNorm have a mathematically zero gradient: what is left is fp32 rounding noise)
        assert abs(g.double().norm().item() - ref) <= 1e-4 * ref + 1e-7 * float(gold["grad_norm_total"]), k
    assert abs(tot ** 0.5 - float(gold["grad_norm_total"])) <= 1e-5 * float(gold["grad_norm_total"])


def test_atten_unet_oracle_shapes_and_zero_init_quirk():
    from oracle import atten_unet as OA
    shapes = OA.param_shapes()
    assert len(shapes) == 416 and sum(int(np.prod(s)) for s in shapes.values()) == 12562945
    assert shapes["down_blocks.3.attentions.0.transformer_blocks.0.attn2.to_k.weight"] == (128, 5)
    assert "down_blocks.0.downsampler.op.conv.weight" in OA.param_shapes(OA.SMOKE_CFG)


def test_bmgan_oracle_matches_golden():
    """Forward of the small dense U-Net generator + both loss terms against the fixture generated from the reference's
    bmgan_model.py (the backward of this case is covered on the GPU box, where the test time is not a concern)."""
    from oracle import bmgan as OB
    SMALL = dict(input_conv_channel=64, output_conv_channel=64, down_channels=[64, 128, 128, 128], middle_channels=[128],
                 up_channels=[128, 128, 128, 128, 64])
    gold = np.load(os.path.join(GOLD, "bmgan_small_2x64x96x64.npz"))
    shape, seed = tuple(int(v) for v in gold["shape"]), int(gold["seed"])
    torch.manual_seed(seed)
    gen, disc = OB.DenseUnetGenerator(**SMALL).train(), OB.PatchDiscriminatorWrapper().train()
    for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in disc.state_dict().items()]:
        if v.dtype.is_floating_point:                             # same construction order => same seeded init
            ref = float(gold["wsum/" + k])
            assert abs(float(v.double().abs().sum()) - ref) <= 1e-6 * max(1.0, ref), k
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    t1, pet = torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, d, h, w, generator=g) * 2 - 1
    z = torch.randn(n, 8, generator=g)
    with torch.no_grad():
        loss, adv, l1, fake = OB.generator_step(gen, disc, t1, pet, z)
    assert abs(loss.item() - float(gold["g_loss"])) <= 1e-5 * abs(float(gold["g_loss"]))
    assert abs(adv.item() - float(gold["g_adv"])) <= 1e-5 * abs(float(gold["g_adv"]))
    assert np.abs(fake.numpy()[:, :, ::2, ::2, ::2] - gold["fake_sample"]).max() < 1e-4


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "unet", "utils")), reason="reference checkout not mounted")
@pytest.mark.parametrize("cfg_name", ["TRAINING_JSON", "SMOKE_CFG", "ATTN_ONLY_CFG"])
def test_atten_unet_oracle_matches_live_reference(cfg_name):
    from oracle import atten_unet as OA
    from oracle import monai_stub
    monai_stub.install_atten()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from unet.utils.atten_unet_model import AttenUNet
    cfg = getattr(OA, cfg_name)
    model = AttenUNet(**cfg).train()
    OA.randomize_(model.named_parameters(), seed=3)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == OA.param_shapes(cfg)
    assert list(model.state_dict()) == list(OA.param_shapes(cfg))                 # registration order too
    x, ctx, tgt = _atten_inputs((1, 16, 32, 24), 4, cfg["cross_attention_dim"] or 1)
    if not cfg["with_conditioning"]:
        ctx = None
    y = model(x, ctx)
    loss = (y - tgt).abs().mean()
    loss.backward()
    lo, yo, go = OA.train_step(x, ctx, tgt, {k: v.detach().clone() for k, v in model.state_dict().items()}, cfg)
    assert abs(float(loss) - float(lo)) < 1e-6 and (y.detach() - yo).abs().max().item() < 1e-5
    for k, p in model.named_parameters():
        g = p.grad if p.grad is not None else torch.zeros_like(p)
        assert torch.allclose(g, go[k], atol=1e-6, rtol=1e-4), k


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "bl_methods", "BMGAN")), reason="reference checkout not mounted")
def test_bmgan_oracle_matches_live_reference():
    import importlib
    from oracle import bmgan as OB
    from oracle import monai_stub
    monai_stub.install()
    p = os.path.join(REF, "bl_methods", "BMGAN")
    if p not in sys.path:
        sys.path.insert(0, p)
    ref = importlib.import_module("bmgan_model")
    tiny = dict(input_conv_channel=8, output_conv_channel=8, down_channels=[8, 16, 16, 16], middle_channels=[16],
                up_channels=[16, 16, 16, 16, 8])
    torch.manual_seed(2)
    rg, rd, re_ = ref.dense_unet_generator(**tiny).train(), ref.patch_discriminator().train(), ref.ResNet_encoder().train()
    og, od, oe = OB.DenseUnetGenerator(**tiny).train(), OB.PatchDiscriminatorWrapper().train(), OB.ResNetEncoder().train()
    og.load_state_dict(rg.state_dict()); od.load_state_dict(rd.state_dict()); oe.load_state_dict(re_.state_dict())   # same keys
    g = torch.Generator().manual_seed(2)
    t1, z = torch.rand(1, 1, 64, 64, 64, generator=g), torch.randn(1, 8, generator=g)
    yr, yo = rg(t1, z), og(t1, z)
    assert (yr - yo).abs().max().item() < 1e-6
    assert (rd(yr) - od(yo)).abs().max().item() < 1e-6
    vol = torch.rand(1, 1, 128, 128, 128, generator=g)
    (mr, lr_), (mo, lo) = re_(vol), oe(vol)
    assert (mr - mo).abs().max().item() < 1e-5 and (lr_ - lo).abs().max().item() < 1e-5


# ---------------------------------------------------------------------------------------------------------------------
# UnetGenerator3d's other constructor families: the oracle's norm="instance" / biased convolutions / dropout-mask arguments
# against the fixtures generated from the live reference class (tests/golden/make_golden.py families).
# ---------------------------------------------------------------------------------------------------------------------
UNET_FAMILIES = {
    "unet3d_instnorm_drop_nd6_1x64x64x64": ("instance", False, True),
    "unet3d_instaffine_nd5_2x32x32x32": ("instance", True, False),
    "unet3d_batchnorm_drop_nd6_1x64x64x64": ("batch", True, True),
}


def unet_family_state_dict(keys, nd, ngf, seed):
    """Zero tensors of the right shapes under the reference's keys (shapes follow from the key), filled by key."""
    levels = O.level_specs(1, 1, nd, ngf)
    sd = {}
    for k in keys:
        lv = max((l for l in levels if k.startswith(l.prefix)), key=lambda l: len(l.prefix))
        slot, leaf = k[len(lv.prefix):].split(".", 1)
        sl = O._slots(lv)
        if int(slot) == sl["downconv"]:
            shape = (lv.inner_nc, lv.outer_nc, 4, 4, 4) if leaf == "weight" else (lv.inner_nc,)
        elif int(slot) == sl["upconv"]:
            shape = (lv.outer_nc, lv.inner_nc * (1 if lv.innermost else 2), 3, 3, 3) if leaf == "weight" else (lv.outer_nc,)
        else:
            c = lv.inner_nc if int(slot) == sl["downnorm"] else lv.outer_nc
            shape = () if leaf == "num_batches_tracked" else (c,)
        if leaf == "num_batches_tracked":
            sd[k] = torch.zeros((), dtype=torch.long)
        else:
            sd[k] = torch.ones(shape) if leaf == "running_var" else torch.zeros(shape)
    return O.randomize_(sd, seed)


@pytest.mark.parametrize("name", sorted(UNET_FAMILIES))
def test_unet_family_oracle_matches_golden(name):
    norm, affine, drop = UNET_FAMILIES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    shape, seed, nd, ngf = tuple(int(v) for v in gold["shape"]), int(gold["seed"]), int(gold["num_downs"]), int(gold["ngf"])
    sd = unet_family_state_dict([str(k) for k in gold["keys"]], nd, ngf, seed)
    t1, pet = synth_pair(shape, seed)
    params = {k: v.requires_grad_(True) for k, v in sd.items()
              if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd); full.update(params)
    train = name == "unet3d_instaffine_nd5_2x32x32x32"
    y = O.forward(t1, full, num_downs=nd, ngf=ngf, training=train, norm=norm)
    loss = (y - pet).abs().mean()
    loss.backward()
    assert abs(loss.item() - float(gold["loss"])) < 1e-6
    assert np.abs(y.detach().numpy()[:, :, ::2, ::2, ::2] - gold["output_sample"]).max() < 1e-5
    for k, p in params.items():
        ref = float(gold["gradnorm/" + k])
        assert abs(p.grad.double().norm().item() - ref) <= 1e-4 * ref + 1e-7, k


@pytest.mark.parametrize("name", sorted(UNET_FAMILIES))
def test_unet_family_modules_have_the_reference_keys(name):
    """The drop-in containers of the non-default constructor families register the reference's parameters under the reference's
    names, in its order (the fixture stores the live class's state-dict keys); construction needs no GPU."""
    import functools
    import petsyn
    norm, affine, drop = UNET_FAMILIES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    nd, ngf, seed = int(gold["num_downs"]), int(gold["ngf"]), int(gold["seed"])
    if norm == "instance":
        layer = functools.partial(torch.nn.InstanceNorm3d, affine=True) if affine else torch.nn.InstanceNorm3d
    else:
        layer = torch.nn.BatchNorm3d
    m = petsyn.UnetGenerator3d(1, 1, num_downs=nd, ngf=ngf, norm_layer=layer, use_dropout=drop)
    assert list(m.state_dict().keys()) == [str(k) for k in gold["keys"]]
    assert not m.default_family()
    ref_sd = unet_family_state_dict([str(k) for k in gold["keys"]], nd, ngf, seed)
    m.load_state_dict(ref_sd)                                   # a reference checkpoint of that family loads unchanged
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 64, 64, 64))                        # no CPU path
    with pytest.raises(NotImplementedError):
        petsyn.UnetGenerator3d(1, 1, num_downs=nd, ngf=ngf, norm_layer=torch.nn.GroupNorm)


@pytest.mark.parametrize("cfg_name", ["TRAINING_JSON", "SMOKE_CFG", "ATTN_ONLY_CFG", "TWO_LAYER_CFG"])
def test_atten_unet_modules_have_the_reference_keys(cfg_name):
    """petsyn.AttenUNet registers, for every constructor family with a fixture, exactly the parameters the reference class does
    (names, order, shapes; the fixtures' ``wsum/<key>`` entries were written from the live class's named_parameters())."""
    import petsyn
    from oracle import atten_unet as OA
    cfg = getattr(OA, cfg_name)
    m = petsyn.AttenUNet(**cfg)
    shapes = OA.param_shapes(cfg)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in shapes.items()]
    fixture = {"TRAINING_JSON": "atten_unet_2x32x48x32", "SMOKE_CFG": "atten_unet_smoke_1x44x64x44",
               "ATTN_ONLY_CFG": "atten_unet_attnonly_1x32x48x32", "TWO_LAYER_CFG": "atten_unet_twolayer_1x32x48x32"}[cfg_name]
    gold = np.load(os.path.join(GOLD, fixture + ".npz"))
    assert sorted(k[len("wsum/"):] for k in gold.files if k.startswith("wsum/")) == sorted(shapes)
