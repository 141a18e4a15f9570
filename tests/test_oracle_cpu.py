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
    assert abs(float(loss) - float(gold["loss"])) < 1e-6
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
