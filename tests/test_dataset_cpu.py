"""CPU tests of the input data format: the oracle restatement of pair_PET_T1dataset against the fixture
generated from the reference class itself, the C ABI's window arithmetic, the sampler, and the checkpoint dictionary."""
import os

import numpy as np
import pytest
import torch

from oracle import dataset as OD

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dataset_crop12x16x10.npz")
NEED = ['ABETA', 'Age', 'Sex', 'APOE4', 'PTEDUCAT']


def test_oracle_matches_reference_fixture_bit_for_bit():
    g = np.load(GOLD)
    crop = tuple(int(c) for c in g["crop"])
    n = len([k for k in g.files if k.startswith("raw_t1_")])
    assert n >= 6
    for i in range(n):
        t1, pet = OD.preprocess_pair(g[f"raw_t1_{i}"], g[f"raw_pet_{i}"], crop)
        assert t1.shape == (1,) + crop and t1.dtype == np.float32
        assert np.array_equal(t1, g[f"t1_{i}"]) and np.array_equal(pet, g[f"pet_{i}"])
        assert t1.max() == 1.0 and pet.max() == 1.0
    mm = {k: tuple(r) for k, r in zip(NEED, g["min_and_max"]) if not np.isnan(r[0])}
    for raw, want in zip(g["covariates_raw"], g["covariates"]):
        got = OD.normalise_covariates({k: repr(float(v)) for k, v in zip(NEED, raw)}, NEED, mm)
        assert np.array_equal(got, want)


@pytest.mark.reference
def test_oracle_matches_live_reference():
    ref = os.environ.get("PETSYN_REFERENCE", "/root/reference")
    if not os.path.isdir(ref):
        pytest.skip("reference checkout not mounted")
    import sys
    from oracle import monai_stub
    monai_stub.install_dataset()
    sys.path.insert(0, ref)
    try:
        from unet.utils.dataset import pair_PET_T1dataset
    finally:
        sys.path.remove(ref)
    ds = pair_PET_T1dataset(info_csv=os.path.join(ref, "unet/config/pair_t1_AV45_test_with_csf.csv"), crop=True,
                            crop_size=(10, 12, 8), PET_dir="/nonexistent", T1_dir="/nonexistent", return_MRI=False)
    rng = np.random.default_rng(5)
    for shp in [(10, 12, 8), (7, 15, 8), (13, 9, 11), (11, 13, 5)]:
        a, b = rng.random(shp, dtype=np.float32) * 100, rng.random(shp, dtype=np.float32) - 0.5
        t1, pet = ds._preprocess_img(a, b)
        o1, o2 = OD.preprocess_pair(a, b, (10, 12, 8))
        assert np.array_equal(t1.numpy(), o1) and np.array_equal(pet.numpy(), o2)


def test_window_offset_matches_pad_then_crop(petsyn):
    from petsyn_b200 import _cabi
    for roi in (1, 2, 7, 8, 96, 128):
        for raw in list(range(1, 40)) + [90, 96, 107, 128, 149, 224]:
            off = _cabi.lib.petsyn_volume_window_offset(raw, roi)
            assert off == OD.window_offset(raw, roi)
            line = np.arange(1, raw + 1, dtype=np.float32)
            want = OD.pad_center_crop(line[:, None, None], (roi, 1, 1))[:, 0, 0]
            got = np.array([line[o + off] if 0 <= o + off < raw else 0.0 for o in range(roi)], dtype=np.float32)
            assert np.array_equal(got, want), (raw, roi)


def test_volume_prepare_validates_before_any_device_work(petsyn):
    import ctypes as C
    from petsyn_b200 import _cabi
    src = (_cabi.VolumeSrc * 1)()
    rc = _cabi.lib.petsyn_volume_prepare(src, 17, 1, 8, 8, 8, 1, None)
    assert rc == _cabi.E_INVAL and "volumes per call" in _cabi.last_error()
    rc = _cabi.lib.petsyn_volume_prepare(src, 1, 1, 8, 8, 8, 1, None)
    assert rc == _cabi.E_INVAL and "empty" in _cabi.last_error()
    assert C.sizeof(_cabi.VolumeSrc) == 24


def test_sampler_is_torch_distributed_sampler(petsyn):
    from petsyn_b200.data import distributed_indices
    from torch.utils.data import DistributedSampler

    class D:
        def __len__(self):
            return 37
    for world in (1, 2, 3, 8):
        for rank in range(world):
            for shuffle in (True, False):
                s = DistributedSampler(D(), num_replicas=world, rank=rank, shuffle=shuffle, seed=5)
                s.set_epoch(3)
                assert list(s) == distributed_indices(37, rank, world, shuffle, 5, 3)


def test_synthetic_source_is_seeded_and_ragged(petsyn):
    a, b = petsyn.SyntheticPairSource(length=8, seed=3), petsyn.SyntheticPairSource(length=8, seed=3)
    shapes = set()
    for i in range(8):
        t1a, peta, rowa = a[i]
        t1b, petb, rowb = b[i]
        assert np.array_equal(t1a, t1b) and np.array_equal(peta, petb) and rowa == rowb
        shapes.add(t1a.shape)
        assert t1a.dtype == np.float32 and t1a.ndim == 3
    assert len(shapes) > 1
    from petsyn_b200.data import normalise_covariates
    cov = normalise_covariates(rowa, a.NEED_VALUES, a.MIN_AND_MAX)
    assert len(cov) == 5 and 0.0 <= cov[0] <= 1.0 and 0.0 <= cov[1] <= 1.0


def test_adam_state_dict_round_trips_through_torch_adam(petsyn):
    """The 'g_optimizer' entry of the reference checkpoint (train_unet.py:109,297-302) is torch.optim.Adam's state dict."""
    from petsyn_b200.train import adam_state_dict, read_adam_state_dict
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    ref = torch.optim.Adam(ps, lr=5e-4)
    for _ in range(3):
        for p in ps:
            p.grad = torch.randn_like(p)
        ref.step()
    want = ref.state_dict()
    m = [want["state"][i]["exp_avg"].clone() for i in range(2)]
    v = [want["state"][i]["exp_avg_sq"].clone() for i in range(2)]
    ours = adam_state_dict(ps, m, v, 3, 5e-4, (0.9, 0.999), 1e-8)
    assert ours["param_groups"] == want["param_groups"]
    assert set(ours["state"]) == set(want["state"])
    for i in range(2):
        assert set(ours["state"][i]) == set(want["state"][i])
        for k in ("step", "exp_avg", "exp_avg_sq"):
            assert torch.equal(ours["state"][i][k], want["state"][i][k])
    # torch accepts it, and continues exactly like the optimizer it was taken from
    ps2 = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    other = torch.optim.Adam(ps2, lr=1.0)
    import copy
    other.load_state_dict(copy.deepcopy(ours))       # torch aliases the 'step' tensors it is given
    for p, q in zip(ps, ps2):
        g = torch.randn_like(p)
        p.grad, q.grad = g, g.clone()
    ref.step(); other.step()
    for p, q in zip(ps, ps2):
        assert torch.equal(p, q)
    # and back
    m2, v2 = [torch.zeros_like(x) for x in m], [torch.zeros_like(x) for x in v]
    assert read_adam_state_dict(ours, m2, v2) == 3
    assert all(torch.equal(a, b) for a, b in zip(m + v, m2 + v2))
    assert adam_state_dict(ps, m, v, 0, 5e-4, (0.9, 0.999), 1e-8)["state"] == {}
    with pytest.raises(ValueError):
        read_adam_state_dict(ours, m2[:1], v2[:1])
