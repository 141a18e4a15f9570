"""GPU parity of the input data format: ``petsyn_volume_prepare`` and ``PairVolumeLoader`` against the
numpy restatement of ``pair_PET_T1dataset`` (unet/utils/dataset.py:70-139) -- BIT-EXACT (a gather and one IEEE fp32 division
per voxel) -- and the reference checkpoint dictionary (train_unet.py:85-109,297-302) round trip."""
import io
import os

import numpy as np
import pytest
import torch

from oracle import atten_unet as OA
from oracle import dataset as OD

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "dataset_crop12x16x10.npz")


def _prepare(petsyn, raws, crop):
    dev = [torch.from_numpy(r).cuda() for r in raws]
    out = torch.full((len(raws), 1) + tuple(crop), 7.0, dtype=torch.float32, device="cuda")
    vmax = petsyn.volume_prepare(dev, out)
    torch.cuda.synchronize()
    return out.cpu().numpy(), vmax.cpu().numpy()


def test_volume_prepare_matches_reference_fixture_bit_for_bit(petsyn):
    g = np.load(GOLD)
    crop = tuple(int(c) for c in g["crop"])
    n = len([k for k in g.files if k.startswith("raw_t1_")])
    raws = [g[f"raw_t1_{i}"] for i in range(n)] + [g[f"raw_pet_{i}"] for i in range(n)]
    want = [g[f"t1_{i}"] for i in range(n)] + [g[f"pet_{i}"] for i in range(n)]
    got, _ = _prepare(petsyn, raws, crop)              # one ragged batch: every volume has its own extent
    for i, w in enumerate(want):
        assert np.array_equal(got[i], w), i


@pytest.mark.parametrize("crop", [(96, 128, 96), (160, 192, 160)])
def test_volume_prepare_full_size_ragged_batch(petsyn, crop):
    """BASELINE sizes: crop-only, pad-only and mixed axes (odd and even differences) in one batch; > 16 volumes (two C calls)."""
    rng = np.random.default_rng(1)
    d, h, w = crop
    shapes = [(d, h, w), (d + 11, h + 21, w + 11), (d - 5, h - 8, w - 3), (d + 6, h - 7, w + 1), (d - 1, h + 2, w - 10)]
    if crop == (96, 128, 96):
        shapes = shapes + [(107, 149, 107)] * 13          # 18 volumes
    raws = [(rng.random(s, dtype=np.float32) - 0.1) * 2500.0 for s in shapes]
    got, vmax = _prepare(petsyn, raws, crop)
    for i, r in enumerate(raws):
        c = np.ascontiguousarray(OD.pad_center_crop(r, crop))
        assert vmax[i] == c.max()
        assert np.array_equal(got[i, 0], c / np.float32(c.max())), shapes[i]
        assert got[i].max() == 1.0


def test_volume_prepare_edge_cases(petsyn):
    crop = (8, 8, 8)
    neg = -np.abs(np.random.default_rng(2).random((12, 9, 10), dtype=np.float32)) - 1.0      # all negative, crop only
    neg_padded = neg[:6, :6, :6].copy()                                                      # all negative + zero padding
    one = np.full((1, 1, 1), 5.0, np.float32)                                                # a single voxel, padded all round
    zero = np.zeros((8, 8, 8), np.float32)                                                   # 0 / 0 = NaN, as torch
    got, vmax = _prepare(petsyn, [neg, neg_padded, one, zero], crop)
    for i, r in enumerate([neg, neg_padded, one]):
        c = np.ascontiguousarray(OD.pad_center_crop(r, crop))
        with np.errstate(divide="ignore", invalid="ignore"):
            want = c / np.float32(c.max())
        assert vmax[i] == c.max()
        assert np.array_equal(got[i, 0], want, equal_nan=True), i
        ok = ~np.isnan(want)                                  # -0.0 where padding is divided by a negative maximum; NaN signs are free
        assert np.array_equal(np.signbit(got[i, 0])[ok], np.signbit(want)[ok]), i
    assert vmax[1] == 0.0 and vmax[2] == 5.0 and got[2].sum() == 1.0
    assert np.isnan(got[3]).all()
    with pytest.raises(ValueError):
        petsyn.volume_prepare([torch.zeros(4, 4, 4, device="cuda", dtype=torch.float64)],
                              torch.zeros(1, 1, 8, 8, 8, device="cuda"))
    with pytest.raises(ValueError):
        petsyn.volume_prepare([torch.zeros(4, 4, 4, device="cuda")], torch.zeros(2, 1, 8, 8, 8, device="cuda"))


def test_loader_batches_follow_the_dataset_contract(petsyn):
    """Every batch of PairVolumeLoader == collate of pair_PET_T1dataset.__getitem__ over DistributedSampler's indices."""
    from petsyn_b200.data import distributed_indices
    src = petsyn.SyntheticPairSource(length=11, base_shape=(30, 40, 28), jitter=5, seed=9)
    crop, bs = (32, 32, 32), 2
    for world, rank in ((1, 0), (2, 1)):
        ld = petsyn.PairVolumeLoader(src, bs, "cuda", crop_size=crop, need_values=src.NEED_VALUES,
                                     min_and_max=src.MIN_AND_MAX, rank=rank, world_size=world, seed=4)
        ld.set_epoch(2)
        idx = distributed_indices(len(src), rank, world, True, 4, 2)
        nb = 0
        burn = torch.empty(1 << 22, device="cuda")
        for k, (t1, pet, info, subj, d1, d2) in enumerate(ld):
            burn.normal_()                                     # consumer-side work between batches
            assert t1.shape == (bs, 1) + crop and pet.shape == t1.shape and info.shape == (bs, 5)
            t1c, petc, infoc = t1.cpu().numpy(), pet.cpu().numpy(), info.cpu().numpy()
            for j in range(bs):
                a, b, row = src[idx[k * bs + j]]
                o1, o2 = OD.preprocess_pair(a, b, crop)
                assert np.array_equal(t1c[j], o1) and np.array_equal(petc[j], o2)
                assert np.array_equal(infoc[j], OD.normalise_covariates(row, src.NEED_VALUES, src.MIN_AND_MAX))
                assert subj[j] == row["Subject"] and d1[j] == row["T1_date"] and d2[j] == row["PET_date"]
            nb += 1
        assert nb == len(ld) == len(idx) // bs                # drop_last=True (train_unet.py:120)
        assert ld.h2d_bytes > 0


def test_checkpoint_round_trip_in_the_reference_format(petsyn):
    from petsyn_b200.train import AttenUNetTrainer
    shape, lr = (2, 32, 48, 32), 5e-4
    g = torch.Generator().manual_seed(1)
    batches = [(torch.rand(2, 1, 32, 48, 32, generator=g).cuda(), torch.rand(2, 1, 5, generator=g).cuda(),
                torch.rand(2, 1, 32, 48, 32, generator=g).cuda()) for _ in range(3)]

    def fresh(seed):
        m = petsyn.AttenUNet(**OA.TRAINING_JSON)
        OA.randomize_(m.named_parameters(), seed=seed)
        m = m.cuda().train()
        return m, AttenUNetTrainer(m, lr=lr, example_input=batches[0][0])

    m1, t1 = fresh(1)
    for b in batches[:2]:
        t1.step(*b)
    buf = io.BytesIO()
    torch.save(t1.checkpoint(epoch=6, eval_loss=0.125), buf)
    ck = torch.load(io.BytesIO(buf.getvalue()), map_location="cuda", weights_only=False)
    assert set(ck) == {"unet", "epoch", "g_optimizer", "eval_loss"} and ck["epoch"] == 6       # train_unet.py:297-301
    assert list(ck["unet"]) == list(m1.state_dict())
    # torch.optim.Adam over the module's parameters accepts 'g_optimizer' as is (train_unet.py:109)
    opt = torch.optim.Adam(m1.parameters(), lr=1.0)
    opt.load_state_dict(ck["g_optimizer"])
    assert opt.param_groups[0]["lr"] == lr and len(opt.state) == len(list(m1.parameters()))
    assert all(float(s["step"]) == 2.0 for s in opt.state.values())
    # a DDP-saved checkpoint carries the 'module.' prefix (train_unet.py:72-74): written on request, accepted on load
    ck_ddp = t1.checkpoint(epoch=6, eval_loss=0.125, ddp_prefix=True)
    assert list(ck_ddp["unet"]) == ["module." + k for k in m1.state_dict()]
    ck["unet"] = ck_ddp["unet"]
    m2, t2 = fresh(2)
    assert t2.load_checkpoint(ck) == 7
    for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(a, b), k
    assert torch.equal(t2.m, t1.m) and torch.equal(t2.v, t1.v) and int(t2.step_dev.item()) == 2
    la, lb = t1.step(*batches[2]).item(), t2.step(*batches[2]).item()
    assert abs(la - lb) < 1e-3
    for (k, a), b in zip(m1.state_dict().items(), m2.state_dict().values()):
        assert (a - b).abs().max().item() <= 2.5 * lr, k      # same trajectory up to reduction-order noise under Adam
