"""Input side of the path: the output contract of ``pair_PET_T1dataset`` + ``DataLoader`` (unet/utils/dataset.py:14-139,
train_unet.py:111-127), produced on the device.

The reference loads synchronously (``num_workers=0``): NIfTI read -> MONAI pad/crop on the CPU -> ``/ max`` on the CPU ->
collate -> ``.to(device)`` inside the step.  Here the raw arrays go from pinned staging memory to the device on a copy
stream and ``petsyn_volume_prepare`` (``csrc/volume_prep.cu``) does pad -> centre crop -> ``/ max`` there, one batch ahead of
the step that consumes it.  What a batch looks like is unchanged:

    t1_img  [B, 1, 96, 128, 96] fp32      pet_img  same      info  [B, n_covariates] fp32      subject, t1_date, pet_date

Reading NIfTI files (SimpleITK) stays outside: a *source* is any sequence of
``(t1_raw, pet_raw, row)`` with fp32 ``[d, h, w]`` arrays and ``row`` the CSV line as a dict.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from ._cabi import VolumeSrc, check, lib, ptr, require_cuda


def normalise_covariates(row: Dict[str, str], need_values: Sequence[str], min_and_max: Dict[str, Sequence[float]]) -> List[float]:
    """dataset.py:127-135: min-max scaling (in Python float = float64) of the keys that have a range, the others raw."""
    infos = []
    for k in need_values:
        v = float(row[k])
        if k in min_and_max:
            v = (v - min_and_max[k][0]) / (min_and_max[k][1] - min_and_max[k][0])
        infos.append(v)
    return infos


def distributed_indices(n: int, rank: int, world_size: int, shuffle: bool = True, seed: int = 0, epoch: int = 0) -> List[int]:
    """The index list ``torch.utils.data.DistributedSampler(dataset)`` hands to rank ``rank`` (train_unet.py:116; its
    drop_last is False: the seeded permutation is padded by wrapping around to a multiple of the world size, then strided)."""
    if shuffle:
        g = torch.Generator()
        g.manual_seed(seed + epoch)
        idx = torch.randperm(n, generator=g).tolist()
    else:
        idx = list(range(n))
    total = -(-n // world_size) * world_size
    pad = total - len(idx)
    if pad > 0:
        idx += (idx * (-(-pad // len(idx))))[:pad]
    return idx[rank:total:world_size]


def volume_prepare(raw: Sequence[torch.Tensor], out: torch.Tensor, vmax: Optional[torch.Tensor] = None,
                   stream: Optional[torch.cuda.Stream] = None) -> torch.Tensor:
    """``_preprocess_img`` (dataset.py:70-105) for a batch of raw device volumes ``[d_i, h_i, w_i]`` fp32 into
    ``out [B, 1, D, H, W]`` fp32 (pad -> centre crop -> divide by the maximum of the cropped volume)."""
    require_cuda()
    n = len(raw)
    if out.dim() != 5 or out.shape[0] != n or out.shape[1] != 1:
        raise ValueError(f"out must be [{n}, 1, D, H, W], got {tuple(out.shape)}")
    if not (out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()):
        raise ValueError("out must be a contiguous fp32 CUDA tensor")
    if vmax is None:
        vmax = torch.empty(n, dtype=torch.float32, device=out.device)
    arr = (VolumeSrc * n)()
    for i, r in enumerate(raw):
        if not (r.is_cuda and r.dtype == torch.float32 and r.is_contiguous() and r.dim() == 3):
            raise ValueError("raw volumes must be contiguous fp32 CUDA tensors [d, h, w]")
        arr[i].data, arr[i].d, arr[i].h, arr[i].w = r.data_ptr(), r.shape[0], r.shape[1], r.shape[2]
    s = (stream or torch.cuda.current_stream(out.device)).cuda_stream
    for lo in range(0, n, 16):                              # the C ABI takes up to 16 volumes per call
        cnt = min(16, n - lo)
        sub = (VolumeSrc * cnt).from_buffer(arr, lo * C.sizeof(VolumeSrc))
        check(lib.petsyn_volume_prepare(sub, cnt, ptr(out[lo:]), out.shape[2], out.shape[3], out.shape[4],
                                        ptr(vmax[lo:]), s), "volume_prepare")
    return vmax


class SyntheticPairSource:
    """Seeded stand-in for the NIfTI folders: raw T1 / PET arrays of slightly varying extents around the reference's
    registered grid (160x224x160 at 1 mm resampled to 1.5 mm, preprocess/reg_to_T1.py:29,53 => ~107x149x107) and CSV-like
    covariate rows in the ranges of ``unet/config/AV45_min_and_max.pkl``."""

    NEED_VALUES = ("ABETA", "Age", "Sex", "APOE4", "PTEDUCAT")
    MIN_AND_MAX = {"ABETA": (0.0, 2000.0), "Age": (55.0, 93.84589134246576), "PTEDUCAT": (8.0, 20.0)}

    def __init__(self, length: int = 64, base_shape: Tuple[int, int, int] = (107, 149, 107), jitter: int = 6, seed: int = 777,
                 pool: int = 4):
        self.length, self.base, self.jitter, self.seed = length, base_shape, jitter, seed
        rng = np.random.default_rng(seed)
        # a small pool of distinct random volumes at the maximum extent; items are shifted windows of them (cheap to build)
        mx = tuple(b + jitter for b in base_shape)
        self._pool = [(rng.random(mx, dtype=np.float32) * 3000.0, rng.random(mx, dtype=np.float32) * 8.0)
                      for _ in range(pool)]

    def __len__(self) -> int:
        return self.length

    def __getitem__(self, i: int):
        rng = np.random.default_rng(self.seed * 1000003 + i)
        shp = tuple(int(b + rng.integers(-self.jitter, self.jitter + 1)) for b in self.base)
        a, b = self._pool[i % len(self._pool)]
        t1, pet = a[:shp[0], :shp[1], :shp[2]], b[:shp[0], :shp[1], :shp[2]]        # strided views: the loader gathers them
        row = {"Subject": f"synthetic_{i:04d}", "T1_date": "2000-01-01", "PET_date": "2000-01-02",
               "ABETA": repr(float(rng.uniform(200, 1700))), "Age": repr(float(rng.uniform(55, 93))),
               "Sex": repr(float(rng.integers(0, 2))), "APOE4": repr(float(rng.integers(0, 3))),
               "PTEDUCAT": repr(float(rng.integers(8, 21)))}
        return t1, pet, row


class PairVolumeLoader:
    """``DataLoader(pair_PET_T1dataset(...), batch_size, sampler=DistributedSampler, drop_last=True)`` (train_unet.py:111-121)
    with the preprocessing on the device and one batch of look-ahead.

    Per batch: raw arrays -> pinned staging slab (host memcpy) -> one async H2D per volume on the copy stream ->
    ``volume_prepare`` on the copy stream -> event; the consumer's stream waits on the event only.  Three staging / output
    slots rotate, so the host staging, H2D and preparation of batch k+1 overlap the training step of batch k-1 / k.
    """

    def __init__(self, source, batch_size: int, device, crop_size: Tuple[int, int, int] = (96, 128, 96),
                 need_values: Sequence[str] = (), min_and_max: Optional[Dict[str, Sequence[float]]] = None,
                 rank: int = 0, world_size: int = 1, shuffle: bool = True, seed: int = 0, drop_last: bool = True,
                 max_raw_voxels: Optional[int] = None):
        require_cuda()
        self.source, self.bs, self.dev = source, batch_size, torch.device(device)
        self.crop = tuple(crop_size)
        self.need_values, self.min_and_max = list(need_values), dict(min_and_max or {})
        self.rank, self.world, self.shuffle, self.seed, self.drop_last = rank, world_size, shuffle, seed, drop_last
        self.epoch = 0
        if max_raw_voxels is None:
            max_raw_voxels = int(np.prod([c + c // 2 for c in self.crop]))        # raw extents up to 1.5x the crop
        self.cap = max_raw_voxels
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        nslots = 3          # batch k+1 is staged while step k-1 still runs: its slot was last read by step k-2
        mk = lambda *shape: torch.empty(*shape, dtype=torch.float32, device=self.dev)
        self.slots = [{
            "pin": torch.empty(2 * batch_size, self.cap, dtype=torch.float32).pin_memory(),
            "raw": mk(2 * batch_size, self.cap),
            "t1": mk(batch_size, 1, *self.crop), "pet": mk(batch_size, 1, *self.crop),
            "info_pin": torch.empty(batch_size, max(1, len(self.need_values)), dtype=torch.float32).pin_memory(),
            "info": mk(batch_size, max(1, len(self.need_values))),
            "vmax": mk(2 * batch_size),
            "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
        } for _ in range(nslots)]
        for sl in self.slots:
            sl["pin_np"] = sl["pin"].numpy()               # shares the pinned allocation
        self.h2d_bytes = 0

    # -------------------------------------------------------------------------------------------- sampler
    def set_epoch(self, epoch: int) -> None:
        """``DistributedSampler.set_epoch`` (train_unet.py:131)."""
        self.epoch = epoch

    def _indices(self) -> List[int]:
        return distributed_indices(len(self.source), self.rank, self.world, self.shuffle, self.seed, self.epoch)

    def __len__(self) -> int:
        per_rank = -(-len(self.source) // self.world)
        return per_rank // self.bs if self.drop_last else -(-per_rank // self.bs)

    # -------------------------------------------------------------------------------------------- pipeline
    def _stage(self, slot: dict, items: List[int]) -> dict:
        """Host side of one batch: copy into pinned memory, enqueue H2D + device preparation on the copy stream."""
        slot["free"].synchronize()                       # the step that consumed this slot three batches ago is done
        b = len(items)
        meta, shapes = [], []
        for j, i in enumerate(items):
            t1, pet, row = self.source[i]
            for k, a in ((2 * j, t1), (2 * j + 1, pet)):
                a = np.asarray(a)
                if a.ndim != 3 or a.size > self.cap:
                    raise ValueError(f"raw volume {a.shape} does not fit the staging slab ({self.cap} voxels)")
                # one pass: gathers a strided view and converts the dtype straight into pinned memory
                np.copyto(slot["pin_np"][k, :a.size].reshape(a.shape), a, casting="unsafe")
                shapes.append(a.shape)
            cov = normalise_covariates(row, self.need_values, self.min_and_max)
            if cov:
                slot["info_pin"][j, :len(cov)] = torch.tensor(cov, dtype=torch.float)
            meta.append((row.get("Subject"), row.get("T1_date"), row.get("PET_date")))
        with torch.cuda.stream(self.copy_stream):
            raws = []
            for k, shp in enumerate(shapes):
                nvox = int(np.prod(shp))
                slot["raw"][k, :nvox].copy_(slot["pin"][k, :nvox], non_blocking=True)
                raws.append(slot["raw"][k, :nvox].view(*shp))
                self.h2d_bytes += 4 * nvox
            slot["info"].copy_(slot["info_pin"], non_blocking=True)
            volume_prepare(raws[0::2], slot["t1"][:b], slot["vmax"][:b], self.copy_stream)
            volume_prepare(raws[1::2], slot["pet"][:b], slot["vmax"][b:2 * b], self.copy_stream)
            slot["ready"].record(self.copy_stream)
        return {"slot": slot, "b": b, "meta": meta}

    def __iter__(self) -> Iterator[tuple]:
        idx = self._indices()
        batches = [idx[i:i + self.bs] for i in range(0, len(idx), self.bs)]
        if self.drop_last and batches and len(batches[-1]) < self.bs:
            batches.pop()
        # a previous (possibly abandoned) iteration may have left consumer work in flight on the slots: order this epoch's
        # copies and preparation kernels after everything the consumer has enqueued so far
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        pending = self._stage(self.slots[0], batches[0]) if batches else None
        for k in range(len(batches)):
            cur = pending
            pending = self._stage(self.slots[(k + 1) % len(self.slots)], batches[k + 1]) if k + 1 < len(batches) else None
            slot, b = cur["slot"], cur["b"]
            consumer = torch.cuda.current_stream(self.dev)
            consumer.wait_event(slot["ready"])
            info = slot["info"][:b, :len(self.need_values)] if self.need_values else []
            subjects, t1_dates, pet_dates = (list(x) for x in zip(*cur["meta"]))
            yield slot["t1"][:b], slot["pet"][:b], info, subjects, t1_dates, pet_dates
            slot["free"].record(torch.cuda.current_stream(self.dev))     # everything the consumer enqueued so far
