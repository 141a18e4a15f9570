"""CPU restatement of the input data format of the path: ``pair_PET_T1dataset`` (unet/utils/dataset.py:14-139).

TEST INFRASTRUCTURE ONLY.  numpy throughout (address arithmetic + one IEEE fp32 division per voxel: the bar is bit-exact).

``_preprocess_img`` (:70-105) applies MONAI ``SpatialPad`` and ``CenterSpatialCrop`` -- MONAI is not part of the checkout
and no version is pinned, so the two transforms follow *published upstream* MONAI 1.x:
  SpatialPad(method="symmetric")   width = max(roi - size, 0); pad (width // 2, width - width // 2) with zeros
  CenterSpatialCrop                 start = max(size // 2 - roi // 2, 0); keep [start, start + roi)
**Parity unpinned** behind that boundary; the composition (pad -> crop -> divide each image by ITS OWN maximum taken AFTER
the crop -> covariate min-max) is pinned against the reference class itself imported over ``oracle.monai_stub``
(``tests/golden/make_golden_dataset.py`` -> ``tests/golden/dataset_*.npz``).
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np


def window_offset(raw: int, roi: int) -> int:
    """Output voxel ``o`` of pad->crop reads raw voxel ``o + window_offset`` (outside the raw volume: 0)."""
    width = max(roi - raw, 0)
    before = width // 2
    start = max((raw + width) // 2 - roi // 2, 0)
    return start - before


def pad_center_crop(img: np.ndarray, roi: Sequence[int]) -> np.ndarray:
    """dataset.py:81-85 on one [d, h, w] volume."""
    pads = []
    for s, r in zip(img.shape, roi):
        width = max(r - s, 0)
        pads.append((width // 2, width - width // 2))
    img = np.pad(img, pads, mode="constant", constant_values=0)
    sl = []
    for s, r in zip(img.shape, roi):
        start = max(s // 2 - r // 2, 0)
        sl.append(slice(start, start + r))
    return img[tuple(sl)]


def preprocess_pair(img1: np.ndarray, img2: np.ndarray, crop_size=(96, 128, 96)) -> Tuple[np.ndarray, np.ndarray]:
    """``_preprocess_img`` with crop=True, random_crop=False, resize=False (the configuration every script uses,
    train_unet.py:111-114): returns two [1, d, h, w] fp32 arrays."""
    out = []
    for img in (img1, img2):
        c = np.ascontiguousarray(pad_center_crop(np.asarray(img, dtype=np.float32), crop_size))
        out.append((c / np.float32(c.max()))[None])          # :97-101  img / torch.max(img)
    return out[0], out[1]


def normalise_covariates(line: Dict[str, str], need_values: Sequence[str], min_and_max: Dict[str, Sequence[float]]) -> np.ndarray:
    """dataset.py:127-137: float64 min-max per listed key, then ``torch.tensor(infos, dtype=torch.float)``."""
    infos = []
    for k in need_values:
        v = float(line[k])
        if k in min_and_max:
            v = (v - min_and_max[k][0]) / (min_and_max[k][1] - min_and_max[k][0])
        infos.append(v)
    return np.asarray(infos, dtype=np.float64).astype(np.float32)
