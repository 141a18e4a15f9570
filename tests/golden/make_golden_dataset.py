"""Generate the data-format golden fixture from the LIVE reference (authoring container only).

    python tests/golden/make_golden_dataset.py

``unet/utils/dataset.py`` is imported UNMODIFIED over ``oracle/monai_stub.install_dataset()`` (MONAI's SpatialPad /
CenterSpatialCrop restated from upstream; SimpleITK is only touched by the NIfTI read, which is outside the path) and
``pair_PET_T1dataset._preprocess_img`` plus the covariate normalisation of ``__getitem__`` are run on small seeded raw
volumes whose extents exercise pad-only, crop-only and mixed axes, odd and even differences.
"""
import os
import pickle
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PETSYN_REFERENCE", "/root/reference")

CROP = (12, 16, 10)
RAW_SHAPES = [(12, 16, 10), (15, 21, 13), (9, 11, 7), (14, 13, 10), (13, 16, 17), (11, 22, 9)]


def main():
    from oracle import monai_stub
    monai_stub.install_dataset()
    sys.path.insert(0, REF)
    from unet.utils.dataset import pair_PET_T1dataset

    csv_path = os.path.join(REF, "unet", "config", "pair_t1_AV45_training_with_csf.csv")
    need = ['ABETA', 'Age', 'Sex', 'APOE4', 'PTEDUCAT']                     # train_unet.py:62 (AV45)
    mm = pickle.load(open(os.path.join(REF, "unet", "config", "AV45_min_and_max.pkl"), "rb"))
    ds = pair_PET_T1dataset(info_csv=csv_path, crop=True, crop_size=CROP, PET_dir="/nonexistent", T1_dir="/nonexistent",
                            min_and_max=mm, need_values=need, return_MRI=False)
    assert len(ds) == 0            # no image folders here: rows are skipped by the existence check (:55-56)
    out = {"crop": np.array(CROP)}
    rng = np.random.default_rng(777)
    for i, shp in enumerate(RAW_SHAPES):
        a = rng.random(shp, dtype=np.float32) * 3000.0
        b = (rng.random(shp, dtype=np.float32) - 0.25) * 7.0             # PET-like range with some negatives
        t1, pet = ds._preprocess_img(a, b)
        out[f"raw_t1_{i}"], out[f"raw_pet_{i}"] = a, b
        out[f"t1_{i}"], out[f"pet_{i}"] = t1.numpy(), pet.numpy()
    # covariates: feed the first CSV rows through __getitem__ (return_MRI=False skips the NIfTI read)
    import csv
    rows = []
    with open(csv_path, "r", encoding="utf-8") as f:
        for k, row in enumerate(csv.DictReader(f)):
            if k >= 8:
                break
            rows.append(row)
    ds.lines = [dict(row, T1_ImagePath="", PET_ImagePath="") for row in rows]
    infos = np.stack([ds[k][2].numpy() for k in range(len(rows))])
    out["covariates_raw"] = np.array([[float(r[k]) for k in need] for r in rows], dtype=np.float64)
    out["covariates"] = infos
    out["min_and_max"] = np.array([[mm[k][0], mm[k][1]] if k in mm else [np.nan, np.nan] for k in need], dtype=np.float64)
    path = os.path.join(HERE, "dataset_crop12x16x10.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
