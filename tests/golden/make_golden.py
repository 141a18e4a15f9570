"""Generate the committed golden fixtures from the LIVE reference.

Run in the authoring container only (needs /root/reference, which is not on the GPU box):

    python tests/golden/make_golden.py

For each case the unmodified reference module (``unet/utils/unet_model.py``) is constructed
under ``torch.manual_seed(seed)``, run forward + L1 loss + backward on seeded synthetic
volumes, and the results stored as small ``.npz`` files next to this script:
output volume (or a strided sample of it), loss, per-parameter gradient norms, updated
BatchNorm running statistics and a per-tensor weight checksum (to detect RNG drift between
torch builds -- the weights themselves are re-drawn from the seed, never stored).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PETSYN_REFERENCE", "/root/reference")

CASES = {
    # name: (ngf, (N, D, H, W), seed, store_full_output)
    "unet3d_ngf8_1x32x48x32": (8, (1, 32, 48, 32), 777, True),
    "unet3d_ngf32_2x32x48x32": (32, (2, 32, 48, 32), 777, True),
    "unet3d_ngf64_1x32x32x48": (64, (1, 32, 32, 48), 777, True),
    # BASELINE config 1: full width, 96x112x96, batch 1 -- scalars + a strided sample
    "unet3d_ngf64_1x96x112x96": (64, (1, 96, 112, 96), 777, False),
}


def synth_pair(shape, seed):
    """t1 and pet ~ U[0,1) drawn consecutively from one seeded generator (SURVEY 8d cfg 1)."""
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    t1 = torch.rand(n, 1, d, h, w, generator=g)
    pet = torch.rand(n, 1, d, h, w, generator=g)
    return t1, pet


def main():
    sys.path.insert(0, REF)
    from unet.utils.unet_model import UnetGenerator3d  # the reference, unmodified

    torch.set_num_threads(os.cpu_count() or 1)
    for name, (ngf, shape, seed, full) in CASES.items():
        torch.manual_seed(seed)
        model = UnetGenerator3d(1, 1, num_downs=4, ngf=ngf)
        model.train()
        t1, pet = synth_pair(shape, seed)
        wsum = {k: float(v.double().abs().sum()) for k, v in model.state_dict().items()
                if v.dtype.is_floating_point}
        y = model(t1.clone())
        loss = torch.nn.L1Loss()(y, pet)
        loss.backward()
        out = {
            "loss": np.float64(loss.item()),
            "grad_norm_total": np.float64(
                torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())).item()),
            "shape": np.array(shape), "ngf": np.int64(ngf), "seed": np.int64(seed),
        }
        for k, p in model.named_parameters():
            out["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        for k, v in model.state_dict().items():
            if k.endswith("running_mean") or k.endswith("running_var"):
                out["buffer/" + k] = v.numpy().copy()
        for k, v in wsum.items():
            out["wsum/" + k] = np.float64(v)
        yd = y.detach().numpy()
        out["output" if full else "output_sample"] = yd if full else yd[:, :, ::8, ::8, ::8].copy()
        # eval-mode forward with the (updated) running statistics: the inference path
        model.eval()
        with torch.no_grad():
            ye = model(t1.clone()).numpy()
        out["output_eval" if full else "output_eval_sample"] = ye if full else ye[:, :, ::8, ::8, ::8].copy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={out['loss']:.6f} gradnorm={out['grad_norm_total']:.5f} -> {os.path.getsize(path)/1e3:.0f} kB")


# Other constructor families (unet_model.py:42-45 biased convolutions under InstanceNorm3d, :17 / :87-88 Dropout(0.5) with more
# than five levels).  Parameters are drawn by key (oracle.unet3d.randomize_), so the fixture does not depend on constructor
# RNG order; eval() mode where a Dropout layer exists (its mask is not reproducible across implementations; InstanceNorm3d
# without running statistics behaves the same in both modes).
FAMILIES = {
    # name: (norm, affine, use_dropout, num_downs, ngf, shape, seed, train_mode)
    "unet3d_instnorm_drop_nd6_1x64x64x64": ("instance", False, True, 6, 8, (1, 64, 64, 64), 31, False),
    "unet3d_instaffine_nd5_2x32x32x32": ("instance", True, False, 5, 8, (2, 32, 32, 32), 32, True),
    "unet3d_batchnorm_drop_nd6_1x64x64x64": ("batch", True, True, 6, 8, (1, 64, 64, 64), 33, False),
}


def families():
    import functools
    sys.path.insert(0, REF)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from unet.utils.unet_model import UnetGenerator3d  # the reference, unmodified
    from oracle import unet3d as O
    torch.set_num_threads(os.cpu_count() or 1)
    for name, (norm, affine, drop, nd, ngf, shape, seed, train) in FAMILIES.items():
        if norm == "instance":
            layer = functools.partial(torch.nn.InstanceNorm3d, affine=True) if affine else torch.nn.InstanceNorm3d
        else:
            layer = torch.nn.BatchNorm3d
        model = UnetGenerator3d(1, 1, num_downs=nd, ngf=ngf, norm_layer=layer, use_dropout=drop)
        O.randomize_(model.state_dict(), seed)
        model.train(train)
        t1, pet = synth_pair(shape, seed)
        y = model(t1.clone())
        loss = torch.nn.L1Loss()(y, pet)
        loss.backward()
        out = {"loss": np.float64(loss.item()), "shape": np.array(shape), "ngf": np.int64(ngf), "seed": np.int64(seed),
               "num_downs": np.int64(nd), "keys": np.array(list(model.state_dict().keys()))}
        for k, p in model.named_parameters():
            out["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        out["output_sample"] = y.detach().numpy()[:, :, ::2, ::2, ::2].copy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={out['loss']:.6f} -> {os.path.getsize(path)/1e3:.0f} kB")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "families":
        families()
    else:
        main()
