"""Generate the BMGAN golden fixtures from the LIVE reference (authoring container only).

    python tests/golden/make_golden_bmgan.py

``bl_methods/BMGAN/bmgan_model.py`` is imported UNMODIFIED with ``oracle/monai_stub.py`` standing in for the
un-vendored ``monai`` / ``monai_diffusion`` packages (parity unpinned behind that boundary).  One generator
step of ``train_bmgan.py:141-161`` (LSGAN + 20*L1, LPIPS dropped) and the discriminator phase :183-200 are run on
seeded synthetic volumes; losses, outputs and per-parameter gradient norms are stored.
"""
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PETSYN_REFERENCE", "/root/reference")

SMALL = dict(input_conv_channel=64, output_conv_channel=64, down_channels=[64, 128, 128, 128], middle_channels=[128],
             up_channels=[128, 128, 128, 128, 64])
# "full": the reference's default constructor (247.6 M parameters) on the reference crop = BASELINE configs[2]'s generator
CASES = {"bmgan_small_2x64x96x64": (SMALL, (2, 64, 96, 64), 777, 2),
         "bmgan_full_1x96x128x96": ({}, (1, 96, 128, 96), 777, 3)}


def synth(shape, seed):
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    return (torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, d, h, w, generator=g) * 2 - 1,
            torch.randn(n, 8, generator=g))


def main():
    from oracle import bmgan as OB
    from oracle import monai_stub
    monai_stub.install()
    sys.path.insert(0, os.path.join(REF, "bl_methods", "BMGAN"))
    ref = importlib.import_module("bmgan_model")
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    for name, (cfg, shape, seed, stride) in CASES.items():
        if only and name not in only:
            continue
        torch.manual_seed(seed)
        gen = ref.dense_unet_generator(**cfg).train()
        disc = ref.patch_discriminator().train()
        t1, pet, z = synth(shape, seed)
        out = {"shape": np.array(shape), "seed": np.int64(seed)}
        for k, v in list(gen.state_dict().items()) + [("D." + k, v) for k, v in disc.state_dict().items()]:
            if v.dtype.is_floating_point:
                out["wsum/" + k] = np.float64(v.double().abs().sum().item())
        loss, adv, l1, fake = OB.generator_step(gen, disc, t1, pet, z)
        loss.backward()
        out.update(g_loss=np.float64(loss.item()), g_adv=np.float64(adv.item()), g_l1=np.float64(l1.item()),
                   fake_sample=fake.detach().numpy()[:, :, ::stride, ::stride, ::stride].copy(),
                   stride=np.int64(stride))
        for k, p in gen.named_parameters():
            out["gradnorm/" + k] = np.float64(p.grad.double().norm().item())
        dl = OB.discriminator_step(disc, fake.detach(), pet)
        out["d_loss"] = np.float64(dl.item())
        for k, p in disc.named_parameters():
            out["gradnorm/D." + k] = np.float64(p.grad.double().norm().item())
        for k, v in disc.state_dict().items():
            if "running" in k:
                out["buffer/D." + k] = v.numpy().copy()
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: g_loss={loss.item():.6f} adv={adv.item():.6f} l1={l1.item():.6f} d_loss={dl.item():.6f} "
              f"-> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
