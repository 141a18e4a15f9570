"""Generate the AttenUNet golden fixture from the LIVE reference (authoring container only).

    python tests/golden/make_golden_atten.py

``unet/utils/atten_unet_model.py`` is imported UNMODIFIED over ``oracle/monai_stub.install_atten()``; the model is built
from ``unet/config/training.json``'s ``atten_unet_def`` with ``cross_attention_dim = 5`` (train_unet.py:64-68); every
parameter is re-drawn by name (``oracle.atten_unet.randomize_``) because the reference's zero-initialised modules make
the default-initialised network output exactly 0.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PETSYN_REFERENCE", "/root/reference")


def synth(shape, seed, cdim=5):
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    return (torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, cdim, generator=g),
            torch.rand(n, 1, d, h, w, generator=g))


def main():
    from oracle import atten_unet as OA
    from oracle import monai_stub
    monai_stub.install_atten()
    sys.path.insert(0, REF)
    from unet.utils.atten_unet_model import AttenUNet
    cfg = json.load(open(os.path.join(REF, "unet", "config", "training.json")))["atten_unet_def"]
    cfg["cross_attention_dim"] = 5
    assert cfg == OA.TRAINING_JSON, (cfg, OA.TRAINING_JSON)
    torch.set_num_threads(os.cpu_count() or 1)
    cases = (("atten_unet_2x32x48x32", (2, 32, 48, 32), 777, 1, cfg),
             # BASELINE configs[1] at full size (the bench default): output stored on a stride-3 lattice, small gradients whole
             ("atten_unet_2x96x128x96", (2, 96, 128, 96), 777, 3, cfg),
             # the reference's own smoke configuration (atten_unet_model.py:2034-2051: conv-form resampling, 8-channel heads,
             # 2-D context) on a volume whose coarsest grid is odd (11 x 16 x 11)
             ("atten_unet_smoke_1x44x64x44", (1, 44, 64, 44), 777, 2, OA.SMOKE_CFG),
             # with_conditioning=False: Attn{Down,Mid,Up}Block / AttentionBlock instead of the SpatialTransformer, no context
             ("atten_unet_attnonly_1x32x48x32", (1, 32, 48, 32), 777, 1, OA.ATTN_ONLY_CFG),
             # transformer_num_layers = 2: two BasicTransformerBlocks per SpatialTransformer
             ("atten_unet_twolayer_1x32x48x32", (1, 32, 48, 32), 777, 1, OA.TWO_LAYER_CFG))
    only = sys.argv[1:]
    for name, shape, seed, stride, case_cfg in cases:
        if only and name not in only:
            continue
        model = AttenUNet(**case_cfg).train()
        OA.randomize_(model.named_parameters(), seed=seed)
        x, ctx, tgt = synth(shape, seed, case_cfg["cross_attention_dim"] or 1)
        if case_cfg is OA.SMOKE_CFG:
            ctx = ctx[:, 0]                       # [N, C]: the x.dim() < 3 -> unsqueeze branch (:110-112)
        if not case_cfg["with_conditioning"]:
            ctx = None
        y = model(x, ctx)
        loss = torch.nn.L1Loss()(y, tgt)
        loss.backward()
        out = {"shape": np.array(shape), "seed": np.int64(seed), "loss": np.float64(loss.item()),
               "stride": np.int64(stride), "output": y.detach().numpy()[:, :, ::stride, ::stride, ::stride].copy(),
               "output_absmax": np.float64(y.detach().abs().max().item())}
        tot = 0.0
        for k, p in model.named_parameters():
            gnorm = 0.0 if p.grad is None else p.grad.double().norm().item()
            out["gradnorm/" + k] = np.float64(gnorm)
            out["wsum/" + k] = np.float64(p.detach().double().abs().sum().item())
            if stride > 1 and p.grad is not None and p.numel() <= 8192:
                out["grad/" + k] = p.grad.numpy().copy()
            tot += gnorm ** 2
        out["grad_norm_total"] = np.float64(tot ** 0.5)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: loss={loss.item():.6f} gradnorm={tot ** 0.5:.5f} -> {os.path.getsize(path) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
