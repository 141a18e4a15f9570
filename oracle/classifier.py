"""CPU restatement of the pMCI/sMCI classifier used for synthesize -> classify (BASELINE configs[4]).

**Parity unpinned.**  The reference scripts import ``DiffusionModelEncoder`` from the authors' un-vendored
fork (``pet_for_classification/train_atten_encoder_MCI.py:87,169``); the vendored copy
(``unet/utils/atten_unet_model.py:1863-2032``) cannot run.  This follows what the vendored source defines -- conv_in, one
down block per level built from the vendored ResnetBlock / SpatialTransformer (restated in ``oracle/atten_unet.py``), EVERY
level followed by a down-sampling ResnetBlock (:1966), flatten in NCDHW order, Linear -> ReLU -> Dropout -> Linear (:1987)
-- with the time embedding left out (the vendored ResnetBlock has no such input; the scripts pass zeros).
Test infrastructure only.
"""
from typing import Dict

import torch
import torch.nn.functional as F

from . import atten_unet as OA

TRAINING_ATTEN_JSON = dict(  # pet_for_classification/config/training_atten.json + cross_attention_dim (train_...MCI.py:85-86)
    spatial_dims=3, in_channels=1, out_channels=2, num_channels=[16, 32, 64, 128, 128], num_res_blocks=2,
    attention_levels=[False, False, False, True, True], norm_num_groups=16, norm_eps=1e-6, resblock_updown=True,
    num_head_channels=[0, 0, 0, 32, 32], with_conditioning=True, transformer_num_layers=1, upcast_attention=False,
    cross_attention_dim=5)


def forward(x: torch.Tensor, context: torch.Tensor, sd: Dict[str, torch.Tensor], cfg=TRAINING_ATTEN_JSON) -> torch.Tensor:
    """eval-mode forward: logits [N, out_channels]."""
    ch = cfg["num_channels"]
    nres = cfg["num_res_blocks"]
    nres = [nres] * len(ch) if isinstance(nres, int) else list(nres)
    att, hc = cfg["attention_levels"], cfg["num_head_channels"]
    hc = [hc] * len(ch) if isinstance(hc, int) else list(hc)
    groups, eps = cfg["norm_num_groups"], cfg["norm_eps"]
    if context.dim() < 3:
        context = context.unsqueeze(1)
    h = OA._conv(sd, "conv_in.", x, 1)
    for i in range(len(ch)):
        for j in range(nres[i]):
            h = OA.resnet(sd, f"down_blocks.{i}.resnets.{j}.", h, groups, eps)
            if att[i]:
                h = OA.transformer(sd, f"down_blocks.{i}.attentions.{j}.", h, context, groups, eps, ch[i] // hc[i])
        h = OA.resnet(sd, f"down_blocks.{i}.downsampler.", h, groups, eps, down=True)
    h = h.reshape(h.shape[0], -1)
    h = F.relu(F.linear(h, sd["out.0.weight"], sd["out.0.bias"]))
    return F.linear(h, sd["out.3.weight"], sd["out.3.bias"])
