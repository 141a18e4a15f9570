"""configs[4] building block: the classifier restatement (petsyn.DiffusionModelEncoder) against its CPU restatement
(oracle/classifier.py).  BOTH are restatements of a class the reference does not ship in runnable form:
this test pins the CUDA path to the restated algorithm, not to the authors' fork -- parity unpinned."""
import pytest
import torch

from oracle import atten_unet as OA
from oracle import classifier as OC

pytestmark = pytest.mark.gpu


def test_classifier_logits_match_restatement(petsyn):
    cfg = dict(OC.TRAINING_ATTEN_JSON)
    shape = (2, 1, 32, 64, 32)                       # -> 128 x 1x2x1 = 256 features after five down-samplings
    model = petsyn.DiffusionModelEncoder(**cfg, head_in_features=256).eval()
    OA.randomize_(model.named_parameters(), seed=13)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert {"conv_in.conv.weight", "time_embed.0.weight", "down_blocks.4.downsampler.conv1.conv.weight",
            "down_blocks.3.attentions.1.transformer_blocks.0.attn2.to_v.weight", "out.0.weight", "out.3.bias"} <= set(sd)
    g = torch.Generator().manual_seed(13)
    x, ctx = torch.rand(shape, generator=g), torch.rand(2, 1, 5, generator=g)
    ref = OC.forward(x, ctx, sd, cfg)
    # peer: the same restatement under bf16 autocast on the GPU
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        peer = OC.forward(x.cuda(), ctx.cuda(), {k: v.cuda() for k, v in sd.items()}, cfg).float().cpu()
    model = model.cuda()
    with torch.no_grad():
        out = model(x.cuda(), torch.zeros(2, device="cuda"), ctx.cuda()).cpu()
    err, err_peer = (out - ref).abs().max().item(), (peer - ref).abs().max().item()
    print("logits ours/ref/peer", out.tolist(), ref.tolist(), peer.tolist())
    assert out.shape == (2, 2)
    assert err <= 2.0 * err_peer + 2e-2 * ref.abs().max().item(), (err, err_peer)


def test_classifier_contracts(petsyn):
    cfg = dict(OC.TRAINING_ATTEN_JSON)
    with pytest.raises(ValueError):
        petsyn.DiffusionModelEncoder(**{**cfg, "cross_attention_dim": None})
    m = petsyn.DiffusionModelEncoder(**cfg).cuda().eval()                 # head_in_features = 4096 as written (:1987)
    with torch.no_grad(), pytest.raises(ValueError):                      # 96x128x96 -> 4608 features
        m(torch.rand(1, 1, 96, 128, 96, device="cuda"), None, torch.rand(1, 1, 5, device="cuda"))
    with pytest.raises(NotImplementedError):                              # inference only
        m(torch.rand(1, 1, 32, 64, 32, device="cuda"), None, torch.rand(1, 1, 5, device="cuda"))
