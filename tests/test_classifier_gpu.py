"""configs[4] building block: the classifier restatement (petsyn.DiffusionModelEncoder) against its CPU restatement
(oracle/classifier.py).  BOTH are restatements of a class the reference does not ship in runnable form:
these tests pin the CUDA path (inference logits and the training step) to the restated algorithm, not to the authors' fork --
parity unpinned."""
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


def test_classifier_training_step_matches_restatement(petsyn):
    """The step of pet_for_classification/train_atten_encoder_MCI.py:169-175 -- logits -> weighted cross-entropy -> backward --
    through the public module + autograd, against torch autograd on the CPU restatement (Dropout switched off on both sides:
    its mask is random; a second pass with the reference's p = 0.1 checks that training mode runs and drops units)."""
    cfg = dict(OC.TRAINING_ATTEN_JSON)
    shape = (2, 1, 32, 64, 32)
    model = petsyn.DiffusionModelEncoder(**cfg, head_in_features=256).train()
    OA.randomize_(model.named_parameters(), seed=21)
    model.out[2].p = 0.0
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(21)
    x, ctx = torch.rand(shape, generator=g), torch.rand(2, 1, 5, generator=g)
    gts = torch.tensor([0, 1])
    wts = torch.tensor([1.0, 2.5])                                         # weighted CE (class imbalance, :120-130)
    po = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = torch.nn.functional.cross_entropy(OC.forward(x, ctx, po, cfg), gts, weight=wts)
    lo.backward()
    pp = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lp = torch.nn.functional.cross_entropy(OC.forward(x.cuda(), ctx.cuda(), pp, cfg).float(), gts.cuda(), weight=wts.cuda())
    lp.backward()
    model = model.cuda()
    pred = model(x.cuda(), torch.zeros(2, device="cuda"), ctx.cuda())
    loss = torch.nn.functional.cross_entropy(pred, gts.cuda(), weight=wts.cuda())
    loss.backward()
    torch.cuda.synchronize()
    print("classifier CE ours/oracle/peer", loss.item(), lo.item(), lp.item())
    assert abs(loss.item() - lo.item()) <= max(2.0 * abs(lp.item() - lo.item()), 2e-2 * abs(lo.item()))
    tot = tot_o = tot_p = 0.0
    for k, p in model.named_parameters():
        go = po[k].grad if po[k].grad is not None else torch.zeros_like(po[k])
        gp = pp[k].grad if pp[k].grad is not None else torch.zeros_like(pp[k])
        tot += p.grad.double().norm().item() ** 2; tot_o += go.double().norm().item() ** 2; tot_p += gp.double().norm().item() ** 2
        if k.startswith("time_embed."):
            assert float(p.grad.abs().max()) == 0.0, k
    print("classifier grad-norm ours/oracle/peer", tot ** 0.5, tot_o ** 0.5, tot_p ** 0.5)
    assert abs(tot ** 0.5 - tot_o ** 0.5) <= max(2.0 * abs(tot_p ** 0.5 - tot_o ** 0.5), 5e-2 * tot_o ** 0.5)
    named = dict(model.named_parameters())
    for k in ("out.0.weight", "out.3.weight", "conv_in.conv.weight", "down_blocks.4.resnets.1.conv2.conv.weight"):
        a, b = named[k].grad.double().cpu().flatten(), po[k].grad.double().flatten()
        assert (torch.dot(a, b) / (a.norm() * b.norm())).item() > 0.97, k
    # the reference's Dropout(0.1) in training mode: runs, drops about a tenth of the 512 hidden units, gradients finite
    model.out[2].p = 0.1
    model.zero_grad()
    torch.nn.functional.cross_entropy(model(x.cuda(), None, ctx.cuda()), gts.cuda(), weight=wts.cuda()).backward()
    eng = next(iter(model._engines.values()))
    drop = [op for op in eng.tape.ops if type(op).__name__ == "DropoutOp"][0]
    frac = float((drop.mask == 0).float().mean())
    assert 0.03 < frac < 0.2 and all(torch.isfinite(p.grad).all() for p in model.parameters())
