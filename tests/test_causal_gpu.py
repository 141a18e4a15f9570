"""The two decoders of the causal_synthesis model -- LABELLED RESTATEMENTS, parity UNPINNED: the classes live
in the authors' un-vendored fork.  These tests check the CUDA path against ``oracle/causal.py``, a plain-PyTorch
statement of the SAME restated graphs (``training_causal.json:40-74``), peer-calibrated against bf16 autocast: output, the
gradient w.r.t. the LATENT (the decoders compose with an encoder through autograd), parameter gradients; then one step of
``train_unify_causal_gen.py:213-247``'s generator losses (reparameterisation, L1 on both reconstructions, KL as written)."""
import pytest
import torch

from oracle import atten_unet as OA
from oracle import causal as OC

pytestmark = pytest.mark.gpu


def _latent(n, d, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, d, h, w, generator=g), torch.rand(n, 1, 5, generator=g), \
        torch.rand(n, 1, 8 * d, 8 * h, 8 * w, generator=g)


def _check(model, fwd_oracle, z, ctx, tgt, what):
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    po = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    zo = z.clone().requires_grad_(True)
    yo = fwd_oracle(zo, po)
    lo = (yo - tgt).abs().mean()
    lo.backward()
    pp = {k: v.clone().cuda().requires_grad_(True) for k, v in sd.items()}
    zp = z.clone().cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        yp = fwd_oracle(zp, pp)
    (yp.float() - tgt.cuda()).abs().mean().backward()
    model = model.cuda().train()
    zc = z.clone().cuda().requires_grad_(True)
    y = model(zc) if ctx is None else model(zc, ctx.cuda())
    loss = torch.nn.functional.l1_loss(y, tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    err, perr = (y.detach().cpu() - yo.detach()).abs(), (yp.detach().float().cpu() - yo.detach()).abs()
    print(what, "out err ours max/mean", err.max().item(), err.mean().item(), "peer", perr.max().item(), perr.mean().item(),
          "loss", loss.item(), lo.item())
    assert y.shape == tgt.shape
    assert err.max().item() <= 2.0 * perr.max().item() + 5e-3 and err.mean().item() <= 2.0 * perr.mean().item() + 5e-4
    assert abs(loss.item() - lo.item()) <= 2e-3
    a, b, c = zc.grad.double().cpu().flatten(), zo.grad.double().flatten(), zp.grad.double().cpu().flatten()
    cos, cos_p = (torch.dot(a, b) / (a.norm() * b.norm())).item(), (torch.dot(c, b) / (c.norm() * b.norm())).item()
    print(what, "latent gradient: norm ours/oracle/peer", a.norm().item(), b.norm().item(), c.norm().item(), "cos", cos, cos_p)
    assert 1 - cos <= max(2.0 * (1 - cos_p), 2e-2)
    assert abs(a.norm() - b.norm()).item() <= max(2.0 * abs(c.norm() - b.norm()).item(), 5e-2 * b.norm().item())
    tot = tot_o = tot_p = 0.0
    for k, p in model.named_parameters():
        go = po[k].grad if po[k].grad is not None else torch.zeros_like(po[k])
        gp = pp[k].grad if pp[k].grad is not None else torch.zeros_like(pp[k])
        tot += p.grad.double().norm().item() ** 2; tot_o += go.double().norm().item() ** 2; tot_p += gp.double().norm().item() ** 2
    print(what, "grad-norm ours/oracle/peer", tot ** 0.5, tot_o ** 0.5, tot_p ** 0.5)
    assert abs(tot ** 0.5 - tot_o ** 0.5) <= max(2.0 * abs(tot_p ** 0.5 - tot_o ** 0.5), 3e-2 * tot_o ** 0.5)


def test_t1_decoder_matches_restated_oracle(petsyn):
    torch.manual_seed(1)
    dec = petsyn.Decoder(**OC.T1_DECODER)
    assert sum(p.numel() for p in dec.parameters()) == 1808321 and "blocks.0.conv.weight" in dec.state_dict()
    OA.randomize_(dec.named_parameters(), seed=7)
    z, _, tgt = _latent(2, 4, 6, 4, 3)
    _check(dec, lambda zz, sd: OC.t1_decoder_forward(zz, sd), z, None, tgt, "T1 decoder 2x3x4x6x4 -> 32x48x32:")


def test_pet_decoder_matches_restated_oracle(petsyn):
    torch.manual_seed(2)
    dec = petsyn.DiffusionModelDecoder(**OC.PET_DECODER)
    OA.randomize_(dec.named_parameters(), seed=8)
    z, ctx, tgt = _latent(2, 4, 6, 4, 4)
    _check(dec, lambda zz, sd: OC.pet_decoder_forward(zz, ctx.to(zz.device), sd), z, ctx, tgt,
           "PET decoder 2x3x4x6x4 -> 32x48x32:")
    sd = dec.state_dict()
    zero = [k for k in sd if ".attn2.to_q." in k or ".attn2.to_k." in k or ".transformer_blocks.0.norm2." in k]
    named = dict(dec.named_parameters())
    assert len(zero) == 2 * 4 and all(float(named[k].grad.abs().max()) == 0.0 for k in zero)


def test_causal_generator_losses_at_the_reference_crop(petsyn):
    """train_unify_causal_gen.py:213-247 from the latent on, at the reference sizes (batch 2, latent 6x12x16x12, volumes
    96x128x96): reparameterisation, both decoders, L1 + L1 + kl_weight * KL(z_mu, z_sigma) as written, one backward."""
    torch.manual_seed(3)
    t1_dec, pet_dec = petsyn.Decoder(**OC.T1_DECODER).cuda().train(), petsyn.DiffusionModelDecoder(**OC.PET_DECODER).cuda().train()
    OA.randomize_(t1_dec.named_parameters(), seed=1)
    OA.randomize_(pet_dec.named_parameters(), seed=2)
    g = torch.Generator().manual_seed(5)
    latent = (torch.randn(2, 6, 12, 16, 12, generator=g) * 0.5).cuda().requires_grad_(True)      # stands for t1_encoder(t1_img)
    t1, pet, info = torch.rand(2, 1, 96, 128, 96, generator=g).cuda(), torch.rand(2, 1, 96, 128, 96, generator=g).cuda(), \
        torch.rand(2, 1, 5, generator=g).cuda()
    z_mu, z_sigma = latent[:, :3], latent[:, 3:]
    t1_rec = t1_dec(petsyn.reparameterize(z_mu, z_sigma))
    rec_pet = pet_dec(petsyn.reparameterize(z_mu, z_sigma), info)
    loss = torch.nn.functional.l1_loss(rec_pet, pet) + torch.nn.functional.l1_loss(t1_rec, t1) \
        + 0.001 * petsyn.kl_divergence(z_mu, z_sigma)
    loss.backward()
    torch.cuda.synchronize()
    assert t1_rec.shape == t1.shape and rec_pet.shape == pet.shape
    assert torch.isfinite(loss) and torch.isfinite(latent.grad).all() and float(latent.grad.abs().max()) > 0
    for m in (t1_dec, pet_dec):
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    assert abs(petsyn.kl_divergence(z_mu, z_sigma).item() - OC.kl_divergence(z_mu.detach().cpu(), z_sigma.detach().cpu()).item()) < 1e-3
