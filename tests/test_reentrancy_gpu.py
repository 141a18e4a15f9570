"""The drop-in modules hold ONE set of saved activations per (shape, device) engine.  These tests replay the call patterns
of the reference scripts that put two forward passes of the same module in front of one backward pass, through the public
``nn.Module`` API + autograd, and check them against the CPU oracle:

  * ``train_bmgan.py:170-180``  E phase: ``encoder(pet_img)``, ``encoder(fake_pet)``, summed KL, ONE ``backward()``;
  * ``train_bmgan.py:243-247``  evaluation: ``discriminator(fake)`` then ``discriminator(real)`` under ``no_grad`` before
    either result is consumed (the outputs must not alias an engine-owned buffer);
  * D(fake) + D(real) summed into one loss (the pattern of ``train_unet.py:179-184`` when written with one backward);
  * ``UnetGenerator3d`` (BatchNorm): the re-run that restores the first node's activations must not move the running
    statistics a second time; backward through ``eval()``-mode BatchNorm raises instead of returning batch-statistics
    gradients.
"""
import copy

import pytest
import torch

from oracle import bmgan as OB
from oracle import unet3d as OU

pytestmark = pytest.mark.gpu


def _gnorm(grads):
    return sum(g.double().norm().item() ** 2 for g in grads) ** 0.5


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()


def test_encoder_two_forwards_one_backward_matches_oracle(petsyn):
    """train_bmgan.py:170-180 through petsyn.ResNet_encoder + autograd."""
    torch.manual_seed(11)
    enc = petsyn.ResNet_encoder().train()
    oe = OB.ResNetEncoder().train()
    oe.load_state_dict(enc.state_dict())
    g = torch.Generator().manual_seed(11)
    pet, fake = torch.rand(1, 1, 128, 128, 128, generator=g), torch.rand(1, 1, 128, 128, 128, generator=g) * 2 - 1

    mu_r, lv_r = oe(pet)
    mu_f, lv_f = oe(fake)
    lo = (OB.kl_divergence(mu_r, lv_r) + OB.kl_divergence(mu_f, lv_f)).mean()        # :174-176
    lo.backward()
    ref = {k: p.grad.clone() for k, p in oe.named_parameters()}

    enc = enc.cuda()
    mu_r, lv_r = enc(pet.cuda())
    mu_f, lv_f = enc(fake.cuda())                      # overwrites the engine's activations of the first call
    loss = (OB.kl_divergence(mu_r, lv_r) + OB.kl_divergence(mu_f, lv_f)).mean()
    loss.backward()                                    # ONE backward through both nodes
    torch.cuda.synchronize()
    ours = {k: p.grad.clone() for k, p in enc.named_parameters()}

    # the same quantity computed the way BmganTrainer does it (forward, backward, forward, backward): must agree closely
    enc.zero_grad()
    for vol in (pet, fake):
        mu, lv = enc(vol.cuda())
        OB.kl_divergence(mu, lv).mean().backward()
    seq = {k: p.grad.clone() for k, p in enc.named_parameters()}

    assert abs(loss.item() - lo.item()) <= 2e-2 * abs(lo.item()) + 1e-3
    tot, tot_ref = _gnorm(ours.values()), _gnorm(ref.values())
    tot_seq = _gnorm(seq.values())
    print("E phase: loss", loss.item(), lo.item(), "grad-norm ours", tot, "sequential", tot_seq, "oracle", tot_ref)
    # reductions are reproducible (csrc/det_reduce.cuh), so the restored-activation path must give the SAME numbers as the
    # interleaved one, not merely close ones (fp32 sums of the two nodes' gradients are formed in the same order by autograd)
    assert abs(tot - tot_seq) <= 1e-6 * tot_seq
    assert abs(tot - tot_ref) <= 5e-2 * tot_ref
    big = sorted(ref, key=lambda k: ref[k].norm().item(), reverse=True)[:8]
    for k in big:
        # against the fp32 oracle: this encoder ends in 2x2x2 = 8-voxel InstanceNorms, its gradients are ill-conditioned
        assert _cos(ours[k], ref[k]) > 0.95, (k, _cos(ours[k], ref[k]))
        assert _cos(ours[k], seq[k]) > 0.999999, (k, _cos(ours[k], seq[k]))


def test_discriminator_no_grad_outputs_do_not_alias(petsyn):
    """train_bmgan.py:243-247: logits_fake must survive the call that computes logits_real."""
    torch.manual_seed(3)
    disc = petsyn.patch_discriminator().cuda().eval()
    gen = petsyn.dense_unet_generator(input_conv_channel=16, output_conv_channel=16, down_channels=[16, 16, 16, 16],
                                      middle_channels=[16], up_channels=[16, 16, 16, 16, 16]).cuda().eval()
    g = torch.Generator().manual_seed(3)
    a, b = torch.rand(1, 1, 64, 64, 64, generator=g).cuda(), (torch.rand(1, 1, 64, 64, 64, generator=g) * 2 - 1).cuda()
    with torch.no_grad():
        la = disc(a)
        keep = la.clone()
        lb = disc(b)
        assert la.data_ptr() != lb.data_ptr()
        assert torch.equal(la, keep)
        assert not torch.equal(la, lb)
        z = torch.randn(1, 8, generator=g).cuda()
        ya = gen(a, z)
        keep = ya.clone()
        yb = gen(b, z)
        assert ya.data_ptr() != yb.data_ptr() and torch.equal(ya, keep) and not torch.equal(ya, yb)


def test_discriminator_fake_plus_real_one_backward(petsyn):
    torch.manual_seed(5)
    disc = petsyn.patch_discriminator().train()
    od = OB.PatchDiscriminatorWrapper().train()
    od.load_state_dict(disc.state_dict())
    g = torch.Generator().manual_seed(5)
    fake, real = torch.rand(2, 1, 64, 64, 64, generator=g) * 2 - 1, torch.rand(2, 1, 64, 64, 64, generator=g)
    lo = 0.5 * (OB.lsgan_loss(od(fake), False) + OB.lsgan_loss(od(real), True))
    lo.backward()
    disc = disc.cuda()
    loss = 0.5 * (OB.lsgan_loss(disc(fake.cuda()), False) + OB.lsgan_loss(disc(real.cuda()), True))
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - lo.item()) <= 2e-2 * abs(lo.item())
    ref = {k: p.grad for k, p in od.named_parameters()}
    ours = {k: p.grad for k, p in disc.named_parameters()}
    tot, tot_ref = _gnorm(ours.values()), _gnorm(ref.values())
    print("D fake+real: loss", loss.item(), lo.item(), "grad-norm", tot, tot_ref)
    assert abs(tot - tot_ref) <= 0.1 * tot_ref
    for k in sorted(ref, key=lambda k: ref[k].norm().item(), reverse=True)[:4]:
        assert _cos(ours[k], ref[k]) > 0.98, (k, _cos(ours[k], ref[k]))
    # BatchNorm running statistics moved exactly twice (two forward calls), not three times (the restoring re-run)
    for (k, v), (_, vo) in zip(disc.state_dict().items(), od.state_dict().items()):
        if k.endswith("num_batches_tracked"):
            assert int(v) == int(vo) == 2, (k, int(v), int(vo))
        elif "running" in k:
            assert torch.allclose(v.cpu(), vo, rtol=2e-2, atol=2e-3), k


def test_unet3d_two_forwards_one_backward_and_eval_backward(petsyn):
    torch.manual_seed(9)
    ngf, shape = 16, (1, 1, 32, 32, 32)
    model = petsyn.UnetGenerator3d(1, 1, num_downs=4, ngf=ngf).train()
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(9)
    xa, xb, tgt = (torch.rand(shape, generator=g) for _ in range(3))
    # oracle: gradients of L1(G(xa)) + L1(G(xb)) are the sum of the two single-input gradients
    _, _, ga, sd1 = OU.train_step(xa, tgt, sd0, num_downs=4, ngf=ngf)
    sd_mid = dict(sd0)
    sd_mid.update({k: v for k, v in sd1.items() if "running" in k or "num_batches" in k})
    _, _, gb, sd2 = OU.train_step(xb, tgt, sd_mid, num_downs=4, ngf=ngf)
    model = model.cuda()
    ya, yb = model(xa.cuda()), model(xb.cuda())
    loss = torch.nn.functional.l1_loss(ya, tgt.cuda()) + torch.nn.functional.l1_loss(yb, tgt.cuda())
    loss.backward()
    torch.cuda.synchronize()
    tot = _gnorm([p.grad for p in model.parameters()])
    tot_ref = _gnorm([ga[k] + gb[k] for k in ga])
    print("unet3d two forwards: grad-norm", tot, tot_ref)
    assert abs(tot - tot_ref) <= 3e-2 * tot_ref
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == 2, (k, int(v))
        elif "running" in k:
            assert torch.allclose(v.cpu(), sd2[k], rtol=2e-2, atol=2e-3), k
    # eval(): BatchNorm is affine in its input; the batch-statistics backward would be silently wrong -> must raise
    model.eval()
    y = model(xa.cuda())
    with pytest.raises(NotImplementedError):
        y.mean().backward()
