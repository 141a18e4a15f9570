"""Model-level parity of the CUDA UnetGenerator3d against the CPU oracle and the committed goldens.

Tolerances (bf16 operands, fp32 accumulation, fp32 master weights): synthesized-PET max-abs error
<= 3e-2 (tanh range), mean-abs <= 3e-3; loss abs error <= 2e-3; per-parameter grad-norm relative error <= 5e-2 and
global grad-norm relative error <= 2e-2.
"""
import os

import numpy as np
import pytest
import torch

from oracle import unet3d as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

OUT_MAX, OUT_MEAN, LOSS_ABS, GN_PARAM, GN_TOTAL = 3e-2, 3e-3, 2e-3, 5e-2, 2e-2


def synth_pair(shape, seed):
    g = torch.Generator().manual_seed(seed)
    n, d, h, w = shape
    return torch.rand(n, 1, d, h, w, generator=g), torch.rand(n, 1, d, h, w, generator=g)


def build(petsyn, ngf, seed=777):
    torch.manual_seed(seed)
    m = petsyn.UnetGenerator3d(1, 1, num_downs=4, ngf=ngf)
    return m


@pytest.mark.parametrize("name", ["unet3d_ngf32_2x32x48x32", "unet3d_ngf64_1x32x32x48"])
def test_train_step_matches_oracle_and_golden(name, petsyn):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    ngf, shape, seed = int(gold["ngf"]), tuple(int(v) for v in gold["shape"]), int(gold["seed"])
    model = build(petsyn, ngf, seed)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    # the seeded init must be the reference's (same RNG order): compare weight checksums with the golden
    for k, v in sd_cpu.items():
        if v.dtype.is_floating_point:
            assert abs(float(v.double().abs().sum()) - float(gold["wsum/" + k])) <= 1e-6 * max(1.0, float(gold["wsum/" + k])), k
    t1, pet = synth_pair(shape, seed)
    loss_o, y_o, grads_o, bufs_o = O.train_step(t1, pet, sd_cpu, num_downs=4, ngf=ngf)
    assert abs(float(loss_o) - float(gold["loss"])) < 1e-5          # oracle is pinned to the reference fixture

    model = model.cuda().train()
    y = model(t1.cuda())
    loss = torch.nn.functional.l1_loss(y, pet.cuda())
    loss.backward()
    torch.cuda.synchronize()
    yc = y.detach().cpu()
    err = (yc - y_o).abs()
    assert err.max().item() <= OUT_MAX and err.mean().item() <= OUT_MEAN, (err.max().item(), err.mean().item())
    gerr = (yc.numpy() - gold["output"])
    assert np.abs(gerr).max() <= OUT_MAX
    assert abs(loss.item() - float(gold["loss"])) <= LOSS_ABS
    tot, tot_ref = 0.0, 0.0
    for k, p in model.named_parameters():
        gn = p.grad.double().norm().item()
        ref = float(gold["gradnorm/" + k])
        tot += gn ** 2
        tot_ref += ref ** 2
        assert abs(gn - ref) <= GN_PARAM * ref + 1e-6, (k, gn, ref)
        # direction too.  nn.L1Loss' gradient is sign(y - t)/N: a bf16-sized output error flips the sign on the
        # ~1-2 % of voxels with |y - t| < 1e-2, so the direction bound under L1 is loose; the smooth-loss test below
        # (test_gradient_direction_smooth_loss) is the tight one.
        go = grads_o[k].double().flatten()
        cos = torch.dot(p.grad.double().cpu().flatten(), go) / (gn * go.norm().item() + 1e-30)
        assert cos.item() > 0.98, (k, cos.item())
    assert abs(tot ** 0.5 - tot_ref ** 0.5) <= GN_TOTAL * tot_ref ** 0.5
    # BatchNorm running statistics follow nn.BatchNorm3d (momentum 0.1, unbiased variance)
    sd = model.state_dict()
    for k in gold.files:
        if k.startswith("buffer/"):
            ref = torch.from_numpy(gold[k])
            got = sd[k[len("buffer/"):]].cpu()
            assert (got - ref).abs().max().item() <= 2e-2 * (ref.abs().max().item() + 1e-3), k

    # eval-mode (inference) forward with running statistics
    model.eval()
    with torch.no_grad():
        ye = model(t1.cuda()).cpu().numpy()
    assert np.abs(ye - gold["output_eval"]).max() <= 5e-2


def test_gradient_direction_smooth_loss(petsyn):
    """Per-parameter gradient direction against the fp32 oracle under a smooth loss (0.5*mean((y-t)^2)), calibrated
    with a PEER as SURVEY 8d prescribes: the same oracle graph run by PyTorch on the GPU under bf16 autocast (cuDNN).
    Accept when our angular error (1 - cos) is <= 2x the peer's, or below 5e-3 outright; global cosine >= 0.999."""
    ngf, shape, seed = 32, (2, 32, 32, 48), 11
    model = build(petsyn, ngf, seed)
    sd_cpu = {k: v.clone() for k, v in model.state_dict().items()}
    t1, pet = synth_pair(shape, seed)

    def oracle_grads(device, autocast):
        params = {k: v.clone().to(device).requires_grad_(True) for k, v in sd_cpu.items()
                  if v.dtype.is_floating_point and "running" not in k}
        full = {k: v.to(device) for k, v in sd_cpu.items()}
        full.update(params)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            y_ = O.forward(t1.to(device), full, num_downs=4, ngf=ngf, training=True)
        (0.5 * ((y_.float() - pet.to(device)) ** 2).mean()).backward()
        return {k: v.grad.double().cpu().flatten() for k, v in params.items()}

    ref = oracle_grads("cpu", False)
    peer = oracle_grads("cuda", True)
    model = model.cuda().train()
    y = model(t1.cuda())
    (0.5 * ((y - pet.cuda()) ** 2).mean()).backward()
    cosf = lambda a, b: (torch.dot(a, b) / (a.norm() * b.norm() + 1e-30)).item()
    dot = na = nb = 0.0
    for k, p in model.named_parameters():
        a, b = p.grad.double().cpu().flatten(), ref[k]
        ours, theirs = 1.0 - cosf(a, b), 1.0 - cosf(peer[k], b)
        assert ours <= max(5e-3, 2.0 * theirs), (k, ours, theirs)
        assert abs(a.norm().item() - b.norm().item()) <= GN_PARAM * b.norm().item() + 1e-9, k
        dot += torch.dot(a, b).item(); na += (a ** 2).sum().item(); nb += (b ** 2).sum().item()
    assert dot / (na * nb) ** 0.5 > 0.999


def test_state_dict_roundtrip_and_errors(petsyn):
    torch.manual_seed(1)
    m = petsyn.UnetGenerator3d(1, 1, num_downs=4, ngf=16)
    assert list(m.state_dict().keys()) == O.state_dict_keys(1, 1, 4, 16)
    ref_sd = O.init_state_dict(1, 1, 4, 16, seed=5)
    m.load_state_dict(ref_sd)                       # reference-format checkpoint loads unchanged
    with pytest.raises(AssertionError):
        petsyn.UnetGenerator3d(1, 2, num_downs=4)  # unet_model.py:12
    m = m.cuda()
    with pytest.raises(ValueError):
        m(torch.zeros(1, 1, 24, 32, 32, device="cuda"))   # 24 is not divisible by 16
    with pytest.raises(RuntimeError):
        m.cpu()(torch.zeros(1, 1, 32, 32, 32))             # no CPU path


def test_full_size_cfg1_scalars(petsyn):
    """BASELINE config 1 (ngf 64, 96x112x96, batch 1): loss / global grad-norm against the reference fixture."""
    gold = np.load(os.path.join(GOLD, "unet3d_ngf64_1x96x112x96.npz"))
    model = build(petsyn, 64, 777).cuda().train()
    t1, pet = synth_pair((1, 96, 112, 96), 777)
    y = model(t1.cuda())
    loss = torch.nn.functional.l1_loss(y, pet.cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(gold["loss"])) <= LOSS_ABS
    tot = sum(p.grad.double().norm().item() ** 2 for p in model.parameters()) ** 0.5
    assert abs(tot - float(gold["grad_norm_total"])) <= GN_TOTAL * float(gold["grad_norm_total"])
    samp = y.detach().cpu().numpy()[:, :, ::8, ::8, ::8]
    assert np.abs(samp - gold["output_sample"]).max() <= OUT_MAX


def test_trainer_eager_and_graph_match_autograd_adam(petsyn):
    """The fused train step (flat arenas + fused Adam), eager and CUDA-graph replayed, follows the same trajectory as
    the autograd path driven by torch.optim.Adam (train_unet.py:149-168 with L1 only)."""
    from petsyn_b200.train import Unet3dTrainer
    ngf, shape, seed, steps = 32, (1, 32, 32, 32), 3, 3
    batches = [synth_pair(shape, 100 + i) for i in range(steps)]

    def run(mode):
        model = build(petsyn, ngf, seed).cuda().train()
        losses = []
        if mode == "autograd":
            opt = torch.optim.Adam(model.parameters(), lr=5e-4)
            for t1, pet in batches:
                opt.zero_grad()
                loss = torch.nn.functional.l1_loss(model(t1.cuda()), pet.cuda())
                loss.backward()
                opt.step()
                losses.append(loss.item())
        else:
            tr = Unet3dTrainer(model, lr=5e-4, example_input=batches[0][0].cuda())
            if mode == "graph":
                tr.capture()
            for t1, pet in batches:
                losses.append(tr.step(t1.cuda(), pet.cuda()).item())
        return losses, {k: v.detach().float().cpu().clone() for k, v in model.state_dict().items()}

    la, sa = run("autograd")
    le, se = run("eager")
    lg, sg = run("graph")
    # fp32 add-reductions (split-K / weight-gradient TMA reduce, statistics atomics) complete in a run-dependent order, and
    # Adam's first steps turn the sign of every near-zero gradient element into a full lr-sized move: trajectories
    # agree to O(steps * lr) per weight, not bit-wise -- bounds below are ~2x the spread seen across repeated runs
    for a, b, c in zip(la, le, lg):
        assert abs(a - b) < 4e-3 and abs(a - c) < 4e-3, (la, le, lg)
    for k in sa:
        ref = sa[k]
        scale = ref.abs().max().item() + 1e-6
        # Adam's first steps move every weight by ~lr regardless of gradient scale, so compare against lr-sized motion
        tol = 8e-2 * scale + 3e-2 if "running" in k else 2e-3 * scale + 4e-3   # running stats of the 2x2x2 bottleneck amplify weight drift
        assert (se[k] - ref).abs().max().item() <= tol, k
        assert (sg[k] - se[k]).abs().max().item() <= tol, k
    assert int(sg["model.model.1.model.2.num_batches_tracked"]) == steps   # capture() restored the BN counters


# ---------------------------------------------------------------------------------------------------------------------
# The other constructor families (unet_model.py:42-45 InstanceNorm3d => biased convolutions; :17, :87-88 Dropout(0.5) with
# more than five levels) against fixtures from the live reference class and, with the dropout masks the CUDA path drew,
# against the oracle in training mode.
# ---------------------------------------------------------------------------------------------------------------------
def _family_model(petsyn, norm, affine, drop, nd, ngf, seed):
    import functools
    if norm == "instance":
        layer = functools.partial(torch.nn.InstanceNorm3d, affine=True) if affine else torch.nn.InstanceNorm3d
    else:
        layer = torch.nn.BatchNorm3d
    m = petsyn.UnetGenerator3d(1, 1, num_downs=nd, ngf=ngf, norm_layer=layer, use_dropout=drop)
    O.randomize_(m.state_dict(), seed)
    return m


FAMILIES = {
    "unet3d_instnorm_drop_nd6_1x64x64x64": ("instance", False, True, False),
    "unet3d_instaffine_nd5_2x32x32x32": ("instance", True, False, True),
    "unet3d_batchnorm_drop_nd6_1x64x64x64": ("batch", True, True, False),
}


def _check_grads(model, grads_ref, gradnorm_ref=None):
    tot = tot_ref = 0.0
    for k, p in model.named_parameters():
        assert p.grad is not None, k
        gn = p.grad.double().norm().item()
        ref = grads_ref[k].double().norm().item() if gradnorm_ref is None else float(gradnorm_ref["gradnorm/" + k])
        tot += gn ** 2
        tot_ref += ref ** 2
        # a convolution bias in front of a normalisation has a mathematically zero gradient (the norm removes the mean):
        # both sides hold rounding noise there
        if ref > 1e-3 * max(1.0, tot_ref ** 0.5):
            assert abs(gn - ref) <= 8e-2 * ref + 1e-5, (k, gn, ref)
            if grads_ref is not None:
                go = grads_ref[k].double().flatten()
                cos = torch.dot(p.grad.double().cpu().flatten(), go) / (gn * go.norm().item() + 1e-30)
                assert cos.item() > 0.97, (k, cos.item())
    assert abs(tot ** 0.5 - tot_ref ** 0.5) <= 3e-2 * tot_ref ** 0.5, (tot ** 0.5, tot_ref ** 0.5)


@pytest.mark.parametrize("name", sorted(FAMILIES))
def test_constructor_families_match_golden(name, petsyn):
    norm, affine, drop, train = FAMILIES[name]
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    shape, seed, nd, ngf = tuple(int(v) for v in gold["shape"]), int(gold["seed"]), int(gold["num_downs"]), int(gold["ngf"])
    model = _family_model(petsyn, norm, affine, drop, nd, ngf, seed)
    assert list(model.state_dict().keys()) == [str(k) for k in gold["keys"]]       # the reference's keys, in its order
    assert not model.default_family()
    t1, pet = synth_pair(shape, seed)
    model = model.cuda().train(train)
    if norm == "batch" and not train:
        # eval() BatchNorm is inference only (the backward kernels implement batch statistics): forward against the fixture
        with torch.no_grad():
            y = model(t1.cuda())
        assert np.abs(y.cpu().numpy()[:, :, ::2, ::2, ::2] - gold["output_sample"]).max() <= 5e-2
        with pytest.raises(NotImplementedError):
            torch.nn.functional.l1_loss(model(t1.cuda()), pet.cuda()).backward()
        return
    y = model(t1.cuda())
    loss = torch.nn.functional.l1_loss(y, pet.cuda())
    loss.backward()
    torch.cuda.synchronize()
    err = np.abs(y.detach().cpu().numpy()[:, :, ::2, ::2, ::2] - gold["output_sample"])
    # the coarsest normalisations of these small cases see 8 voxels per (sample, channel): bf16 rounding of the conv output is
    # amplified by the division by a tiny standard deviation, hence a mean bound twice the default family's
    assert err.max() <= 5e-2 and err.mean() <= 6e-3, (err.max(), err.mean())
    assert abs(loss.item() - float(gold["loss"])) <= LOSS_ABS
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    params = {k: v.requires_grad_(True) for k, v in sd_cpu.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd_cpu); full.update(params)
    yo = O.forward(t1, full, num_downs=nd, ngf=ngf, training=train, norm=norm)
    (yo - pet).abs().mean().backward()
    _check_grads(model, {k: p.grad for k, p in params.items()}, gold)


@pytest.mark.parametrize("norm", ["instance", "batch"])
def test_dropout_training_step_matches_oracle_with_the_same_masks(norm, petsyn):
    nd, ngf, shape, seed = 6, 8, (2, 64, 64, 64), 41
    model = _family_model(petsyn, norm, True, True, nd, ngf, seed).cuda().train()
    t1, pet = synth_pair(shape, seed)
    y = model(t1.cuda())
    loss = torch.nn.functional.mse_loss(y, pet.cuda())
    loss.backward()
    torch.cuda.synchronize()
    eng = model.engine_for(t1.cuda())
    from petsyn_b200.graph import DropoutOp
    drops = [op for op in eng.tape.ops if isinstance(op, DropoutOp)]
    assert len(drops) == nd - 5 and all(op.mask is not None for op in drops)
    masks = {}
    for op in drops:                      # the op acts on the up half of cat_i: level i = log2(D / buffer depth)
        b = op.x
        lvl = int(np.log2(shape[1] // b.d))
        m = op.mask.float().view(b.n, b.d, b.h, b.w, -1).permute(0, 4, 1, 2, 3).cpu()
        frac = (m == 0).float().mean().item()
        assert 0.4 < frac < 0.6 and set(m.unique().tolist()) <= {0.0, 2.0}
        masks[lvl] = m
    sd_cpu = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    if norm == "batch":                   # the forward above already moved the running statistics: oracle starts from the initial ones
        for k in sd_cpu:
            if k.endswith("running_mean"): sd_cpu[k].zero_()
            if k.endswith("running_var"): sd_cpu[k].fill_(1.0)
    params = {k: v.requires_grad_(True) for k, v in sd_cpu.items() if v.dtype.is_floating_point and "running_" not in k}
    full = dict(sd_cpu); full.update(params)
    bufs = {}
    yo = O.forward(t1, full, num_downs=nd, ngf=ngf, training=True, norm=norm, dropout_masks=masks, new_buffers=bufs)
    lo = torch.nn.functional.mse_loss(yo, pet)
    lo.backward()
    err = (y.detach().cpu() - yo.detach()).abs()
    assert err.max().item() <= 5e-2 and err.mean().item() <= 6e-3, (err.max().item(), err.mean().item())
    assert abs(loss.item() - lo.item()) <= LOSS_ABS
    _check_grads(model, {k: p.grad for k, p in params.items()})
    if norm == "batch":
        sd = model.state_dict()
        for k, v in bufs.items():
            if "running" in k:
                assert (sd[k].cpu() - v).abs().max().item() <= 2e-2 * (v.abs().max().item() + 1e-3), k
    # eval(): the Dropout layers are the identity
    model.eval()
    with torch.no_grad():
        model(t1.cuda())
    assert all(op.mask is None for op in drops)
