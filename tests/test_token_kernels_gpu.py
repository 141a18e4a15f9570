"""Kernel-level tests of the token-stream / resampling kernels of the covariate-conditioned generator, each against the
same op in plain PyTorch fp32 on the SAME bf16-rounded operands (VERDICT r1 "weak" #5): LayerNorm fwd / bwd (widths that
are and are not multiples of 32), GEGLU fwd / bwd, the covariate bias (cross-attention over a length-1 context) fwd / bwd,
2x resampling (average pooling / nearest up-sampling on channel slices, with and without accumulation)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _lib():
    from petsyn_b200._cabi import check, lib, ptr, stream_ptr
    return check, lib, ptr, stream_ptr


def bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("rows,c", [(4608, 128), (1000, 64), (333, 16), (77, 40)])
def test_layernorm_fwd_bwd(petsyn, rows, c):
    check, lib, ptr, sp = _lib()
    g = torch.Generator(device=DEV).manual_seed(rows + c)
    x = bf(torch.randn(rows, c, device=DEV, generator=g) * 2 + 0.5)
    dy = bf(torch.randn(rows, c, device=DEV, generator=g))
    gamma = 1 + 0.1 * torch.randn(c, device=DEV, generator=g)
    beta = 0.1 * torch.randn(c, device=DEV, generator=g)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    check(lib.petsyn_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows, c, 1e-5, sp()))
    xr = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.layer_norm(xr, (c,), gr, br, 1e-5)
    assert (y.float() - yr).abs().max().item() <= 2e-2 * (1 + yr.abs().max().item())       # one bf16 rounding of the output
    assert torch.allclose(mean, xr.detach().mean(1), atol=1e-4, rtol=1e-4)
    yr.backward(dy.float())
    for accumulate in (False, True):
        dx = bf(torch.full((rows, c), 0.25, device=DEV)) if accumulate else torch.empty_like(x)
        dgam, dbet = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
        check(lib.petsyn_layernorm_bwd(ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(dgam), ptr(dbet), rows, c,
                                       int(accumulate), sp()))
        want = xr.grad + (0.25 if accumulate else 0.0)
        assert (dx.float() - want).abs().max().item() <= 2e-2 * (1 + want.abs().max().item())
        assert torch.allclose(dgam, gr.grad, rtol=2e-3, atol=2e-3 * gr.grad.abs().max().item())
        assert torch.allclose(dbet, br.grad, rtol=2e-3, atol=2e-3 * br.grad.abs().max().item())
    # reproducible: a second launch gives the same bits
    d2, b2 = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    check(lib.petsyn_layernorm_bwd(ptr(x), ptr(dy), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(d2), ptr(b2), rows, c, 0, sp()))
    assert torch.equal(d2, dgam) and torch.equal(b2, dbet)


@pytest.mark.parametrize("rows,f", [(4608, 512), (129, 64)])
def test_geglu_fwd_bwd(petsyn, rows, f):
    check, lib, ptr, sp = _lib()
    g = torch.Generator(device=DEV).manual_seed(f)
    h = bf(torch.randn(rows, 2 * f, device=DEV, generator=g) * 1.5)
    dout = bf(torch.randn(rows, f, device=DEV, generator=g))
    out, dh = torch.empty(rows, f, dtype=torch.bfloat16, device=DEV), torch.empty_like(h)
    check(lib.petsyn_geglu_fwd(ptr(h), ptr(out), rows, f, sp()))
    hr = h.float().requires_grad_(True)
    a, gate = hr.chunk(2, dim=-1)
    ref = a * F.gelu(gate)                                            # exact erf GELU (MONAI MLPBlock act="GEGLU")
    assert (out.float() - ref).abs().max().item() <= 1e-2 * (1 + ref.abs().max().item())
    ref.backward(dout.float())
    check(lib.petsyn_geglu_bwd(ptr(h), ptr(dout), ptr(dh), rows, f, sp()))
    assert (dh.float() - hr.grad).abs().max().item() <= 1e-2 * (1 + hr.grad.abs().max().item())


@pytest.mark.parametrize("n,L,c,cctx", [(2, 2304, 128, 5), (3, 130, 64, 6), (1, 77, 16, 3)])
def test_covariate_bias_fwd_bwd(petsyn, n, L, c, cctx):
    """tokens += to_out(to_v(context)) per sample (atten_unet_model.py:156-175 with a length-1 context)."""
    check, lib, ptr, sp = _lib()
    g = torch.Generator(device=DEV).manual_seed(c + L)
    ctx = torch.rand(n, cctx, device=DEV, generator=g)
    wv, wo = torch.randn(c, cctx, device=DEV, generator=g) * 0.3, torch.randn(c, c, device=DEV, generator=g) * 0.1
    bo = torch.randn(c, device=DEV, generator=g) * 0.1
    tok = bf(torch.randn(n * L, c, device=DEV, generator=g))
    t0 = tok.clone()
    vbuf, bias = torch.empty(n, c, device=DEV), torch.empty(n, c, device=DEV)
    check(lib.petsyn_covariate_bias_fwd(ptr(ctx), ptr(wv), ptr(wo), ptr(bo), ptr(vbuf), ptr(bias), ptr(tok), n, cctx, c, L, sp()))
    wvr, wor, bor = (t.clone().requires_grad_(True) for t in (wv, wo, bo))
    b_ref = F.linear(F.linear(ctx, wvr), wor, bor)                                            # [n, c]
    assert torch.allclose(bias, b_ref.detach(), rtol=1e-4, atol=1e-5)
    want = t0.float().view(n, L, c) + b_ref.detach()[:, None, :]
    assert (tok.float().view(n, L, c) - want).abs().max().item() <= 1e-2 * (1 + want.abs().max().item())
    dt = bf(torch.randn(n * L, c, device=DEV, generator=g))
    (dt.float().view(n, L, c).sum(1) * b_ref).sum().backward()
    dbias, dwv, dwo, dbo = torch.empty(n, c, device=DEV), torch.empty_like(wv), torch.empty_like(wo), torch.empty_like(bo)
    check(lib.petsyn_covariate_bias_bwd(ptr(ctx), ptr(wo), ptr(vbuf), ptr(dt), ptr(dbias), ptr(dwv), ptr(dwo), ptr(dbo), n, cctx,
                                        c, L, sp()))
    for a, b in ((dwv, wvr.grad), (dwo, wor.grad), (dbo, bor.grad)):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-3 * b.abs().max().item()), (a - b).abs().max().item()


@pytest.mark.parametrize("c,coff_s,cs_s,coff_d,cs_d", [(16, 0, 16, 0, 16), (16, 16, 32, 8, 48), (32, 0, 32, 32, 64)])
def test_resample2_on_channel_slices(petsyn, c, coff_s, cs_s, coff_d, cs_d):
    check, lib, ptr, sp = _lib()
    n, d, h, w = 2, 6, 10, 8
    g = torch.Generator(device=DEV).manual_seed(c + cs_d)
    big = bf(torch.randn(n, d, h, w, cs_s, device=DEV, generator=g))
    small = bf(torch.randn(n, d // 2, h // 2, w // 2, cs_s, device=DEV, generator=g))
    ncdhw = lambda t, off: t[..., off:off + c].float().permute(0, 4, 1, 2, 3)
    for accumulate in (0, 1):
        # average pooling: big -> small grid
        dst = bf(torch.randn(n, d // 2, h // 2, w // 2, cs_d, device=DEV, generator=g))
        keep = dst.clone()
        check(lib.petsyn_resample2(ptr(big), cs_s, coff_s, ptr(dst), cs_d, coff_d, n, d // 2, h // 2, w // 2, c, 0, 0.125,
                                   accumulate, sp()))
        want = F.avg_pool3d(ncdhw(big, coff_s), 2, 2) + (ncdhw(keep, coff_d) if accumulate else 0)
        assert (ncdhw(dst, coff_d) - want).abs().max().item() <= 1e-2 * (1 + want.abs().max().item())
        mask = torch.ones(cs_d, dtype=torch.bool, device=DEV)
        mask[coff_d:coff_d + c] = False
        assert torch.equal(dst[..., mask], keep[..., mask])                       # the other channels are untouched
        # nearest up-sampling: small -> big grid
        dst = bf(torch.randn(n, d, h, w, cs_d, device=DEV, generator=g))
        keep = dst.clone()
        check(lib.petsyn_resample2(ptr(small), cs_s, coff_s, ptr(dst), cs_d, coff_d, n, d, h, w, c, 1, 1.0, accumulate, sp()))
        want = F.interpolate(ncdhw(small, coff_s), scale_factor=2.0, mode="nearest") + (ncdhw(keep, coff_d) if accumulate else 0)
        assert (ncdhw(dst, coff_d) - want).abs().max().item() <= 1e-2 * (1 + want.abs().max().item())
        assert torch.equal(dst[..., mask], keep[..., mask])
