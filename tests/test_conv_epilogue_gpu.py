"""Fused epilogues of the small-channel 3x3x3 convolutions (petsyn_conv_fprop_epi / petsyn_conv_dgrad_epi) against the separate
passes they replace: the residual sum + GroupNorm statistics of a ResnetBlock (atten_unet_model.py:641-662) and the reduction
pass of the normalisation backward.  The stored tensors must be bit-identical to the un-fused kernels' (same tile arithmetic);
the sums are compared with float64 sums of the stored bf16 values."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def from_ndhwc(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


# n, d, h, w, cin, cout: exact tiles / ragged in w and h (tile = 8 (w) x 16 (h)) / two channel atoms either side
CASES = [
    (2, 16, 32, 48, 16, 16),
    (2, 13, 36, 28, 16, 16),
    (2, 12, 32, 40, 32, 16),
    (2, 12, 40, 28, 16, 32),
    (3, 12, 32, 40, 32, 32),
    (2, 12, 32, 40, 48, 16),
]


@pytest.mark.parametrize("n,d,h,w,cin,cout", CASES)
def test_fprop_residual_and_statistics(petsyn, n, d, h, w, cin, cout):
    from petsyn_b200._cabi import ConvEpilogue, ptr
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5 + cin + cout)
    x = torch.randn(n, d, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    # the output is a channel slice of a wider (concat) buffer; the residual another slice of another buffer
    ybuf = torch.zeros(n, d, h, w, cout + 16, dtype=torch.bfloat16, device=dev)
    res = torch.randn(n, d, h, w, cout + 8, generator=g).to(dev).to(torch.bfloat16)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, 3, 1, 1, y_cstride=cout + 16, y_coff=16)
    assert plan.kernel_path[0] == 1 and plan.epi_ok[0]
    plan.pack(wt, need_dgrad=False)
    # un-fused: conv, then the sum in fp32 of the bf16 conv output?  No -- the fused kernel adds BEFORE rounding, as a conv
    # with a fused residual would; reference = bf16(conv_fp32 + bias + res)
    ref = F.conv3d(from_ndhwc(x).float(), wt.to(torch.bfloat16).float(), bias, padding=1) + from_ndhwc(res[..., 8:]).float()
    st1 = torch.zeros(n, 2, cout + 16, dtype=torch.float64, device=dev)      # consumer normalises the whole concat buffer
    st2 = torch.zeros(n, 2, cout, dtype=torch.float64, device=dev)           # a second consumer of just this slice
    e = ConvEpilogue()
    e.side, e.side_cstride, e.side_coff, e.add_side = ptr(res), cout + 8, 8, 1
    e.stats1, e.stats1_c, e.stats1_coff = ptr(st1), cout + 16, 16
    e.stats2, e.stats2_c, e.stats2_coff = ptr(st2), cout, 0
    plan.fprop_epi(x, ybuf, bias, e)
    torch.cuda.synchronize()
    got = from_ndhwc(ybuf[..., 16:]).float()
    err = (got - ref).abs()
    assert err.max().item() <= 2e-2 * ref.abs().max().item() and err.mean().item() <= 4e-3 * ref.abs().mean().item()
    assert ybuf[..., :16].abs().max().item() == 0.0
    # statistics of the values AS STORED
    s = ybuf[..., 16:].double().reshape(n, -1, cout)
    want = torch.stack([s.sum(1), (s * s).sum(1)], 1)
    assert torch.allclose(st2, want, rtol=1e-6, atol=1e-6 * want.abs().max().item())
    assert torch.allclose(st1[:, :, 16:], want, rtol=1e-6, atol=1e-6 * want.abs().max().item())
    assert st1[:, :, :16].abs().max().item() == 0.0
    # statistics only (conv1 -> norm2): the stored tensor is bit-identical to the plain kernel's
    y_plain = torch.zeros_like(ybuf)
    plan.fprop(x, y_plain, bias)
    y_stats = torch.zeros_like(ybuf)
    st2.zero_()
    e2 = ConvEpilogue()
    e2.stats1, e2.stats1_c, e2.stats1_coff = ptr(st2), cout, 0
    plan.fprop_epi(x, y_stats, bias, e2)
    torch.cuda.synchronize()
    assert torch.equal(y_plain, y_stats)
    s = y_stats[..., 16:].double().reshape(n, -1, cout)
    want = torch.stack([s.sum(1), (s * s).sum(1)], 1)
    assert torch.allclose(st2, want, rtol=1e-6, atol=1e-6 * want.abs().max().item())
    # in place: the residual sits in the output slice itself (the 1x1 skip convolution wrote it there)
    y_inpl = torch.zeros_like(ybuf)
    y_inpl[..., 16:] = res[..., 8:]
    e3 = ConvEpilogue()
    e3.side, e3.side_cstride, e3.side_coff, e3.add_side = ptr(y_inpl), cout + 16, 16, 1
    plan.fprop_epi(x, y_inpl, bias, e3)
    torch.cuda.synchronize()
    assert torch.equal(y_inpl, ybuf)
    # run to run: the double accumulators make the sums independent of the order the CTAs finish in
    st_a = torch.zeros_like(st2); st_b = torch.zeros_like(st2)
    for st in (st_a, st_b):
        e2.stats1 = ptr(st)
        plan.fprop_epi(x, y_stats, bias, e2)
    torch.cuda.synchronize()
    assert torch.equal(st_a, st_b)


@pytest.mark.parametrize("n,d,h,w,cin,cout", CASES)
@pytest.mark.parametrize("act", ["silu", "none"])
def test_dgrad_with_norm_backward_reduction(petsyn, n, d, h, w, cin, cout, act):
    """dx = dgrad(dy) plus (sum g, sum g * zhat) per (sample, channel), g = dx * act'(z * scale + shift), against
    petsyn_normact_bwd's own reduction pass on the same tensors (and a float64 formula)."""
    from petsyn_b200._cabi import ConvEpilogue, ptr
    ops = petsyn.ops
    if cin not in (16, 32):
        pytest.skip("the data-gradient epilogue covers 16 / 32 input channels")
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(9 + cin + cout)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cout * 27) ** 0.5).to(dev)
    dy = torch.randn(n, d, h, w, cout, generator=g).to(dev).to(torch.bfloat16)
    z = torch.randn(n, d, h, w, cin + 16, generator=g).to(dev).to(torch.bfloat16)      # z = channels [16, 16 + cin) of a buffer
    scale = (torch.rand(n, cin, generator=g) + 0.5).to(dev)
    shift = torch.randn(n, cin, generator=g).to(dev)
    mean = torch.randn(n, cin, generator=g).to(dev) * 0.1
    rstd = (torch.rand(n, cin, generator=g) + 0.5).to(dev)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, 3, 1, 1)
    assert plan.kernel_path[1] == 1 and plan.epi_ok[1]
    plan.pack(wt)
    dx_plain = torch.zeros(n, d, h, w, cin, dtype=torch.bfloat16, device=dev)
    plan.dgrad(dy, dx_plain)
    dx = torch.zeros_like(dx_plain)
    bsums = torch.zeros(n, 2, cin, dtype=torch.float64, device=dev)
    e = ConvEpilogue()
    e.side, e.side_cstride, e.side_coff = ptr(z), cin + 16, 16
    e.norm_scale, e.norm_shift, e.norm_mean, e.norm_rstd = ptr(scale), ptr(shift), ptr(mean), ptr(rstd)
    e.norm_act = ops.ACT_SILU if act == "silu" else ops.ACT_NONE
    e.bsums = ptr(bsums)
    plan.dgrad_epi(dy, dx, e)
    torch.cuda.synchronize()
    assert torch.equal(dx, dx_plain)
    zz = z[..., 16:].double().reshape(n, -1, cin)
    b = zz * scale.double()[:, None] + shift.double()[:, None]
    if act == "silu":
        s = torch.sigmoid(b)
        dact = s * (1 + b * (1 - s))
    else:
        dact = torch.ones_like(b)
    gg = dx.double().reshape(n, -1, cin) * dact
    s0 = gg.sum(1)
    s1 = ((gg * zz).sum(1) - mean.double() * s0) * rstd.double()
    want = torch.stack([s0, s1], 1)
    tol = 2e-4 * want.abs().max().item()          # fp32 partial sums per work item, __expf
    assert (bsums - want).abs().max().item() <= tol, ((bsums - want).abs().max().item(), tol)
    b2 = torch.zeros_like(bsums)
    e.bsums = ptr(b2)
    plan.dgrad_epi(dy, dx, e)
    torch.cuda.synchronize()
    assert torch.equal(b2, bsums)                  # reproducible


# gather-form (igemm) kernel: statistics targets only.  n, d, h, w, cin, cout, k, stride, pad -- ragged tiles, stride 2,
# 1x1x1 (the transformer linears' shape), two channel tiles (cout 256 = 2 x 128)
IGEMM_CASES = [
    (2, 22, 36, 28, 64, 64, 3, 1, 1),
    (2, 20, 34, 30, 64, 128, 3, 1, 1),
    (2, 44, 60, 52, 32, 64, 4, 2, 1),
    (3, 14, 20, 18, 128, 256, 3, 1, 1),
    (2, 18, 32, 32, 96, 32, 1, 1, 0),
    # few voxels: split-K, the statistics come from the finish pass (splitk_finish_stats_kernel)
    (2, 6, 8, 6, 128, 128, 3, 1, 1),
    (3, 5, 7, 6, 64, 256, 3, 1, 1),
    (1, 12, 16, 12, 256, 64, 4, 2, 1),
    (2, 6, 8, 6, 128, 1024, 1, 1, 0),
]


@pytest.mark.parametrize("n,d,h,w,cin,cout,k,s,p", IGEMM_CASES)
def test_gather_form_statistics_epilogue(petsyn, n, d, h, w, cin, cout, k, s, p):
    from petsyn_b200._cabi import ConvEpilogue, ptr
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(11 + cin + cout + k)
    x = torch.randn(n, d, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, k, s, p)
    assert plan.epi_stats_ok and plan.kernel_path[0] == 0 and not plan.epi_ok[0]
    if n * d * h * w // s ** 3 <= 2048:
        assert plan.workspace is not None          # the plan splits K
    plan.pack(wt, need_dgrad=False)
    od, oh, ow = plan.out_dims
    y_plain = torch.zeros(n, od, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    plan.fprop(x, y_plain, bias)
    y = torch.zeros_like(y_plain)
    st1 = torch.zeros(n, 2, cout + 8, dtype=torch.float64, device=dev)
    st2 = torch.zeros(n, 2, cout, dtype=torch.float64, device=dev)
    e = ConvEpilogue()
    e.stats1, e.stats1_c, e.stats1_coff = ptr(st1), cout + 8, 8
    e.stats2, e.stats2_c, e.stats2_coff = ptr(st2), cout, 0
    plan.fprop_epi(x, y, bias, e)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain)
    v = y.double().reshape(n, -1, cout)
    want = torch.stack([v.sum(1), (v * v).sum(1)], 1)
    tol = 1e-6 * want.abs().max().item()
    assert torch.allclose(st2, want, rtol=1e-6, atol=tol)
    assert torch.allclose(st1[:, :, 8:], want, rtol=1e-6, atol=tol) and st1[:, :, :8].abs().max().item() == 0.0
    st3 = torch.zeros_like(st2)
    e2 = ConvEpilogue()
    e2.stats1, e2.stats1_c = ptr(st3), cout
    plan.fprop_epi(x, y, bias, e2)
    torch.cuda.synchronize()
    assert torch.equal(st3, st2)                   # reproducible
    e3 = ConvEpilogue()
    e3.stats1, e3.stats1_c = ptr(st3), cout
    e3.side, e3.side_cstride, e3.add_side = ptr(y_plain), cout, 1
    with pytest.raises(ValueError):
        plan.fprop_epi(x, y, bias, e3)             # this kernel takes statistics targets only
