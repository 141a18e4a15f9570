"""Parity of the bandwidth-bound kernels (norm/act/concat, stem, head, losses, Adam) against plain PyTorch fp32."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(got, ref):
    return ((got.float() - ref.float()).abs().max() / (ref.float().abs().max() + 1e-12)).item()


def test_batchnorm_act_concat_fwd_bwd(petsyn):
    ops = petsyn.ops
    g = torch.Generator().manual_seed(3)
    rows, c = 4096 + 37, 128
    z = (torch.randn(rows, c, generator=g) * 1.7 + 0.3).to(DEV).to(torch.bfloat16)
    gamma = (torch.rand(c, generator=g) + 0.5).to(DEV)
    beta = torch.randn(c, generator=g).to(DEV)
    rm, rv = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    sums = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    scale, shift, mean, rstd = (torch.empty(c, device=DEV) for _ in range(4))
    ops.bn_stats(z, sums, rows, c)
    ops.bn_finalize(sums, gamma, beta, rm, rv, scale, shift, mean, rstd, rows, c, 1e-5, 0.1, True)
    a = torch.empty(rows, c, dtype=torch.bfloat16, device=DEV)
    cat = torch.zeros(rows, 2 * c, dtype=torch.bfloat16, device=DEV)
    ops.norm_act_fwd(z, scale, shift, a, c, 0, ops.ACT_LRELU, cat, 2 * c, c, ops.ACT_RELU, 0.2, rows, c)
    torch.cuda.synchronize()

    z32 = z.float().requires_grad_(True)
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm2, rv2 = torch.zeros(c, device=DEV), torch.ones(c, device=DEV)
    b = F.batch_norm(z32, rm2, rv2, gm, bt, True, 0.1, 1e-5)
    a_ref, s_ref = F.leaky_relu(b, 0.2), F.relu(b)
    assert rel(a, a_ref) < 1e-2 and rel(cat[:, c:], s_ref) < 1e-2
    assert cat[:, :c].abs().max().item() == 0
    assert rel(rm, rm2) < 1e-4 and rel(rv, rv2) < 1e-4

    g1 = torch.randn(rows, c, generator=g).to(DEV).to(torch.bfloat16)
    g2buf = torch.randn(rows, 2 * c, generator=g).to(DEV).to(torch.bfloat16)
    (a_ref * g1.float()).sum().add((s_ref * g2buf[:, c:].float()).sum()).backward()
    bsums = torch.zeros(2 * c, device=DEV, dtype=torch.float64)
    dz = torch.empty(rows, c, dtype=torch.bfloat16, device=DEV)
    dgamma, dbeta = torch.empty(c, device=DEV), torch.empty(c, device=DEV)
    ops.norm_act_bwd(z, scale, shift, mean, rstd, gamma, g1, c, 0, ops.ACT_LRELU, g2buf, 2 * c, c, ops.ACT_RELU, 0.2,
                     bsums, dz, dgamma, dbeta, rows, c)
    torch.cuda.synchronize()
    assert rel(dz, z32.grad) < 2e-2
    assert rel(dgamma, gm.grad) < 2e-3 and rel(dbeta, bt.grad) < 2e-3


def test_stem_and_head(petsyn):
    ops = petsyn.ops
    g = torch.Generator().manual_seed(5)
    n, d, h, w, c = 2, 8, 12, 16, 64
    x = torch.rand(n, 1, d, h, w, generator=g).to(DEV)
    ws = (torch.randn(c, 1, 4, 4, 4, generator=g) / 8).to(DEV)
    stem = ops.StemConv(n, d, h, w, c, DEV)
    stem.pack(ws)
    y = torch.empty(n, d // 2, h // 2, w // 2, c, dtype=torch.bfloat16, device=DEV)
    stem.fprop(x, y)
    ref = F.conv3d(x, ws, None, stride=2, padding=1)
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 1.5e-2
    dy = torch.randn(ref.shape, generator=g).to(DEV).to(torch.bfloat16)
    ws2 = ws.clone().requires_grad_(True)
    F.conv3d(x, ws2, None, stride=2, padding=1).backward(dy.float())
    dw = torch.empty_like(ws)
    stem.wgrad(dy.permute(0, 2, 3, 4, 1).contiguous(), dw)
    assert rel(dw, ws2.grad) < 1e-2

    # head: relu'd input (bf16, NDHWC) -> up x2 -> conv3(C->1) -> tanh
    ch = 2 * c
    xin = torch.randn(n, d, h, w, ch, generator=g).relu().to(DEV).to(torch.bfloat16)
    wh = (torch.randn(1, ch, 3, 3, 3, generator=g) / (ch * 27) ** 0.5 * 3).to(DEV)
    head = ops.HeadConv(n, d, h, w, ch, DEV)
    head.pack(wh)
    yo = torch.empty(n, 1, 2 * d, 2 * h, 2 * w, device=DEV)
    head.fprop(xin, yo)
    x32 = xin.float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    wh2 = wh.clone().requires_grad_(True)
    yr = torch.tanh(F.conv3d(F.interpolate(x32, scale_factor=2, mode="nearest"), wh2, None, padding=1))
    assert (yo - yr).abs().max().item() < 2e-2          # weights are bf16 inside the GEMM
    go = torch.randn(yr.shape, generator=g).to(DEV)
    yr.backward(go)
    dx = torch.empty_like(xin)
    dwh = torch.empty_like(wh)
    head.backward(xin, yo, go, dx, dwh)
    assert rel(dx.permute(0, 4, 1, 2, 3), x32.grad) < 2e-2
    assert rel(dwh, wh2.grad) < 1e-2


def test_losses_and_adam(petsyn):
    ops = petsyn.ops
    g = torch.Generator().manual_seed(9)
    y = torch.rand(3, 1, 8, 9, 11, generator=g).to(DEV)
    t = torch.rand(3, 1, 8, 9, 11, generator=g).to(DEV)
    loss = torch.zeros(1, device=DEV)
    dy = torch.empty_like(y)
    ops.l1_loss_fwd_bwd(y, t, loss, dy)
    y2 = y.clone().requires_grad_(True)
    lr = F.l1_loss(y2, t)
    lr.backward()
    assert abs(loss.item() - lr.item()) < 1e-6 and (dy - y2.grad).abs().max().item() < 1e-9

    x = torch.randn(2, 1, 4, 6, 4, generator=g).to(DEV)
    loss.zero_()
    dx = torch.empty_like(x)
    ops.mse_const_fwd_bwd(x, 1.0, loss, dx)
    x2 = x.clone().requires_grad_(True)
    lm = F.mse_loss(x2, torch.ones_like(x2))
    lm.backward()
    assert abs(loss.item() - lm.item()) < 1e-5 and (dx - x2.grad).abs().max().item() < 1e-7

    p = torch.randn(10007, generator=g).to(DEV)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=5e-4, betas=(0.9, 0.999), eps=1e-8)
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    p2, m2, v2 = p.clone(), torch.zeros_like(p), torch.zeros_like(p)
    for step in range(1, 4):
        gr = torch.randn(10007, generator=g).to(DEV)
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step(p, gr, m, v, 5e-4, 0.9, 0.999, 1e-8, step)
        ops.adam_step(p2, gr, m2, v2, 5e-4, 0.9, 0.999, 1e-8, 0, step_dev=torch.tensor([step], dtype=torch.int32, device=DEV))
    assert (p - pr.detach()).abs().max().item() < 1e-6
    assert torch.equal(p, p2)
    out = torch.zeros(1, device=DEV)
    ops.sumsq(p, out)
    assert abs(out.item() - (p.double() ** 2).sum().item()) / out.item() < 1e-5


def test_generalised_normact(petsyn):
    """Descriptor API: InstanceNorm (per-sample statistics) + PReLU with a device-resident slope + residual add, forward
    and backward (dz, d(res), d(slope)) against torch fp32 autograd on the same bf16-rounded tensors."""
    from petsyn_b200 import graph
    ops = petsyn.ops
    g = torch.Generator().manual_seed(0)
    n, d, h, w, c = 2, 5, 6, 5, 128
    z, out, res = (graph.Buf(n, d, h, w, c, DEV, nm) for nm in "zor")
    z.t.copy_((torch.randn(z.rows, c, generator=g) * 2 + 0.5).to(DEV))
    res.t.copy_(torch.randn(z.rows, c, generator=g).to(DEV))
    alpha = torch.nn.Parameter(torch.tensor([0.25], device=DEV))
    op = graph.NormActOp(z, "instance", ops.ACT_PRELU, [out.sl()], res=res.sl(), slope_param=alpha)
    op.grad_slope = torch.zeros(1, device=DEV)
    op.fwd(True)
    out.g.copy_(torch.randn(z.rows, c, generator=g).to(DEV))
    op.bwd()
    torch.cuda.synchronize()
    ncl = lambda t: t.float().view(n, d * h * w, c).permute(0, 2, 1).contiguous()
    zz, rr = ncl(z.t).requires_grad_(True), ncl(res.t).requires_grad_(True)
    a2 = alpha.detach().clone().requires_grad_(True)
    o = F.prelu(F.instance_norm(zz, eps=1e-5), a2) + rr
    o.backward(ncl(out.g))
    assert rel(ncl(out.t), o) < 1e-2
    assert rel(ncl(z.g), zz.grad) < 1e-2
    assert (ncl(res.g) - rr.grad).abs().max().item() == 0.0
    assert abs(op.grad_slope.item() - a2.grad.item()) <= 1e-4 * abs(a2.grad.item())


@pytest.mark.parametrize("n,L,H", [(2, 96, 4), (1, 200, 2), (2, 2304, 4), (1, 77, 1)])
def test_flash_attention_matches_torch(n, L, H, petsyn):
    """Tensor-core flash attention (csrc/attention_mma.cu) through the C ABI against fp32 softmax(q k^T * scale) v of the
    same bf16-rounded q, k, v (atten_unet_model.py:137-154), forward and backward; L not a multiple of the 64-row tile
    exercises the masking of padded keys / queries.  P and dS are rounded to bf16 for the second GEMM of each product:
    tolerance 2e-2 of the reference's max (forward), 3e-2 (gradients)."""
    from petsyn_b200._cabi import check, lib, ptr, stream_ptr
    dev = torch.device("cuda:0")
    hd = 32
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(n * L, 3 * H * hd, generator=g).to(dev).to(torch.bfloat16)
    dout = torch.randn(n * L, H * hd, generator=g).to(dev).to(torch.bfloat16)
    scale = hd ** -0.5
    out = torch.empty(n * L, H * hd, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(n, H, L, dtype=torch.float32, device=dev)
    delta = torch.empty_like(lse)
    dqkv = torch.zeros_like(qkv)
    check(lib.petsyn_attention_fwd(ptr(qkv), ptr(out), ptr(lse), n, L, H, hd, scale, stream_ptr()), "attention_fwd")
    check(lib.petsyn_attention_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(delta), ptr(dqkv), n, L, H, hd, scale,
                                   stream_ptr()), "attention_bwd")
    torch.cuda.synchronize()

    x = qkv.float().view(n, L, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)   # [3, n, H, L, hd]
    q, k, v = x[0], x[1], x[2]
    s = torch.einsum("nhld,nhmd->nhlm", q, k) * scale
    ref = torch.einsum("nhlm,nhmd->nhld", torch.softmax(s, -1), v)                                  # [n, H, L, hd]
    ref.backward(dout.float().view(n, L, H, hd).permute(0, 2, 1, 3))
    ref_out = ref.detach().permute(0, 2, 1, 3).reshape(n * L, H * hd)
    ref_lse = torch.logsumexp(s.detach(), -1)
    ref_dqkv = x.grad.permute(1, 3, 0, 2, 4).reshape(n * L, 3 * H * hd)

    def rel(a, b):
        return ((a.float() - b).abs().max() / b.abs().max()).item()

    assert rel(out, ref_out) <= 2e-2, rel(out, ref_out)
    assert (lse - ref_lse).abs().max().item() <= 2e-3
    for i, name in enumerate("qkv"):
        sl = slice(i * H * hd, (i + 1) * H * hd)
        assert rel(dqkv[:, sl], ref_dqkv[:, sl]) <= 3e-2, (name, rel(dqkv[:, sl], ref_dqkv[:, sl]))


@pytest.mark.parametrize("shape", [(2, 1, 20, 24, 40), (1, 1, 9, 13, 37), (1, 1, 32, 48, 32)])
def test_ssim_loss_matches_oracle(shape, petsyn):
    """Single-scale SSIM (Gaussian 5-tap window, sigma 0.5, data_range 1: output_predict.py:73) and the gradient of the
    loss 1 - mean(SSIM) against the float64 CPU oracle (oracle/ssim.py, autograd); MAE / PSNR against their definitions.
    fp32 arithmetic: 1e-5 on the value, 1e-4 of the gradient's max."""
    from oracle import ssim as OS
    ops = petsyn.ops
    g = torch.Generator().manual_seed(9)
    x = torch.rand(shape, generator=g)
    y = (x + 0.2 * torch.randn(shape, generator=g)).clamp(0, 1)
    xd = x.double().requires_grad_(True)
    loss_o = OS.ssim_loss(xd, y.double())
    loss_o.backward()
    xc, yc = x.cuda(), y.cuda()
    dx = torch.empty_like(xc)
    crit = ops.SsimLoss(shape, xc.device)
    mean_ssim = crit(xc, yc, dx, grad_scale=1.0)
    torch.cuda.synchronize()
    assert abs((1.0 - mean_ssim.item()) - loss_o.item()) <= 1e-5
    gref = xd.grad.float()
    assert (dx.cpu() - gref).abs().max().item() <= 1e-4 * gref.abs().max().item()
    assert abs(crit(xc, yc).item() - mean_ssim.item()) <= 1e-6          # evaluation-only call (no gradient, no workspace)
    mae, psnr = ops.eval_metrics(xc, yc)
    assert abs(mae.item() - (x - y).abs().mean().item()) <= 1e-6
    assert abs(psnr.item() - (10 * torch.log10(1.0 / ((x - y) ** 2).mean())).item()) <= 1e-3


def test_ms_ssim_matches_oracle(petsyn):
    """The reference's evaluation metric (output_predict.py:73,126: MS-SSIM, kernel 5, sigma 0.5, data_range 1) composed from
    the SSIM kernel's per-sample contrast-structure sums and 2x average pooling, against the float64 oracle."""
    from oracle import ssim as OS
    g = torch.Generator().manual_seed(21)
    shape = (2, 1, 80, 96, 88)
    x = torch.rand(shape, generator=g)
    y = (x + 0.1 * torch.randn(shape, generator=g)).clamp(0, 1)
    ref = OS.ms_ssim(x.double(), y.double()).item()
    got = petsyn.ops.ms_ssim(x.cuda(), y.cuda()).item()
    assert abs(got - ref) <= 1e-4, (got, ref)
    with pytest.raises(ValueError):
        petsyn.ops.ms_ssim(x[..., :40].cuda(), y[..., :40].cuda())
