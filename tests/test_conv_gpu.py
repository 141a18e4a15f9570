"""Kernel-level parity of the tcgen05 implicit-GEMM convolution family (through the C ABI) against
PyTorch fp32 convolutions of the same bf16-rounded operands.

Tolerance: operands are bf16 (8-bit mantissa), accumulation is fp32; the packed weights are rounded to
bf16 (after tap merging for Upsample+Conv).  We require max|err| <= 2e-2 * max|ref| and
mean|err| <= 4e-3 * mean|ref| -- about 4x the bf16 rounding noise of a K~10^3 dot product.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

MAX_REL, MEAN_REL = 2e-2, 4e-3


def to_ndhwc(t):
    return t.permute(0, 2, 3, 4, 1).contiguous()


def from_ndhwc(t):
    return t.permute(0, 4, 1, 2, 3).contiguous()


def close(got, ref, what):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    mx, mean = err.max().item() / (ref.abs().max().item() + 1e-12), err.mean().item() / (ref.abs().mean().item() + 1e-12)
    assert mx <= MAX_REL and mean <= MEAN_REL, f"{what}: max-rel {mx:.3e} mean-rel {mean:.3e}"


def ref_forward(op, x, w, k, s, p):
    if op == "conv":
        return F.conv3d(x, w, None, stride=s, padding=p)
    if op == "upconv":
        return F.conv3d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, stride=1, padding=1)
    return F.conv_transpose3d(x, w, None, stride=2, padding=1)


CASES = [
    # op, n, d, h, w, cin, cout, k, s, p
    ("conv", 1, 8, 12, 16, 64, 128, 3, 1, 1),
    ("conv", 2, 8, 12, 16, 64, 64, 4, 2, 1),
    ("conv", 1, 6, 7, 6, 192, 256, 3, 1, 1),       # ragged boxes, 3 K-chunks, 2 N tiles
    ("conv", 1, 8, 8, 8, 128, 64, 1, 1, 0),
    ("conv", 1, 12, 8, 4, 64, 128, 3, 2, 1),       # BMGAN-style k3 s2
    ("upconv", 1, 4, 6, 8, 128, 64, 3, 1, 1),
    ("upconv", 2, 3, 5, 4, 64, 128, 3, 1, 1),
    ("convt", 1, 4, 6, 8, 64, 64, 4, 2, 1),
    ("convt", 1, 3, 4, 3, 128, 128, 4, 2, 1),
    ("conv", 2, 8, 12, 16, 16, 16, 3, 1, 1),       # AttenUNet full-resolution layers: small-channel wgrad (8 taps per M tile)
    ("conv", 1, 8, 8, 12, 32, 48, 3, 1, 1),        # 4 taps per M tile, N = 48
    ("conv", 1, 8, 8, 8, 48, 16, 3, 1, 1),         # 2 taps per M tile with 32 idle rows
    ("conv", 1, 4, 8, 8, 16, 32, 1, 1, 0),         # 1x1 skip connection
    ("upconv", 1, 4, 6, 8, 32, 32, 3, 1, 1),       # up-sampling ResnetBlock conv1 (8 phases x 8 merged taps)
    ("conv", 2, 3, 4, 3, 128, 128, 3, 2, 1),       # odd extents with stride 2 (ResNet_encoder: 3 -> 2)
    ("conv", 1, 6, 8, 6, 256, 256, 4, 2, 1),       # bottleneck-like: few voxels, long K -> split-K fprop
    ("upconv", 1, 3, 4, 3, 256, 256, 3, 1, 1),     # split-K dgrad (64 taps x 4 chunks, 36 voxels)
    # slab kernels (persistent depth sweep, taps as shifted views of smem-resident halo slabs): need >= 19k / 38k voxels
    ("conv", 2, 16, 32, 48, 16, 16, 3, 1, 1),      # 1 channel atom, exact tiles, 2 samples (depth halo must not leak)
    ("conv", 1, 24, 40, 40, 48, 32, 3, 1, 1),      # 3 atoms in, ragged tiles in w and h
    ("conv", 1, 20, 32, 64, 32, 48, 3, 1, 1),      # N = 48; dgrad runs 48 -> 32
    ("conv", 3, 7, 48, 40, 16, 64, 3, 1, 1),       # short depth, 3 samples, N = 64
    ("conv", 2, 16, 32, 48, 96, 32, 3, 1, 1),      # slab wgrad with two input-channel groups of 3 atoms
    ("conv", 1, 24, 40, 40, 64, 48, 3, 1, 1),      # two groups of 2 atoms; slab fprop with 4 atoms
    ("conv", 2, 12, 32, 48, 80, 16, 3, 1, 1),      # 5 atoms: the second group's last atom is zero-filled by TMA
    ("conv", 2, 16, 32, 48, 48, 16, 1, 1, 0),      # 1x1x1 skip connection on the slab kernels (no halo, one tap)
    ("conv", 2, 16, 32, 48, 64, 64, 3, 1, 1),      # 64 -> 64 with many voxels: gather-form kernel (weights do not fit the slab kernels)
    ("conv", 1, 24, 40, 40, 32, 32, 1, 1, 0),      # ... ragged tiles
    ("conv", 2, 12, 32, 48, 96, 32, 1, 1, 0),      # ... wgrad over two channel groups; fprop / dgrad stay gather-form
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(map(str, c)))
def test_conv_family(case, petsyn):
    op, n, d, h, w, cin, cout, k, s, p = case
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.randn(n, cin, d, h, w, generator=g).to(dev)
    wshape = (cin, cout, k, k, k) if op == "convt" else (cout, cin, k, k, k)
    wt = (torch.randn(*wshape, generator=g) / (cin * k ** 3) ** 0.5).to(dev)
    xb = x.to(torch.bfloat16)
    x32 = xb.float().requires_grad_(True)
    w32 = wt.clone().requires_grad_(True)
    y_ref = ref_forward(op, x32, w32, k, s, p)
    dy = torch.randn(y_ref.shape, generator=g).to(dev).to(torch.bfloat16)
    y_ref.backward(dy.float())

    opcode = {"conv": ops.OP_CONV, "upconv": ops.OP_UPCONV, "convt": ops.OP_CONVT}[op]
    plan = ops.ConvPlan(opcode, n, d, h, w, cin, cout, k, s, p)
    assert plan.out_dims == tuple(y_ref.shape[2:])
    plan.pack(wt)
    x_cl = to_ndhwc(xb)
    y = torch.empty(n, *plan.out_dims, cout, dtype=torch.bfloat16, device=dev)
    plan.fprop(x_cl, y)
    torch.cuda.synchronize()
    close(from_ndhwc(y), y_ref.detach(), f"{op} fprop")

    dx = torch.empty_like(x_cl)
    plan.dgrad(to_ndhwc(dy), dx)
    torch.cuda.synchronize()
    close(from_ndhwc(dx), x32.grad, f"{op} dgrad")

    dw = torch.empty_like(wt)
    plan.wgrad(x_cl, to_ndhwc(dy), dw)
    torch.cuda.synchronize()
    close(dw, w32.grad, f"{op} wgrad")


def test_conv_channel_slices_and_bias(petsyn):
    """x and y as channel slices of wider NDHWC buffers (copy-free concat) + bias + LeakyReLU epilogue."""
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7)
    n, d, h, w, cin, cout = 1, 8, 8, 8, 64, 64
    xbuf = torch.randn(n, d, h, w, 192, generator=g).to(dev).to(torch.bfloat16)
    ybuf = torch.zeros(n, d, h, w, 128, dtype=torch.bfloat16, device=dev)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, 3, 1, 1, x_cstride=192, x_coff=64, y_cstride=128,
                        y_coff=64, act=ops.ACT_LRELU, slope=0.2)
    plan.pack(wt, need_dgrad=False)
    plan.fprop(xbuf, ybuf, bias)
    torch.cuda.synchronize()
    x32 = from_ndhwc(xbuf[..., 64:128]).float()
    ref = F.leaky_relu(F.conv3d(x32, wt, bias, padding=1), 0.2)
    close(from_ndhwc(ybuf[..., 64:]), ref, "sliced fprop")
    assert ybuf[..., :64].abs().max().item() == 0.0   # the other half of the concat buffer is untouched


def test_slab_path_slices_bias_act_accumulate(petsyn):
    """The slab kernels (small-channel k3 s1 p1 convs with many voxels) behind the same contract: x / y / dy / dx as
    channel slices of wider buffers, bias + SiLU epilogue, and the accumulating backward-data variant."""
    from petsyn_b200._cabi import check, lib, ptr, stream_ptr
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(11)
    n, d, h, w, cin, cout = 2, 12, 32, 40, 32, 16
    xbuf = torch.randn(n, d, h, w, 96, generator=g).to(dev).to(torch.bfloat16)
    ybuf = torch.zeros(n, d, h, w, 48, dtype=torch.bfloat16, device=dev)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, 3, 1, 1, x_cstride=96, x_coff=32, y_cstride=48, y_coff=16,
                        dy_cstride=48, dy_coff=32, dx_cstride=96, dx_coff=64, act=ops.ACT_SILU)
    plan.pack(wt)
    plan.fprop(xbuf, ybuf, bias)
    torch.cuda.synchronize()
    x32 = from_ndhwc(xbuf[..., 32:64]).float()
    ref = F.silu(F.conv3d(x32, wt, bias, padding=1))
    close(from_ndhwc(ybuf[..., 16:32]), ref, "slab sliced fprop")
    assert ybuf[..., :16].abs().max().item() == 0.0 and ybuf[..., 32:].abs().max().item() == 0.0

    dybuf = torch.randn(n, d, h, w, 48, generator=g).to(dev).to(torch.bfloat16)
    dxbuf = torch.randn(n, d, h, w, 96, generator=g).to(dev).to(torch.bfloat16)
    before = dxbuf.clone()
    dy32 = from_ndhwc(dybuf[..., 32:48]).float()
    dx_ref = F.conv_transpose3d(dy32, wt, None, padding=1)
    check(lib.petsyn_conv_dgrad_accumulate(plan._h, ptr(dybuf), ptr(plan.w_dgrad), ptr(dxbuf), stream_ptr()), "dgrad_acc")
    torch.cuda.synchronize()
    close(from_ndhwc(dxbuf[..., 64:96]), dx_ref + from_ndhwc(before[..., 64:96]).float(), "slab accumulating dgrad")
    assert torch.equal(dxbuf[..., :64], before[..., :64])
    plan.dgrad(dybuf, dxbuf)
    torch.cuda.synchronize()
    close(from_ndhwc(dxbuf[..., 64:96]), dx_ref, "slab sliced dgrad")

    dw = torch.empty_like(wt)
    plan.wgrad(xbuf, dybuf, dw)
    torch.cuda.synchronize()
    x32r = x32.clone().requires_grad_(False)
    w32 = wt.clone().requires_grad_(True)
    F.conv3d(x32r, w32, None, padding=1).backward(dy32)
    close(dw, w32.grad, "slab sliced wgrad")
    dw2 = dw.clone()
    plan.wgrad(xbuf, dybuf, dw2, accumulate=True)
    torch.cuda.synchronize()
    close(dw2, 2 * w32.grad, "slab accumulating wgrad")


def test_slab_path_fp32_output(petsyn):
    """One-channel network heads keep an fp32 output (cout padded to 16): slab kernel with an fp32 epilogue."""
    ops = petsyn.ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    n, d, h, w, cin, cout = 2, 12, 32, 40, 16, 16
    x = torch.randn(n, d, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, 3, 3, 3, generator=g) / (cin * 27) ** 0.5).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    plan = ops.ConvPlan(ops.OP_CONV, n, d, h, w, cin, cout, 3, 1, 1, y_fp32=True)
    assert plan.kernel_path[0] == 1
    plan.pack(wt, need_dgrad=False)
    y = torch.empty(n, d, h, w, cout, dtype=torch.float32, device=dev)
    plan.fprop(x, y, bias)
    torch.cuda.synchronize()
    ref = F.conv3d(from_ndhwc(x).float(), wt.to(torch.bfloat16).float(), bias, padding=1)
    assert (from_ndhwc(y) - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()      # fp32 epilogue: no bf16 rounding


def test_conv_bad_config_raises(petsyn):
    ops = petsyn.ops
    with pytest.raises(ValueError):
        ops.ConvPlan(ops.OP_CONV, 1, 8, 8, 8, 60, 64, 3, 1, 1)       # cin not a multiple of 8
    with pytest.raises(ValueError):
        ops.ConvPlan(ops.OP_CONV, 1, 8, 8, 8, 64, 64, 3, 1, 3)       # padding >= kernel size
    with pytest.raises(ValueError):
        ops.ConvPlan(ops.OP_CONVT, 1, 8, 8, 8, 64, 64, 3, 1, 1)      # unsupported transposed config
