"""Drop-in replacements for the BMGAN baseline networks of the reference (``bl_methods/BMGAN/bmgan_model.py``):
``dense_unet_generator`` (:25-101) and ``patch_discriminator`` (:133-144), running on libpetsyn's sm_100a kernels.

Kept identical to the reference: constructor signatures/defaults, ``forward`` signatures and return values, the module
tree (including the names MONAI's ``ConvDenseBlock`` / ``ResidualUnit`` / ``Convolution`` / ``ADN`` and
GenerativeModels' ``PatchDiscriminator`` give their children) and therefore every ``state_dict`` key and shape, the
seeded default initialisation, fp32 NCDHW tensors at the boundary, autograd differentiability.
Nothing is computed by the container modules: ``forward`` hands the network to a static op tape (``graph.py``) of
tcgen05 implicit-GEMM convolutions (Conv3d k3 s1/s2, k1, ConvTranspose3d k4 s2 p1) and fused
InstanceNorm/BatchNorm + LeakyReLU + residual + dense-concat kernels over channels-last bf16 buffers.

Spatial dims must be divisible by 32 (the reference's own constraint: five stride-2 stages with skip concatenation).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import graph, ops
from ._cabi import check, lib, ptr, stream_ptr
from .graph import Buf, ConvOp, NormActOp, Sl, Tape

LRELU_SLOPE = 0.2


# ------------------------------------------------------------------------------------------------ containers
class _Container(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("petsyn BMGAN blocks are parameter containers; call the top-level network")


class _ADN(nn.Sequential):
    """Children named as MONAI's ADN names them: N (norm), A (activation)."""

    def __init__(self, channels: int, norm: Optional[str], act: bool = True):
        super().__init__()
        if norm == "instance":
            self.add_module("N", nn.InstanceNorm3d(channels))
        elif norm == "batch":
            self.add_module("N", nn.BatchNorm3d(channels))
        if act:
            self.add_module("A", nn.LeakyReLU(LRELU_SLOPE))


class _Convolution(nn.Sequential):
    """MONAI ``Convolution``: child ``conv`` (+ ``adn``)."""

    def __init__(self, cin, cout, k, stride, pad, bias=True, norm: Optional[str] = "instance", act=True, conv_only=False):
        super().__init__()
        self.add_module("conv", nn.Conv3d(cin, cout, k, stride=stride, padding=pad, bias=bias))
        if not conv_only:
            self.add_module("adn", _ADN(cout, norm, act))


class _ResidualUnit(_Container):
    """MONAI ``ResidualUnit`` with one sub-unit: ``conv.unit0`` and ``residual`` (1x1 conv when channels change)."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv = nn.Sequential()
        self.conv.add_module("unit0", _Convolution(cin, cout, 3, 1, 1))
        self.residual = nn.Conv3d(cin, cout, 1, 1, 0, bias=True) if cin != cout else nn.Identity()


class _ConvDenseBlock(nn.Sequential):
    """MONAI ``ConvDenseBlock(channels=[c], num_res_units=1)``: child ``layers0``."""

    def __init__(self, cin, c):
        super().__init__()
        self.add_module("layers0", _ResidualUnit(cin, c))


def _dense_block(cin: int, c: int) -> List[nn.Module]:
    mods: List[nn.Module] = []
    for a in (cin, c):
        mods += [_ConvDenseBlock(a, c), nn.Conv3d(a + c, c, 3, padding=1), nn.InstanceNorm3d(c), nn.LeakyReLU(LRELU_SLOPE)]
    return mods


def _cnl(cin, cout, stride=1) -> List[nn.Module]:
    return [nn.Conv3d(cin, cout, 3, padding=1, stride=stride), nn.InstanceNorm3d(cout), nn.LeakyReLU(LRELU_SLOPE)]


# ------------------------------------------------------------------------------------------------ generator
class dense_unet_generator(nn.Module):
    """B200-native ``dense_unet_generator`` (bmgan_model.py:25-101): same ctor, same keys, same forward contract."""

    def __init__(self, input_channel=9, input_conv_channel=64, output_conv_channel=64,
                 down_layers=5, down_channels=[128, 256, 256, 512],
                 middle_layers=1, middle_channels=[512],
                 up_layers=6, up_channels=[512, 256, 256, 256, 128]):
        super().__init__()
        if len(up_channels) != len(down_channels) + 1:
            raise ValueError("up_channels must have one more entry than down_channels (one per skip connection)")
        ic, oc = input_conv_channel, output_conv_channel
        self.cfg = dict(input_channel=input_channel, ic=ic, oc=oc, down=list(down_channels), mid=middle_channels[-1],
                        up=list(up_channels))
        self.input_layer = nn.Sequential(*(_cnl(input_channel, ic) + _cnl(ic, ic) + _cnl(ic, ic, 2)))
        self.down_layers = nn.ModuleList([])
        cur = ic
        for c in down_channels:
            self.down_layers.append(nn.Sequential(*(_dense_block(cur, c) + _cnl(c, c, 2))))
            cur = c
        self.middle_layers = nn.Sequential(*_dense_block(cur, middle_channels[-1]))
        cur = middle_channels[-1]
        skips = [ic] + list(down_channels)
        self.up_layers = nn.ModuleList([])
        for i, c in enumerate(up_channels):
            self.up_layers.append(nn.Sequential(*(_dense_block(cur + skips[-1 - i], c) + [
                nn.ConvTranspose3d(c, c, kernel_size=4, stride=2, padding=1), nn.InstanceNorm3d(c),
                nn.LeakyReLU(LRELU_SLOPE)])))
            cur = c
        self.output_layer = nn.Sequential(*(_cnl(cur, oc) + _cnl(oc, oc) + [nn.Conv3d(oc, 1, 3, padding=1), nn.Tanh()]))
        self._engines: Dict[Tuple, "_GenEngine"] = ops.EngineCache()

    def engine_for(self, x: torch.Tensor) -> "_GenEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _GenEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def forward(self, x: torch.Tensor, sampled_latent_vector: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("petsyn dense_unet_generator runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        z = sampled_latent_vector.reshape(x.shape[0], -1).contiguous().float()
        if 1 + z.shape[1] != self.cfg["input_channel"]:
            raise ValueError(f"latent vector has {z.shape[1]} entries, expected {self.cfg['input_channel'] - 1}")
        eng = self.engine_for(x)
        params = eng.params
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _TapeFn.apply(eng, x, z, *params)
        return eng.forward(x, z).clone()        # the engine owns its output buffer: hand out a copy (no aliasing across calls)


class _TapeFn(torch.autograd.Function):
    """Autograd bridge: forward runs the tape, backward runs it in reverse and returns input + parameter gradients."""

    @staticmethod
    def forward(ctx, eng, x, extra, *params):
        ctx.need_dx = x.requires_grad
        y = eng.forward(x, extra).clone()
        eng.stamp(ctx, (x, extra))
        return y

    @staticmethod
    def backward(ctx, dy):
        ctx.eng.restore(ctx)
        dx, grads = ctx.eng.backward(dy.contiguous().float(), need_dx=ctx.need_dx)
        return (None, dx, None, *grads)


class _EngineBase:
    """Shared plumbing: parameter list, gradient slots, training-mode lookup."""

    def __init__(self, module: nn.Module, dev):
        self.module = module
        self.dev = dev
        self.tape = Tape()
        self.params: List[nn.Parameter] = []
        self._slots: Optional[List[torch.Tensor]] = None
        self._bind: List[Tuple[object, str, nn.Parameter]] = []     # (op, attribute, parameter)
        # An engine holds ONE set of saved activations.  Every forward bumps `generation`; an autograd node remembers the
        # generation it ran in and its inputs, and `restore` re-runs the forward when a later call has overwritten them.
        self.generation = 0
        self._training_override: Optional[bool] = None

    # ------------------------------------------------------------------------------------------------ re-entrancy
    def training(self) -> bool:
        return self.module.training if self._training_override is None else self._training_override

    def stamp(self, ctx, inputs) -> None:
        """Called by the autograd bridge right after its forward: remember what is needed to redo that forward."""
        ctx.eng, ctx.gen, ctx.mode = self, self.generation, self.training()
        ctx.versions = tuple(p._version for p in self.params)
        ctx.save_for_backward(*inputs)

    def restore(self, ctx) -> None:
        """Called by the autograd bridge before its backward.  Two forwards followed by one backward (the encoder in
        train_bmgan.py:170-180; D(fake) + D(real) summed) leave the activations of the LAST call in the engine: the
        earlier node re-runs its forward from the saved inputs first (same mode; BatchNorm running statistics are not
        moved a second time), so each node backpropagates through its own activations."""
        if self.generation == ctx.gen:
            return
        if tuple(p._version for p in self.params) != ctx.versions:
            raise RuntimeError("a parameter of this petsyn module was modified in place between forward and backward")
        self._training_override = ctx.mode
        graph.RunState.freeze_running_stats = True
        try:
            self.forward(*ctx.saved_tensors)
        finally:
            graph.RunState.freeze_running_stats = False
            self._training_override = None
        ctx.gen = self.generation

    def _conv(self, x: Sl, conv: nn.Module, **kw) -> ConvOp:
        op = ConvOp(x, conv.weight, conv.bias, **kw)
        self.tape.add(op)
        self._bind.append((op, "grad_w", conv.weight))
        if conv.bias is not None:
            self._bind.append((op, "grad_b", conv.bias))
        return op

    def _finish(self) -> None:
        seen = set()
        for _, _, p in self._bind:
            if id(p) not in seen:
                seen.add(id(p))
                self.params.append(p)
        self.tape.finalize()
        self.flops_algorithmic = self.tape.flops()

    def grad_order(self) -> List[nn.Parameter]:
        """Parameters in the order backward produces their gradients (reverse op order), then any parameter no op
        produces a gradient for."""
        order, seen = [], set()
        for op in reversed(self.tape.ops):
            for o, _, p in self._bind:
                if o is op and id(p) not in seen:
                    seen.add(id(p))
                    order.append(p)
        for p in self.params:
            if id(p) not in seen:
                seen.add(id(p))
                order.append(p)
        return order

    def run_backward(self, on_ready=None) -> None:
        """Tape backward; ``on_ready(param)`` fires right after the op producing that parameter's gradient."""
        cb = None
        if on_ready is not None:
            by_op: Dict[int, List[nn.Parameter]] = {}
            for op, _, p in self._bind:
                by_op.setdefault(id(op), []).append(p)
            cb = lambda op: [on_ready(p) for p in by_op.get(id(op), [])]
        self.tape.backward(cb)

    def mark_weights_dirty(self) -> None:
        for op in self.tape.ops:
            if hasattr(op, "_ver"):
                op._ver = None

    def grad_slots(self, out: Optional[Dict[int, torch.Tensor]] = None) -> List[torch.Tensor]:
        """Bind every op's gradient destination; ``out`` maps id(param) -> tensor (flat-arena views), else the engine
        owns the buffers."""
        if out is None:
            if self._slots is None:
                self._slots = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
            out = {id(p): g for p, g in zip(self.params, self._slots)}
        for op, attr, p in self._bind:
            setattr(op, attr, out[id(p)])
        return [out[id(p)] for p in self.params]


class _GenEngine(_EngineBase):
    """Op tape of the dense U-Net generator for one (input shape, device)."""

    CPAD_IN = 16

    def __init__(self, gen: dense_unet_generator, shape, dev):
        super().__init__(gen, dev)
        n, _, D, H, W = shape
        cfg = gen.cfg
        nstage = len(cfg["down"]) + 1
        if D % (1 << nstage) or H % (1 << nstage) or W % (1 << nstage):
            raise ValueError(f"spatial dims {D}x{H}x{W} must be divisible by {1 << nstage}")
        self.shape = shape
        ic, oc, down, mid, up = cfg["ic"], cfg["oc"], cfg["down"], cfg["mid"], cfg["up"]
        for c in [ic, oc, mid] + down + up:
            if c % 8:
                raise ValueError(f"channel widths must be multiples of 8 (got {c})")
        res = lambda i: (D >> i, H >> i, W >> i)
        self.bufs: Dict[str, Buf] = {}

        def B(i, c, name):
            self.bufs[name] = Buf(n, *res(i), c, dev, name)
            return self.bufs[name]

        skips = [ic] + down
        nd = len(down)
        # ---- concat buffers of every dense block: cat0 = [X | v0], cat1 = [y0 | v1] ----
        d_cat0 = [B(i + 1, (skips[i]) + down[i], f"down{i}.cat0") for i in range(nd)]
        d_cat1 = [B(i + 1, 2 * down[i], f"down{i}.cat1") for i in range(nd)]
        d_tail = [B(i + 1, down[i], f"down{i}.y1") for i in range(nd)]
        m_cat0 = B(nd + 1, down[-1] + mid, "mid.cat0")
        m_cat1 = B(nd + 1, 2 * mid, "mid.cat1")
        cur = mid
        u_cat0, u_cat1, u_tail, u_in = [], [], [], []
        for j, c in enumerate(up):
            lvl = nd + 1 - j
            cin = cur + skips[-1 - j]
            u_in.append((cur, skips[-1 - j]))
            u_cat0.append(B(lvl, cin + c, f"up{j}.cat0"))
            u_cat1.append(B(lvl, 2 * c, f"up{j}.cat1"))
            u_tail.append(B(lvl, c, f"up{j}.y1"))
            cur = c
        out_in = B(0, up[-1], "out.in")
        # ---- input layer ----
        self.inp = B(0, self.CPAD_IN, "input")
        il, t = gen.input_layer, self.tape
        a0, a1 = B(0, ic, "in.a0"), B(0, ic, "in.a1")
        c0 = self._conv(self.inp.sl(), il[0], ksize=3, stride=1, pad=1, use_bias=False, need_dx=False, name="in.conv0")
        t.add(NormActOp(c0.z, "instance", ops.ACT_LRELU, [a0.sl()]))
        c1 = self._conv(a0.sl(), il[3], ksize=3, stride=1, pad=1, use_bias=False, name="in.conv1")
        t.add(NormActOp(c1.z, "instance", ops.ACT_LRELU, [a1.sl()]))
        c2 = self._conv(a1.sl(), il[6], ksize=3, stride=2, pad=1, use_bias=False, name="in.conv2")
        # F0 feeds the first dense block (as X) and the last up stage (as skip)
        t.add(NormActOp(c2.z, "instance", ops.ACT_LRELU, [d_cat0[0].sl(0, ic), self._skip_slot(u_cat0, u_in, nd)]))
        # ---- down stages ----
        for i in range(nd):
            self._dense(gen.down_layers[i], d_cat0[i], d_cat1[i], skips[i], down[i], [d_tail[i].sl()], f"down{i}")
            tail = self._conv(d_tail[i].sl(), gen.down_layers[i][8], ksize=3, stride=2, pad=1, use_bias=False,
                              name=f"down{i}.tail")
            nxt = d_cat0[i + 1].sl(0, down[i]) if i + 1 < nd else m_cat0.sl(0, down[i])
            t.add(NormActOp(tail.z, "instance", ops.ACT_LRELU, [nxt, self._skip_slot(u_cat0, u_in, nd - 1 - i)]))
        # ---- middle ----
        self._dense(gen.middle_layers, m_cat0, m_cat1, down[-1], mid, [u_cat0[0].sl(0, mid)], "mid")
        # ---- up stages ----
        for j, c in enumerate(up):
            cin = u_in[j][0] + u_in[j][1]
            self._dense(gen.up_layers[j], u_cat0[j], u_cat1[j], cin, c, [u_tail[j].sl()], f"up{j}")
            tail = self._conv(u_tail[j].sl(), gen.up_layers[j][8], ksize=4, stride=2, pad=1, op=ops.OP_CONVT,
                              use_bias=False, name=f"up{j}.tail")
            nxt = u_cat0[j + 1].sl(0, c) if j + 1 < len(up) else out_in.sl()
            t.add(NormActOp(tail.z, "instance", ops.ACT_LRELU, [nxt]))
        # ---- output layer ----
        ol = gen.output_layer
        o0, o1 = B(0, oc, "out.a0"), B(0, oc, "out.a1")
        c0 = self._conv(out_in.sl(), ol[0], ksize=3, stride=1, pad=1, use_bias=False, name="out.conv0")
        t.add(NormActOp(c0.z, "instance", ops.ACT_LRELU, [o0.sl()]))
        c1 = self._conv(o0.sl(), ol[3], ksize=3, stride=1, pad=1, use_bias=False, name="out.conv1")
        t.add(NormActOp(c1.z, "instance", ops.ACT_LRELU, [o1.sl()]))
        self.head = self._conv(o1.sl(), ol[6], ksize=3, stride=1, pad=1, act=ops.ACT_TANH, y_fp32=True, name="out.head")
        self.y = torch.zeros(n, 1, D, H, W, dtype=torch.float32, device=dev)
        self._finish()

    @staticmethod
    def _skip_slot(u_cat0, u_in, j) -> Sl:
        """Channel slice of up stage j's concat buffer that holds its skip tensor ([feature | skip | v0])."""
        cur, sk = u_in[j]
        return u_cat0[j].sl(cur, sk)

    def _dense(self, blk: nn.Sequential, cat0: Buf, cat1: Buf, cin: int, c: int, dsts: Sequence[Sl], name: str) -> None:
        """get_dense_block (bmgan_model.py:12-23) on pre-allocated concat buffers; X already lives in cat0[:, :cin]."""
        t = self.tape
        for idx, (cat, a) in enumerate(((cat0, cin), (cat1, c))):
            ru: _ResidualUnit = blk[4 * idx].layers0
            x = cat.sl(0, a)
            cu = self._conv(x, ru.conv.unit0.conv, ksize=3, stride=1, pad=1, use_bias=False, name=f"{name}.ru{idx}.conv")
            if isinstance(ru.residual, nn.Conv3d):
                cr = self._conv(x, ru.residual, ksize=1, stride=1, pad=0, use_bias=True, name=f"{name}.ru{idx}.res")
                res = cr.z.sl()
            else:
                res = x
            t.add(NormActOp(cu.z, "instance", ops.ACT_LRELU, [cat.sl(a, c)], res=res))
            cy = self._conv(cat.sl(), blk[4 * idx + 1], ksize=3, stride=1, pad=1, use_bias=False, name=f"{name}.conv{idx}")
            t.add(NormActOp(cy.z, "instance", ops.ACT_LRELU, [cat1.sl(0, c)] if idx == 0 else list(dsts)))

    # ------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, z: torch.Tensor) -> torch.Tensor:
        n, _, D, H, W = self.shape
        self.generation += 1
        check(lib.petsyn_concat_latent(ptr(x), ptr(z), ptr(self.inp.t), D * H * W, n, z.shape[1], self.CPAD_IN,
                                       stream_ptr()), "concat_latent")
        self.tape.forward(self.training())
        check(lib.petsyn_take_channel0(ptr(self.head.zf), ptr(self.y), self.y.numel(), self.head.cout, stream_ptr()),
              "take_channel0")
        return self.y

    def backward(self, dy: torch.Tensor, need_dx: bool = False, out: Optional[Dict[int, torch.Tensor]] = None,
                 on_ready=None):
        grads = self.grad_slots(out)
        check(lib.petsyn_put_channel0_grad(ptr(self.y), ptr(dy), ptr(self.head.zg), dy.numel(), self.head.cout, 1,
                                           stream_ptr()), "put_channel0_grad")
        cb = None
        if on_ready is not None:
            by_op: Dict[int, List[nn.Parameter]] = {}
            for op, _, p in self._bind:
                by_op.setdefault(id(op), []).append(p)
            cb = lambda op: [on_ready(p) for p in by_op.get(id(op), [])]
        self.tape.backward(cb)
        return None, ([g.clone() for g in grads] if out is None else [])

    def grad_order(self) -> List[nn.Parameter]:
        """Parameters in the order backward produces their gradients (reverse op order)."""
        order, seen = [], set()
        for op in reversed(self.tape.ops):
            for o, _, p in self._bind:
                if o is op and id(p) not in seen:
                    seen.add(id(p))
                    order.append(p)
        return order


# ------------------------------------------------------------------------------------------------ discriminator
class _PatchDiscriminatorNet(nn.Sequential):
    """Parameter container with GenerativeModels' ``PatchDiscriminator`` child names and initialisation."""

    def __init__(self, num_channels: int, in_channels: int, num_layers_d: int):
        super().__init__()
        self.add_module("initial_conv", _Convolution(in_channels, num_channels, 4, 2, 1, bias=True, norm=None))
        cin, cout = num_channels, num_channels * 2
        for l_ in range(num_layers_d):
            self.add_module(str(l_), _Convolution(cin, cout, 4, 1 if l_ == num_layers_d - 1 else 2, 1, bias=False,
                                                  norm="batch"))
            cin, cout = cout, cout * 2
        self.add_module("final_conv", _Convolution(cin, 1, 4, 1, 1, bias=True, conv_only=True))
        self.num_layers_d = num_layers_d
        self.apply(self._init)

    @staticmethod
    def _init(m: nn.Module) -> None:
        if isinstance(m, nn.Conv3d):
            nn.init.normal_(m.weight.data, 0.0, 0.02)
        elif isinstance(m, nn.BatchNorm3d):
            nn.init.normal_(m.weight.data, 1.0, 0.02)
            nn.init.constant_(m.bias.data, 0)

    def forward(self, *a, **k):
        raise RuntimeError("petsyn PatchDiscriminator is a parameter container; call patch_discriminator")


class _LastStageOnly(tuple):
    """What ``PatchDiscriminator.forward`` returns: the reference returns the list of ALL stage outputs and every script reads
    ``[-1]`` only (train_unet.py:154,179,182; train_unify_causal_gen.py:231,262,265).  Here only that last entry exists;
    asking for an intermediate feature map raises instead of handing out something else."""

    def __new__(cls, logits, n):
        self = super().__new__(cls, (None,) * (n - 1) + (logits,))
        return self

    def __getitem__(self, i):
        v = super().__getitem__(i)
        if v is None:
            raise NotImplementedError("petsyn PatchDiscriminator exposes the last stage (the patch logits, index -1) only")
        return v


class PatchDiscriminator(_PatchDiscriminatorNet):
    """B200-native drop-in for ``monai_diffusion.generative.networks.nets.PatchDiscriminator`` as the unet / causal scripts
    build it: ``PatchDiscriminator(**model_dict['discriminator'])`` (train_unet.py:74, unet/config/training.json:40-46: 64
    channels x 3 layers; training_causal.json:76-82).  Same constructor signature, child names (``initial_conv``, ``"0"`` ..,
    ``final_conv``: state-dict keys of a reference checkpoint load unchanged) and initialisation; ``forward`` returns a
    sequence whose ``[-1]`` is the patch-logit tensor."""

    def __init__(self, spatial_dims: int, num_channels: int, in_channels: int, out_channels: int = 1, num_layers_d: int = 3,
                 kernel_size: int = 4, activation=("LEAKYRELU", {"negative_slope": 0.2}), norm="BATCH", bias: bool = False,
                 padding=1, dropout=0.0, last_conv_kernel_size=None) -> None:
        act_ok = isinstance(activation, (tuple, list)) and str(activation[0]).upper() == "LEAKYRELU" and \
            abs(float(activation[1].get("negative_slope", 0.01)) - LRELU_SLOPE) < 1e-12
        if spatial_dims != 3 or out_channels != 1 or in_channels != 1 or kernel_size != 4 or not act_ok or \
                str(norm).upper() != "BATCH" or bias or padding != 1 or dropout not in (0, 0.0, None) or \
                last_conv_kernel_size not in (None, 4):
            raise NotImplementedError("petsyn PatchDiscriminator implements the reference's use: 3-D, one channel in / out, "
                                      "kernel 4, LeakyReLU(0.2), BatchNorm, no conv bias, padding 1, no dropout")
        super().__init__(num_channels, in_channels, num_layers_d)
        self._engines: Dict[Tuple, "_DiscEngine"] = ops.EngineCache()

    def engine_for(self, x: torch.Tensor) -> "_DiscEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _DiscEngine(self, tuple(x.shape), x.device, net=self)
            self._engines[key] = eng
        return eng

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("petsyn PatchDiscriminator runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        eng = self.engine_for(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in eng.params)):
            logits = _TapeFn.apply(eng, x, None, *eng.params)
        else:
            logits = eng.forward(x, None).clone()
        return _LastStageOnly(logits, self.num_layers_d + 2)


class patch_discriminator(nn.Module):
    """B200-native ``patch_discriminator`` (bmgan_model.py:133-144): ``PatchDiscriminator(3, 32, 1, num_layers_d=4)``;
    ``forward`` returns the last stage's patch logits ``[N, 1, d, h, w]``."""

    def __init__(self):
        super().__init__()
        self.patch_d = _PatchDiscriminatorNet(32, 1, 4)
        self._engines: Dict[Tuple, "_DiscEngine"] = ops.EngineCache()

    def engine_for(self, x: torch.Tensor) -> "_DiscEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _DiscEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("petsyn patch_discriminator runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        eng = self.engine_for(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in eng.params)):
            return _TapeFn.apply(eng, x, None, *eng.params)
        return eng.forward(x, None).clone()     # e.g. D(fake) then D(real) under no_grad (train_bmgan.py:243-247)


class _StemOp(graph.Op):
    """First PatchGAN conv (1 -> C, k4 s2 p1, bias) on the fp32 NCDHW input through the im2col GEMM."""

    def __init__(self, eng: "_DiscEngine", conv: nn.Conv3d, n, D, H, W, dev):
        self.eng, self.conv = eng, conv
        self.stem = ops.StemConv(n, D, H, W, conv.out_channels, dev)
        self.z = Buf(n, D // 2, H // 2, W // 2, conv.out_channels, dev, "d.stem.z")
        self.dpatches = torch.zeros_like(self.stem.patches)
        self.dbias = torch.zeros(conv.out_channels, dtype=torch.float64, device=dev)
        self.grad_w = self.grad_b = None
        self.acc_dw = False
        self.flops = 2.0 * self.z.rows * conv.out_channels * 64
        self._ver = None

    def repack(self) -> None:
        w = self.conv.weight
        ver = (w._version, w.data_ptr())
        if ver != self._ver:
            self.stem.plan.pack(w.detach().view(self.conv.out_channels, 64, 1, 1, 1), need_dgrad=True)
            self._ver = ver

    def fwd(self, training: bool) -> None:
        st = self.stem
        check(lib.petsyn_stem_im2col_k4s2(ptr(self.eng.x), ptr(st.patches), st.n, st.d, st.h, st.w, stream_ptr()),
              "stem_im2col")
        st.plan.fprop(st.patches, self.z.t, self.conv.bias.detach())

    def bwd(self) -> None:
        eng, st = self.eng, self.stem
        if eng.need_dw:
            st.plan.wgrad(st.patches, self.z.g, self.grad_w, accumulate=self.acc_dw)
            check(lib.petsyn_colsum(ptr(self.z.g), self.z.c, 0, ptr(self.dbias), self.z.rows, self.z.c, stream_ptr()),
                  "colsum")
            if self.acc_dw:
                self.grad_b.add_(self.dbias)
            else:
                self.grad_b.copy_(self.dbias)
        if eng.need_dx:
            st.plan.dgrad(self.z.g, self.dpatches)
            check(lib.petsyn_stem_col2im_k4s2(ptr(self.dpatches), ptr(eng.dx), st.n, st.d, st.h, st.w, stream_ptr()),
                  "stem_col2im")


class _DiscEngine(_EngineBase):
    def __init__(self, disc: nn.Module, shape, dev, net: Optional[_PatchDiscriminatorNet] = None):
        super().__init__(disc, dev)
        n, _, D, H, W = shape
        net = disc.patch_d if net is None else net
        L = net.num_layers_d
        if D % (1 << L) or H % (1 << L) or W % (1 << L):
            raise ValueError(f"spatial dims {D}x{H}x{W} must be divisible by {1 << L}")
        self.shape = shape
        self.x: Optional[torch.Tensor] = None
        self.dx = torch.zeros(n, 1, D, H, W, dtype=torch.float32, device=dev)
        self.need_dx, self.need_dw = True, True
        t = self.tape
        stem = _StemOp(self, net.initial_conv.conv, n, D, H, W, dev)
        t.add(stem)
        self._bind += [(stem, "grad_w", net.initial_conv.conv.weight), (stem, "grad_b", net.initial_conv.conv.bias)]
        a = Buf(n, D // 2, H // 2, W // 2, stem.z.c, dev, "d.a0")
        t.add(NormActOp(stem.z, "none", ops.ACT_LRELU, [a.sl()]))
        self.convs: List[ConvOp] = []
        for l_ in range(L):
            blk = getattr(net, str(l_))
            stride = blk.conv.stride[0]
            cv = self._conv(a.sl(), blk.conv, ksize=4, stride=stride, pad=1, name=f"d.{l_}")
            self.convs.append(cv)
            a = Buf(cv.z.n, cv.z.d, cv.z.h, cv.z.w, cv.z.c, dev, f"d.a{l_ + 1}")
            nm = NormActOp(cv.z, "batch", ops.ACT_LRELU, [a.sl()], bn=blk.adn.N)
            t.add(nm)
            self._bind += [(nm, "grad_gamma", blk.adn.N.weight), (nm, "grad_beta", blk.adn.N.bias)]
        self.head = self._conv(a.sl(), net.final_conv.conv, ksize=4, stride=1, pad=1, y_fp32=True, name="d.final")
        self.convs.append(self.head)
        od, oh, ow = self.head.plan.out_dims
        self.logits = torch.zeros(n, 1, od, oh, ow, dtype=torch.float32, device=dev)
        self._finish()

    def set_mode(self, need_dx: bool, need_dw: bool, accumulate: bool = False) -> None:
        """G phase: frozen weights, gradient w.r.t. the input; D phase: weight gradients only (``accumulate`` for the
        second of the two backward calls of train_bmgan.py:191-196)."""
        self.need_dx, self.need_dw = need_dx, need_dw
        for cv in self.convs:
            cv.need_dw = need_dw
        for op in self.tape.ops:
            if hasattr(op, "acc_dw"):
                op.acc_dw = accumulate

    def forward(self, x: torch.Tensor, _unused=None) -> torch.Tensor:
        self.generation += 1
        self.x = x
        self.tape.forward(self.training())
        check(lib.petsyn_take_channel0(ptr(self.head.zf), ptr(self.logits), self.logits.numel(), self.head.cout,
                                       stream_ptr()), "take_channel0")
        return self.logits

    def backward(self, dlogits: torch.Tensor, need_dx: bool = True, out: Optional[Dict[int, torch.Tensor]] = None,
                 need_dw: Optional[bool] = None, accumulate: bool = False):
        self.need_dx = need_dx
        if need_dw is None:
            need_dw = any(p.requires_grad for p in self.params)
        self.set_mode(need_dx, need_dw, accumulate)
        grads = self.grad_slots(out)
        check(lib.petsyn_put_channel0_grad(None, ptr(dlogits), ptr(self.head.zg), dlogits.numel(), self.head.cout, 0,
                                           stream_ptr()), "put_channel0_grad")
        self.tape.backward()
        dx = self.dx.clone() if need_dx else None
        if out is not None:
            return dx, []
        return dx, [g.clone() if need_dw else None for g in grads]


# ------------------------------------------------------------------------------------------------ encoder
class _ResidualUnitS2(_Container):
    """MONAI ``ResidualUnit(3, cin, c, strides=2, padding=1)``: ``conv.unit0`` (k3 s2) and ``conv.unit1`` (k3 s1), each
    Conv -> InstanceNorm -> PReLU; ``residual`` = Conv3d(cin, c, 3, 2, 1)."""

    def __init__(self, cin, c):
        super().__init__()
        self.conv = nn.Sequential()
        for su, (a, s) in enumerate(((cin, 2), (c, 1))):
            unit = nn.Sequential()
            unit.add_module("conv", nn.Conv3d(a, c, 3, stride=s, padding=1, bias=True))
            adn = nn.Sequential()
            adn.add_module("N", nn.InstanceNorm3d(c))
            adn.add_module("A", nn.PReLU())
            unit.add_module("adn", adn)
            self.conv.add_module(f"unit{su:d}", unit)
        self.residual = nn.Conv3d(cin, c, 3, 2, 1, bias=True)


class ResNet_encoder(nn.Module):
    """B200-native ``ResNet_encoder`` (bmgan_model.py:103-130): PET volume -> (mu, logvar) in R^8.  Like the reference
    (``nn.Linear(128*8, 8)``, :119) it accepts only volumes that six stride-2 stages reduce to 2x2x2."""

    def __init__(self, input_layer_channel=32, channels=[64, 128, 128, 128, 128, 128]):
        super().__init__()
        self.input_layer = nn.Sequential(nn.Conv3d(1, input_layer_channel, 3, padding=1),
                                         nn.InstanceNorm3d(input_layer_channel), nn.ReLU())
        self.resblocks = nn.ModuleList([])
        cur = input_layer_channel
        for c in channels:
            self.resblocks.append(_ResidualUnitS2(cur, c))
            cur = c
        self.linear1 = nn.Linear(128 * 8, 8)
        self.linear2 = nn.Linear(128 * 8, 8)
        self._engines: Dict[Tuple, "_EncEngine"] = ops.EngineCache()

    def engine_for(self, x: torch.Tensor) -> "_EncEngine":
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = _EncEngine(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def forward(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("petsyn ResNet_encoder runs on CUDA (sm_100a) only; there is no CPU path")
        if x.dim() != 5 or x.shape[1] != 1:
            raise ValueError(f"expected x of shape [N, 1, D, H, W], got {tuple(x.shape)}")
        x = x.contiguous().float()
        eng = self.engine_for(x)
        if torch.is_grad_enabled() and any(p.requires_grad for p in eng.params):
            out = _EncFn.apply(eng, x, *eng.params)
        else:
            out = eng.forward(x).clone()
        return out[:, :8], out[:, 8:16]


class _EncFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, *params):
        out = eng.forward(x).clone()
        eng.stamp(ctx, (x,))
        return out

    @staticmethod
    def backward(ctx, dout):
        ctx.eng.restore(ctx)
        grads = ctx.eng.backward(dout.contiguous().float())
        return (None, None, *grads)


class _LinearHeads(graph.Op):
    """linear1 / linear2 on nn.Flatten() of the NCDHW feature map = ONE k=1 conv on the channels-last buffer with the
    weight columns permuted from (c, voxel) to (voxel, c) order; fp32 output [N, 16] = [mu | logvar]."""

    def __init__(self, feat: Buf, lin1: nn.Linear, lin2: nn.Linear, dev):
        self.lin = (lin1, lin2)
        n, c = feat.n, feat.c
        self.vox = feat.rows // n
        self.c = c
        flat = feat.alias(n, 1, 1, 1, self.vox * c, "enc.flat")
        self.flat = flat
        self.plan = ops.ConvPlan(ops.OP_CONV, n, 1, 1, 1, self.vox * c, 16, 1, 1, 0, y_fp32=True)
        self.w = torch.zeros(16, self.vox * c, 1, 1, 1, dtype=torch.float32, device=dev)
        self.dw = torch.zeros_like(self.w)
        self.b = torch.zeros(16, dtype=torch.float32, device=dev)
        self.out = torch.zeros(n, 16, dtype=torch.float32, device=dev)
        self.dout = torch.zeros(n, 16, dtype=torch.bfloat16, device=dev)
        self.db = torch.zeros(16, dtype=torch.float64, device=dev)
        self.grad = {}          # id(param) -> tensor, bound by the engine
        self.acc_dw = False
        self.flops = 2.0 * n * self.vox * c * 16
        self._ver = None

    def repack(self) -> None:
        l1, l2 = self.lin
        ver = (l1.weight._version, l1.weight.data_ptr(), l2.weight._version, l1.bias._version, l2.bias._version)
        if ver == self._ver:
            return
        for i, l in enumerate(self.lin):      # Flatten order is (c, voxel); the buffer is (voxel, c)
            self.w[8 * i:8 * i + 8, :, 0, 0, 0].copy_(
                l.weight.detach().view(8, self.c, self.vox).permute(0, 2, 1).reshape(8, -1))
            self.b[8 * i:8 * i + 8].copy_(l.bias.detach())
        self.plan.pack(self.w, need_dgrad=True)
        self._ver = ver

    def fwd(self, training: bool) -> None:
        self.plan.fprop(self.flat.t, self.out, self.b)

    def grad_writes(self):
        return [("dx", self.flat.sl())]

    def bwd(self) -> None:
        self.plan.wgrad(self.flat.t, self.dout, self.dw)
        check(lib.petsyn_colsum(ptr(self.dout), 16, 0, ptr(self.db), self.dout.shape[0], 16, stream_ptr()), "colsum")
        for i, l in enumerate(self.lin):
            gw = self.dw[8 * i:8 * i + 8, :, 0, 0, 0].view(8, self.vox, self.c).permute(0, 2, 1).reshape(8, -1)
            if self.acc_dw:
                self.grad[id(l.weight)].add_(gw)
                self.grad[id(l.bias)].add_(self.db[8 * i:8 * i + 8])
            else:
                self.grad[id(l.weight)].copy_(gw)
                self.grad[id(l.bias)].copy_(self.db[8 * i:8 * i + 8])
        self.plan.dgrad(self.dout, self.flat.g)


class _EncEngine(_EngineBase):
    CPAD_IN = 16

    def __init__(self, enc: ResNet_encoder, shape, dev):
        super().__init__(enc, dev)
        n, _, D, H, W = shape
        self.shape = shape
        t = self.tape
        self.inp = Buf(n, D, H, W, self.CPAD_IN, dev, "enc.in")
        c0 = self._conv(self.inp.sl(), enc.input_layer[0], ksize=3, stride=1, pad=1, use_bias=False, need_dx=False,
                        name="enc.conv_in")
        a = Buf(n, D, H, W, c0.z.c, dev, "enc.a0")
        t.add(NormActOp(c0.z, "instance", ops.ACT_RELU, [a.sl()]))
        self.prelu_ops: List[Tuple[NormActOp, nn.Parameter]] = []
        for i, ru in enumerate(enc.resblocks):
            u0, u1 = ru.conv.unit0, ru.conv.unit1
            cv0 = self._conv(a.sl(), u0.conv, ksize=3, stride=2, pad=1, use_bias=False, name=f"enc.{i}.u0")
            zc = cv0.z
            h0 = Buf(zc.n, zc.d, zc.h, zc.w, zc.c, dev, f"enc.{i}.h0")
            n0 = NormActOp(zc, "instance", ops.ACT_PRELU, [h0.sl()], slope_param=u0.adn.A.weight)
            t.add(n0)
            cvr = self._conv(a.sl(), ru.residual, ksize=3, stride=2, pad=1, use_bias=True, name=f"enc.{i}.res")
            cv1 = self._conv(h0.sl(), u1.conv, ksize=3, stride=1, pad=1, use_bias=False, name=f"enc.{i}.u1")
            a = Buf(zc.n, zc.d, zc.h, zc.w, zc.c, dev, f"enc.{i}.out")
            n1 = NormActOp(cv1.z, "instance", ops.ACT_PRELU, [a.sl()], res=cvr.z.sl(), slope_param=u1.adn.A.weight)
            t.add(n1)
            for nm, prm in ((n0, u0.adn.A.weight), (n1, u1.adn.A.weight)):
                self._bind.append((nm, "grad_slope", prm))
        if (a.d, a.h, a.w) != (2, 2, 2) or a.c != 128:
            raise ValueError(f"ResNet_encoder needs an input that reduces to 128 x 2x2x2 (got {a.c} x {a.d}x{a.h}x{a.w}); "
                             "the reference hard-codes nn.Linear(128*8, 8)")
        self.heads = _LinearHeads(a, enc.linear1, enc.linear2, dev)
        t.add(self.heads)
        for l in (enc.linear1, enc.linear2):
            self._bind += [(self.heads, "_w", l.weight), (self.heads, "_b", l.bias)]
        self._finish()

    def grad_slots(self, out=None):
        grads = super().grad_slots(out)
        self.heads.grad = {id(p): g for p, g in zip(self.params, grads)}
        return grads

    def set_accumulate(self, acc: bool) -> None:
        for op in self.tape.ops:
            if hasattr(op, "acc_dw"):
                op.acc_dw = acc

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, _, D, H, W = self.shape
        self.generation += 1
        check(lib.petsyn_concat_latent(ptr(x), ptr(x), ptr(self.inp.t), D * H * W, n, 0, self.CPAD_IN, stream_ptr()),
              "concat_latent")
        self.tape.forward(self.training())
        return self.heads.out

    def backward(self, dout: torch.Tensor, out: Optional[Dict[int, torch.Tensor]] = None, accumulate: bool = False):
        grads = self.grad_slots(out)
        self.set_accumulate(accumulate)
        self.heads.dout.copy_(dout)
        self.tape.backward()
        return [g.clone() for g in grads] if out is None else []
