"""Drop-in replacement for the reference's ``unet/utils/unet_model.py`` (UnetGenerator3d and
UnetSkipConnectionBlock3d, reference lines 5-99) running on libpetsyn's sm_100a kernels.

What is kept identical to the reference (the drop-in boundary):
  * constructor signatures and defaults, the ``assert input_nc == output_nc`` (unet_model.py:12);
  * the module tree, hence every ``state_dict`` key/shape and the default initialisation (parameters are held by
    ordinary ``nn.Conv3d`` / ``nn.BatchNorm3d`` containers created in the reference's order, so a seeded
    construction draws the same weights and reference checkpoints load with ``load_state_dict``);
  * ``forward(input)``: fp32 NCDHW in, fp32 NCDHW out, autograd-differentiable, BatchNorm train/eval semantics with
    running-statistics updates, and the reference's in-place-activation aliasing (the skip half of every concat is
    ``LeakyReLU(x)``, ReLU'd again by the parent).

Constructor families: the reference configuration (``norm_layer=nn.BatchNorm3d``, no dropout layer; train_unet.py:59) runs on
the tuned engine ``_Engine``; ``norm_layer=nn.InstanceNorm3d`` (biased convolutions, unet_model.py:42-45), a
``functools.partial`` of either norm (e.g. ``affine=True``) and ``use_dropout=True`` with more than five levels
(unet_model.py:17, 87-88) run on the op tape (``_TapeEngine``).  Other norm classes raise ``NotImplementedError``.

What differs: nothing is computed by the container modules.  ``UnetGenerator3d.forward`` hands the whole nest to an
engine that runs fused CUDA kernels over channels-last bf16 activations (fp32 accumulation, fp32 master weights).
"""
from __future__ import annotations

import functools
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops
from ._cabi import check, lib, ptr, stream_ptr
from .bmgan_model import _EngineBase
from .graph import Buf, DropoutOp, NormActOp

LRELU_SLOPE = 0.2   # unet_model.py:49


class UnetSkipConnectionBlock3d(nn.Module):
    """Parameter container mirroring unet_model.py:37-99 (same children at the same Sequential indices)."""

    def __init__(self, outer_nc, inner_nc, submodule=None, outermost=False, innermost=False,
                 norm_layer=nn.BatchNorm3d, use_dropout=False):
        super().__init__()
        self.outermost = outermost
        self.innermost = innermost
        self.outer_nc, self.inner_nc = outer_nc, inner_nc
        norm_cls = norm_layer.func if isinstance(norm_layer, functools.partial) else norm_layer
        use_bias = norm_cls == nn.InstanceNorm3d                      # unet_model.py:42-45
        if norm_cls not in (nn.BatchNorm3d, nn.InstanceNorm3d):
            raise NotImplementedError("petsyn UnetGenerator3d implements norm_layer = nn.BatchNorm3d / nn.InstanceNorm3d "
                                      "(or a functools.partial of either)")
        self.use_bias = use_bias
        # the default family (bias-free convolutions, plain BatchNorm3d, no dropout layer) runs on the tuned engine below;
        # every other constructor family on the op tape (_TapeEngine)
        self.default_family = norm_layer is nn.BatchNorm3d and not (use_dropout and not (outermost or innermost))

        # creation order == the reference's (downconv, downnorm, upnorm, up conv) so seeded inits coincide
        downconv = nn.Conv3d(outer_nc, inner_nc, kernel_size=4, stride=2, padding=1, bias=use_bias)
        downrelu = nn.LeakyReLU(LRELU_SLOPE, True)
        downnorm = norm_layer(inner_nc)
        uprelu = nn.ReLU(True)
        upnorm = norm_layer(outer_nc)
        upsample = nn.Upsample(scale_factor=2)
        up_in = inner_nc if innermost else inner_nc * 2
        conv = nn.Conv3d(up_in, outer_nc, kernel_size=3, stride=1, padding=1, bias=use_bias)
        dropout = None
        if outermost:
            layers = [downconv, submodule, uprelu, upsample, conv, nn.Tanh()]
        elif innermost:
            layers = [downrelu, downconv, uprelu, upsample, conv, upnorm]
        else:
            layers = [downrelu, downconv, downnorm, submodule, uprelu, upsample, conv, upnorm]
            if use_dropout:                                           # unet_model.py:87-88
                dropout = nn.Dropout(0.5)
                layers.append(dropout)
        self.model = nn.Sequential(*layers)
        # handles for the engine (not registered twice: these are the same module objects)
        self._refs = dict(downconv=downconv, downnorm=None if (outermost or innermost) else downnorm, upconv=conv,
                          upnorm=None if outermost else upnorm, submodule=submodule, dropout=dropout)

    def forward(self, x):
        raise RuntimeError("petsyn blocks are parameter containers; call UnetGenerator3d.forward on the whole generator")


class UnetGenerator3d(nn.Module):
    """B200-native ``UnetGenerator3d`` (reference unet_model.py:5-32): same ctor, same keys, same forward contract."""

    def __init__(self, input_nc, output_nc, num_downs, ngf=64, norm_layer=nn.BatchNorm3d, use_dropout=False):
        super().__init__()
        assert (input_nc == output_nc)
        if input_nc != 1:
            raise NotImplementedError("petsyn UnetGenerator3d implements the reference configuration input_nc == output_nc == 1")
        blk = UnetSkipConnectionBlock3d(ngf * 8, ngf * 8, norm_layer=norm_layer, innermost=True)
        for _ in range(num_downs - 5):
            blk = UnetSkipConnectionBlock3d(ngf * 8, ngf * 8, blk, norm_layer=norm_layer, use_dropout=use_dropout)
        blk = UnetSkipConnectionBlock3d(ngf * 4, ngf * 8, blk, norm_layer=norm_layer)
        blk = UnetSkipConnectionBlock3d(ngf * 2, ngf * 4, blk, norm_layer=norm_layer)
        if num_downs >= 5:
            blk = UnetSkipConnectionBlock3d(ngf, ngf * 2, blk, norm_layer=norm_layer)
            blk = UnetSkipConnectionBlock3d(output_nc, ngf, blk, outermost=True, norm_layer=norm_layer)
        else:
            blk = UnetSkipConnectionBlock3d(output_nc, ngf * 2, blk, outermost=True, norm_layer=norm_layer)
        self.model = blk
        self._engines: Dict[Tuple, "_Engine"] = ops.EngineCache()

    # ------------------------------------------------------------------------------------------------
    def levels(self) -> List[UnetSkipConnectionBlock3d]:
        out, b = [], self.model
        while b is not None:
            out.append(b)
            b = b._refs["submodule"]
        return out

    def default_family(self) -> bool:
        return all(b.default_family for b in self.levels())

    def engine_for(self, x: torch.Tensor):
        key = (tuple(x.shape), x.device.index)
        eng = self._engines.get(key)
        if eng is None:
            eng = (_Engine if self.default_family() else _TapeEngine)(self, tuple(x.shape), x.device)
            self._engines[key] = eng
        return eng

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("petsyn UnetGenerator3d runs on CUDA (sm_100a) only; there is no CPU path")
        if input.dim() != 5 or input.shape[1] != 1:
            raise ValueError(f"expected input of shape [N, 1, D, H, W], got {tuple(input.shape)}")
        x = input.contiguous().float()
        eng = self.engine_for(x)
        if isinstance(eng, _TapeEngine):
            if torch.is_grad_enabled() and any(p.requires_grad for p in eng.params):
                return _TapeUnetFn.apply(eng, x, *eng.params)
            return eng.forward(x).clone()
        params = eng.param_list()
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _UnetFn.apply(x, eng, *params)
        return eng.forward(x, save=False)

    def flops_per_call(self, shape) -> Tuple[float, float]:
        """(algorithmic, executed) forward FLOPs for an input of ``shape`` (needs a CUDA device)."""
        eng = self._engines.get((tuple(shape), torch.cuda.current_device()))
        if eng is None:
            raise RuntimeError("run a forward pass on this shape first")
        return eng.flops_algorithmic, eng.flops_executed


class _UnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, eng, *params):
        y = eng.forward(x, save=True)
        ctx.eng, ctx.gen, ctx.mode = eng, eng.generation, eng.gen.training
        ctx.versions = tuple(p._version for p in params)
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        if eng.generation != ctx.gen:
            # a later forward overwrote this node's activations (forward, forward, backward): redo this node's forward
            # from its saved input, same mode, without moving the BatchNorm running statistics a second time
            if tuple(p._version for p in eng.param_list()) != ctx.versions:
                raise RuntimeError("a parameter of this petsyn module was modified in place between forward and backward")
            eng.forward(ctx.saved_tensors[0], save=True, clone_output=False, training=ctx.mode, update_running=False)
            ctx.gen = eng.generation
        grads = eng.backward(dy.contiguous().float())
        return (None, None, *grads)


class _Norm:
    """Buffers of one BatchNorm3d site."""

    def __init__(self, bn: nn.BatchNorm3d, c: int, dev):
        self.bn = bn
        f = lambda n=c: torch.empty(n, dtype=torch.float32, device=dev)
        self.sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)     # double accumulators (exact sums)
        self.bsums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
        self.scale, self.shift, self.mean, self.rstd = f(), f(), f(), f()
        self.dgamma, self.dbeta = f(), f()


class _Engine:
    """Buffers, plans and the fwd/bwd schedules for one (input shape, device)."""

    def __init__(self, gen: UnetGenerator3d, shape, dev):
        n, _, d, h, w = shape
        self.gen = gen
        self.shape = shape
        self.dev = dev
        lv = gen.levels()
        L = len(lv)
        self.L = L
        if d % (1 << L) or h % (1 << L) or w % (1 << L):
            raise ValueError(f"spatial dims {d}x{h}x{w} must be divisible by {1 << L} (one halving per level)")
        self.lv = lv
        bf = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)
        self.dims = [(d >> i, h >> i, w >> i) for i in range(L + 1)]          # resolution /2^i
        self.rows = [n * a * b * c for (a, b, c) in self.dims]
        outer = [b.outer_nc for b in lv]
        inner = [b.inner_nc for b in lv]
        self.outer, self.inner = outer, inner
        for c in outer[1:] + inner:
            if c % 8:
                raise ValueError(f"channel widths must be multiples of 8 (got {c}); use ngf % 4 == 0")
        # ---- activations (forward) ----
        self.z = [bf(self.rows[i + 1], inner[i]) for i in range(L)]            # raw down-conv outputs
        self.a = [None] + [bf(self.rows[i], outer[i]) for i in range(1, L)]    # LeakyReLU'd block inputs
        self.cat = [None] + [bf(self.rows[i], 2 * outer[i]) for i in range(1, L)]   # [ReLU(up_i) | ReLU(x_i)]
        self.zu = [None] + [bf(self.rows[i], outer[i]) for i in range(1, L)]   # raw up-conv outputs
        self.r = bf(self.rows[L], inner[L - 1])                                # ReLU(innermost down output)
        self.y = torch.empty(n, 1, d, h, w, dtype=torch.float32, device=dev)
        # ---- gradients ----
        self.dz = [bf(self.rows[i + 1], inner[i]) for i in range(L)]
        self.da = [None] + [bf(self.rows[i], outer[i]) for i in range(1, L)]
        self.dcat = [None] + [bf(self.rows[i], 2 * outer[i]) for i in range(1, L)]
        self.dzu = [None] + [bf(self.rows[i], outer[i]) for i in range(1, L)]
        self.dr = bf(self.rows[L], inner[L - 1])
        # ---- norms ----
        self.dnorm: List[Optional[_Norm]] = [None] * L
        self.unorm: List[Optional[_Norm]] = [None] * L
        for i, b in enumerate(lv):
            if b._refs["downnorm"] is not None:
                self.dnorm[i] = _Norm(b._refs["downnorm"], inner[i], dev)
            if b._refs["upnorm"] is not None:
                self.unorm[i] = _Norm(b._refs["upnorm"], outer[i], dev)
        # ---- conv plans ----
        d1, h1, w1 = self.dims[1]
        self.stem = ops.StemConv(n, d, h, w, inner[0], dev)
        self.head = ops.HeadConv(n, d1, h1, w1, 2 * inner[0], dev)
        self.down: List[Optional[ops.ConvPlan]] = [None] * L
        self.up: List[Optional[ops.ConvPlan]] = [None] * L
        for i in range(1, L):
            di, hi, wi = self.dims[i]
            self.down[i] = ops.ConvPlan(ops.OP_CONV, n, di, hi, wi, outer[i], inner[i], 4, 2, 1)
            dj, hj, wj = self.dims[i + 1]
            if i == L - 1:   # innermost: input is r (inner channels, contiguous)
                self.up[i] = ops.ConvPlan(ops.OP_UPCONV, n, dj, hj, wj, inner[i], outer[i], 3, 1, 1)
            else:            # input is cat_{i+1} (2*inner_i channels); dgrad writes dcat_{i+1}
                self.up[i] = ops.ConvPlan(ops.OP_UPCONV, n, dj, hj, wj, 2 * inner[i], outer[i], 3, 1, 1)
        self.flops_algorithmic = sum(p.flops_algorithmic for p in self.down[1:] + self.up[1:])
        self.flops_executed = sum(p.flops_executed for p in self.down[1:] + self.up[1:])
        stem = 2.0 * self.rows[1] * inner[0] * 64
        head = 2.0 * self.rows[0] * (2 * inner[0]) * 27
        self.flops_algorithmic += stem + head
        self.flops_executed += stem + 2.0 * self.rows[1] * (2 * inner[0]) * 27
        self._packed_versions: Dict[int, int] = {}
        self._grads: Optional[List[torch.Tensor]] = None
        self._saved_x: Optional[torch.Tensor] = None
        self.generation = 0             # bumped by every forward: one set of saved activations per engine (see _UnetFn)
        self.fwd_training = True

    # ------------------------------------------------------------------------------------------------ parameters
    def param_list(self) -> List[torch.Tensor]:
        ps = []
        for i, b in enumerate(self.lv):
            ps.append(b._refs["downconv"].weight)
            if self.dnorm[i] is not None:
                ps += [self.dnorm[i].bn.weight, self.dnorm[i].bn.bias]
            ps.append(b._refs["upconv"].weight)
            if self.unorm[i] is not None:
                ps += [self.unorm[i].bn.weight, self.unorm[i].bn.bias]
        return ps

    def _repack(self, need_dgrad: bool) -> None:
        """Refresh the packed bf16 operands of every conv whose fp32 master weight changed."""
        jobs = [(self.stem, self.lv[0]._refs["downconv"], lambda w: self.stem.pack(w)),
                (self.head, self.lv[0]._refs["upconv"], lambda w: self.head.pack(w, need_bwd=need_dgrad))]
        for i in range(1, self.L):
            for plan, conv in ((self.down[i], self.lv[i]._refs["downconv"]), (self.up[i], self.lv[i]._refs["upconv"])):
                jobs.append((plan, conv, lambda w, plan=plan: plan.pack(w, need_dgrad=need_dgrad)))
        for obj, conv, fn in jobs:
            w = conv.weight
            ver = (w._version, w.data_ptr(), need_dgrad)
            if self._packed_versions.get(id(obj)) != ver:
                fn(w.detach())
                self._packed_versions[id(obj)] = ver

    # ------------------------------------------------------------------------------------------------ forward
    def _bn_forward(self, nm: _Norm, z: torch.Tensor, rows: int, c: int, training: bool, update_running: bool) -> None:
        bn = nm.bn
        if bn.momentum is None:
            raise NotImplementedError("BatchNorm3d(momentum=None) (cumulative moving average) is not implemented")
        use_batch = training or bn.running_mean is None
        if use_batch:
            nm.sums.zero_()
            ops.bn_stats(z, nm.sums, rows, c)
            if update_running and bn.track_running_stats and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
        keep = use_batch and not update_running
        ops.bn_finalize(nm.sums, bn.weight, bn.bias, None if keep else bn.running_mean, None if keep else bn.running_var,
                        nm.scale, nm.shift, nm.mean, nm.rstd, rows, c, bn.eps, bn.momentum, use_batch)

    def forward(self, x: torch.Tensor, save: bool, timers=None, clone_output: bool = True,
                training: Optional[bool] = None, update_running: bool = True) -> torch.Tensor:
        L = self.L
        training = self.gen.training if training is None else training
        self.generation += 1
        self.fwd_training = training
        self._repack(need_dgrad=save)
        lv = self.lv
        self._saved_x = x if save else None
        # ---- down path ----
        self.stem.fprop(x, self.z[0])
        for i in range(1, L):
            prev = self.dnorm[i - 1]
            c = self.outer[i]
            if prev is not None:
                self._bn_forward(prev, self.z[i - 1], self.rows[i], c, training, update_running)
            ops.norm_act_fwd(self.z[i - 1], prev.scale if prev else None, prev.shift if prev else None,
                             self.a[i], c, 0, ops.ACT_LRELU, self.cat[i], 2 * c, c, ops.ACT_RELU, LRELU_SLOPE,
                             self.rows[i], c)
            with _timed(timers, f"down{i}.fprop"):
                self.down[i].fprop(self.a[i], self.z[i])
        # ---- innermost: ReLU(z) ----
        ci = self.inner[L - 1]
        ops.norm_act_fwd(self.z[L - 1], None, None, self.r, ci, 0, ops.ACT_RELU, None, 0, 0, ops.ACT_NONE, LRELU_SLOPE,
                         self.rows[L], ci)
        # ---- up path ----
        for i in range(L - 1, 0, -1):
            src = self.r if i == L - 1 else self.cat[i + 1]
            with _timed(timers, f"up{i}.fprop"):
                self.up[i].fprop(src, self.zu[i])
            nm = self.unorm[i]
            c = self.outer[i]
            self._bn_forward(nm, self.zu[i], self.rows[i], c, training, update_running)
            ops.norm_act_fwd(self.zu[i], nm.scale, nm.shift, self.cat[i], 2 * c, 0, ops.ACT_RELU, None, 0, 0,
                             ops.ACT_NONE, LRELU_SLOPE, self.rows[i], c)
        n, _, d, h, w = self.shape
        d1, h1, w1 = self.dims[1]
        cat1 = self.cat[1].view(n, d1, h1, w1, 2 * self.outer[1])
        y = self.y if save else torch.empty_like(self.y)
        self.head.fprop(cat1, y)
        return y.clone() if (save and clone_output) else y

    # ------------------------------------------------------------------------------------------------ backward
    def _grad_buffers(self) -> List[torch.Tensor]:
        if self._grads is None:
            self._grads = [torch.empty_like(p, dtype=torch.float32) for p in self.param_list()]
        return self._grads

    def grad_order(self) -> List[torch.Tensor]:
        """Parameters in the order their gradients are produced by ``backward`` (used to lay out the flat gradient
        arena so that data-parallel buckets complete front to back)."""
        L, lv = self.L, self.lv
        order = [lv[0]._refs["upconv"].weight]
        for i in range(1, L):
            order += [self.unorm[i].bn.weight, self.unorm[i].bn.bias, lv[i]._refs["upconv"].weight]
        for i in range(L - 1, 0, -1):
            order.append(lv[i]._refs["downconv"].weight)
            if self.dnorm[i - 1] is not None:
                order += [self.dnorm[i - 1].bn.weight, self.dnorm[i - 1].bn.bias]
        order.append(lv[0]._refs["downconv"].weight)
        return order

    def mark_weights_dirty(self) -> None:
        """Call after updating the fp32 master weights outside autograd's version tracking (fused optimiser)."""
        self._packed_versions.clear()

    def backward(self, dy: torch.Tensor, out: Optional[Dict[int, torch.Tensor]] = None, on_ready=None,
                 timers=None) -> List[torch.Tensor]:
        """Backward of ``forward(save=True)``.  Gradients are written (not accumulated) into ``out[id(param)]`` when
        given, else into engine-owned buffers whose clones are returned in ``param_list()`` order.  ``on_ready(param)``
        is called right after the kernels producing that parameter's gradient have been enqueued."""
        if out is None:
            grads = self._grad_buffers()
            slot: Dict[int, torch.Tensor] = {id(p): g for p, g in zip(self.param_list(), grads)}
        else:
            grads = None
            slot = out
        for p in self.backward_iter(dy, slot, timers):
            if on_ready is not None:
                on_ready(p)
        return [g.clone() for g in grads] if grads is not None else []

    def _wgrad_dgrad(self, plan, x, dz, gw, dx, timers, name: str) -> None:
        """Weight and data gradient of one conv.  They are independent: outside the profiling leg the data gradient runs
        on a side stream concurrently with the weight gradient (kept as a branch by CUDA-graph capture; joined before
        the caller yields, so a graph-segment cut never separates fork and join)."""
        if timers is None and not os.environ.get("PETSYN_NO_FORK"):
            from .graph import _side_stream
            main, side = torch.cuda.current_stream(), _side_stream(dz.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                plan.dgrad(dz, dx)
            plan.wgrad(x, dz, gw)
            main.wait_stream(side)
            return
        with _timed(timers, f"{name}.wgrad"):
            plan.wgrad(x, dz, gw)
        with _timed(timers, f"{name}.dgrad"):
            plan.dgrad(dz, dx)

    def backward_iter(self, dy: torch.Tensor, slot: Dict[int, torch.Tensor], timers=None):
        """Generator form of backward: enqueues kernels and yields each parameter as soon as the kernels producing
        its gradient (into ``slot[id(param)]``) have been enqueued.  The trainer uses the yield points to cut the
        backward pass into CUDA-graph segments with a bucket all-reduce launched between them."""
        L = self.L
        lv = self.lv
        if not self.fwd_training and any(nm is not None and nm.bn.running_mean is not None for nm in self.dnorm + self.unorm):
            raise NotImplementedError("backward through BatchNorm3d in eval() mode is not implemented; call .train()")
        gw = lambda conv: slot[id(conv.weight)]
        n = self.shape[0]
        d1, h1, w1 = self.dims[1]
        cat1 = self.cat[1].view(n, d1, h1, w1, 2 * self.outer[1])
        self.head.backward(cat1, self.y, dy, self.dcat[1], gw(lv[0]._refs["upconv"]))
        yield lv[0]._refs["upconv"].weight
        # ---- up path, outer -> inner ----
        for i in range(1, L):
            nm = self.unorm[i]
            c = self.outer[i]
            ops.norm_act_bwd(self.zu[i], nm.scale, nm.shift, nm.mean, nm.rstd, nm.bn.weight, self.dcat[i], 2 * c, 0,
                             ops.ACT_RELU, None, 0, 0, ops.ACT_NONE, LRELU_SLOPE, nm.bsums, self.dzu[i],
                             slot[id(nm.bn.weight)], slot[id(nm.bn.bias)], self.rows[i], c)
            yield nm.bn.weight
            yield nm.bn.bias
            src = self.r if i == L - 1 else self.cat[i + 1]
            dsrc = self.dr if i == L - 1 else self.dcat[i + 1]
            self._wgrad_dgrad(self.up[i], src, self.dzu[i], gw(lv[i]._refs["upconv"]), dsrc, timers, f"up{i}")
            yield lv[i]._refs["upconv"].weight
        # ---- innermost: through ReLU(z) ----
        ci = self.inner[L - 1]
        ops.norm_act_bwd(self.z[L - 1], None, None, None, None, None, self.dr, ci, 0, ops.ACT_RELU, None, 0, 0,
                         ops.ACT_NONE, LRELU_SLOPE, None, self.dz[L - 1], None, None, self.rows[L], ci)
        # ---- down path, inner -> outer ----
        for i in range(L - 1, 0, -1):
            self._wgrad_dgrad(self.down[i], self.a[i], self.dz[i], gw(lv[i]._refs["downconv"]), self.da[i], timers,
                              f"down{i}")
            yield lv[i]._refs["downconv"].weight
            prev = self.dnorm[i - 1]
            c = self.outer[i]
            ops.norm_act_bwd(self.z[i - 1], prev.scale if prev else None, prev.shift if prev else None,
                             prev.mean if prev else None, prev.rstd if prev else None,
                             prev.bn.weight if prev else None, self.da[i], c, 0, ops.ACT_LRELU, self.dcat[i], 2 * c, c,
                             ops.ACT_RELU, LRELU_SLOPE, prev.bsums if prev else None, self.dz[i - 1],
                             slot[id(prev.bn.weight)] if prev else None, slot[id(prev.bn.bias)] if prev else None,
                             self.rows[i], c)
            if prev is not None:
                yield prev.bn.weight
                yield prev.bn.bias
        self.stem.wgrad(self.dz[0], gw(lv[0]._refs["downconv"]))
        yield lv[0]._refs["downconv"].weight


# ======================================================================================================================
# The other constructor families -- norm_layer = nn.InstanceNorm3d (convolutions then carry biases, unet_model.py:42-45)
# or a functools.partial of either norm, and use_dropout=True with more than five levels (Dropout(0.5) after the up
# normalisation of the 8*ngf blocks, unet_model.py:17, 87-88) -- are built on the op tape shared with the BMGAN and
# AttenUNet mirrors (graph.py).  Same schedule as _Engine: the activated block input goes to two places in one launch
# (LeakyReLU for the next down convolution, ReLU into the skip half of the parent's concat buffer: the reference's in-place
# LeakyReLU followed by the parent's in-place ReLU, ReLU(LeakyReLU(x)) = ReLU(x)), torch.cat never runs, nn.Upsample is
# folded into the 3x3x3 convolution's gather.  Dropout commutes with the parent's ReLU (its mask is >= 0).
# ======================================================================================================================
class _TapeUnetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng, x, *params):
        y = eng.forward(x).clone()
        eng.stamp(ctx, (x,))
        return y

    @staticmethod
    def backward(ctx, dy):
        ctx.eng.restore(ctx)
        grads = ctx.eng.backward(dy.contiguous().float())
        return (None, None, *grads)


class _TapeEngine(_EngineBase):
    CPAD = 16

    def __init__(self, gen: UnetGenerator3d, shape, dev):
        super().__init__(gen, dev)
        n, _, D, H, W = shape
        lv = gen.levels()
        L = len(lv)
        if D % (1 << L) or H % (1 << L) or W % (1 << L):
            raise ValueError(f"spatial dims {D}x{H}x{W} must be divisible by {1 << L} (one halving per level)")
        outer, inner = [b.outer_nc for b in lv], [b.inner_nc for b in lv]
        for c in outer[1:] + inner:
            if c % 8:
                raise ValueError(f"channel widths must be multiples of 8 (got {c}); use ngf % 4 == 0")
        self.shape, self.gen = shape, gen
        t = self.tape
        B = lambda i, c, name: Buf(n, D >> i, H >> i, W >> i, c, dev, name)
        self.inp = B(0, self.CPAD, "input")
        cat = [None] + [B(i, 2 * outer[i], f"cat{i}") for i in range(1, L)]        # [ReLU(up_i) | ReLU(x_i)]
        # ---- down path ----
        z = self._conv(self.inp.sl(), lv[0]._refs["downconv"], ksize=4, stride=2, pad=1, need_dx=False, name="down0").z
        for i in range(1, L):
            c = outer[i]
            a = B(i, c, f"a{i}")
            self._norm(z, lv[i - 1]._refs["downnorm"], ops.ACT_LRELU, [a.sl(), cat[i].sl(c, c)], act2=ops.ACT_RELU,
                       name=f"down{i - 1}.norm")
            z = self._conv(a.sl(), lv[i]._refs["downconv"], ksize=4, stride=2, pad=1, name=f"down{i}").z
        r = B(L, inner[L - 1], "r")
        t.add(NormActOp(z, "none", ops.ACT_RELU, [r.sl()], name="inner.relu"))
        # ---- up path ----
        src = r.sl()
        for i in range(L - 1, 0, -1):
            c = outer[i]
            zu = self._conv(src, lv[i]._refs["upconv"], ksize=3, stride=1, pad=1, op=ops.OP_UPCONV, name=f"up{i}").z
            self._norm(zu, lv[i]._refs["upnorm"], ops.ACT_RELU, [cat[i].sl(0, c)], name=f"up{i}.norm")
            if lv[i]._refs["dropout"] is not None:
                t.add(DropoutOp(cat[i].sl(0, c), lv[i]._refs["dropout"]))
            src = cat[i].sl()
        self.head = self._conv(src, lv[0]._refs["upconv"], ksize=3, stride=1, pad=1, op=ops.OP_UPCONV, act=ops.ACT_TANH,
                               y_fp32=True, name="up0")
        self.y = torch.zeros(n, 1, D, H, W, dtype=torch.float32, device=dev)
        self._finish()
        self.flops_executed = sum(op.plan.flops_executed for op in self.tape.ops if hasattr(op, "plan"))   # incl. channel padding

    def _norm(self, z, norm: Optional[nn.Module], act: int, dsts, act2: Optional[int] = None, name: str = "") -> None:
        if norm is None:                                          # outermost / innermost down convolutions
            self.tape.add(NormActOp(z, "none", act, dsts, act2=act2, slope=LRELU_SLOPE, name=name))
            return
        if isinstance(norm, nn.BatchNorm3d):
            if norm.weight is None or norm.running_mean is None:
                raise NotImplementedError("BatchNorm3d(affine=False) / (track_running_stats=False) are not implemented")
            op = NormActOp(z, "batch", act, dsts, act2=act2, slope=LRELU_SLOPE, bn=norm, name=name)
            self._bind += [(op, "grad_gamma", norm.weight), (op, "grad_beta", norm.bias)]
        else:
            if norm.track_running_stats:
                raise NotImplementedError("InstanceNorm3d(track_running_stats=True) is not implemented")
            if norm.affine:                                       # per-channel groups of one: GroupNorm(C, C) arithmetic
                shim = SimpleNamespace(weight=norm.weight, bias=norm.bias, num_groups=norm.num_features, eps=norm.eps)
                op = NormActOp(z, "group", act, dsts, act2=act2, slope=LRELU_SLOPE, gn=shim, name=name)
                self._bind += [(op, "grad_gamma", norm.weight), (op, "grad_beta", norm.bias)]
            else:
                op = NormActOp(z, "instance", act, dsts, act2=act2, slope=LRELU_SLOPE, eps=norm.eps, name=name)
        self.tape.add(op)

    def param_list(self) -> List[torch.Tensor]:
        return self.params

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n, _, D, H, W = self.shape
        self.generation += 1
        check(lib.petsyn_concat_latent(ptr(x), ptr(x), ptr(self.inp.t), D * H * W, n, 0, self.CPAD, stream_ptr()),
              "concat_latent")
        self.tape.forward(self.training())
        check(lib.petsyn_take_channel0(ptr(self.head.zf), ptr(self.y), self.y.numel(), self.head.cout, stream_ptr()),
              "take_channel0")
        return self.y

    def backward(self, dy: torch.Tensor, out: Optional[Dict[int, torch.Tensor]] = None, on_ready=None):
        grads = self.grad_slots(out)
        check(lib.petsyn_put_channel0_grad(ptr(self.y), ptr(dy), ptr(self.head.zg), dy.numel(), self.head.cout, 1,
                                           stream_ptr()), "put_channel0_grad")
        self.run_backward(on_ready)
        return [g.clone() for g in grads] if out is None else []


class _timed:
    """Optional CUDA-event bracket around one launch (bench.py's roofline leg); a no-op when ``timers`` is None."""

    def __init__(self, timers, name):
        self.timers, self.name = timers, name

    def __enter__(self):
        if self.timers is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()

    def __exit__(self, *exc):
        if self.timers is not None:
            self.ev[1].record()
            self.timers.setdefault(self.name, []).append(self.ev)
        return False
