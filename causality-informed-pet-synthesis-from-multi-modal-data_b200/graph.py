"""A small static-graph executor for networks made of (Conv3d | ConvTranspose3d) -> (Instance|Batch)Norm -> activation
blocks with residual sums and dense channel concatenation -- the BMGAN generator/discriminator
(``bl_methods/BMGAN/bmgan_model.py:12-144``).

Tensors are channels-last bf16 buffers; an op reads/writes *channel slices* of buffers, so ``torch.cat`` is never
executed: producers write straight into the concat buffer of their consumer(s).  Backward is the reverse op list;
which gradient writes overwrite and which accumulate is decided once, statically, when the tape is finalised (the first
writer of a gradient region in backward order overwrites, later ones add -- through the conv kernels' TMA add-reduce
epilogue or the norm kernels' read-modify-write).  Every FLOP goes through ``ops`` (libpetsyn).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _cabi, ops
from ._cabi import check, lib, ptr, stream_ptr


class Buf:
    """Channels-last bf16 activation buffer [n*d*h*w, c] (+ lazily allocated gradient of the same shape)."""

    def __init__(self, n: int, d: int, h: int, w: int, c: int, device, name: str = ""):
        self.n, self.d, self.h, self.w, self.c = n, d, h, w, c
        self.rows = n * d * h * w
        self.name = name
        self.t = torch.zeros(self.rows, c, dtype=torch.bfloat16, device=device)
        self._g: Optional[torch.Tensor] = None

    @property
    def g(self) -> torch.Tensor:
        if self._g is None:
            self._g = torch.zeros_like(self.t)
        return self._g

    def alias(self, n: int, d: int, h: int, w: int, c: int, name: str = "") -> "Buf":
        """Same storage (values and gradients) seen with another shape, e.g. [N*8, 128] as [N, 1024] for nn.Flatten."""
        assert n * d * h * w * c == self.rows * self.c
        o = object.__new__(Buf)
        o.n, o.d, o.h, o.w, o.c, o.rows, o.name = n, d, h, w, c, n * d * h * w, name or self.name
        o.t = self.t.view(o.rows, c)
        o._g = self.g.view(o.rows, c)
        return o

    def sl(self, off: int = 0, c: Optional[int] = None) -> "Sl":
        return Sl(self, off, self.c - off if c is None else c)


class Sl:
    """Channel slice [off, off+c) of a buffer."""

    def __init__(self, buf: Buf, off: int, c: int):
        assert 0 <= off and off + c <= buf.c, (off, c, buf.c)
        self.buf, self.off, self.c = buf, off, c


class Op:
    def fwd(self, training: bool) -> None: ...
    def bwd(self) -> None: ...
    def grad_writes(self) -> List[Tuple[str, Sl]]:
        """Gradient regions this op writes in backward, in the order it writes them."""
        return []

    def repack(self) -> None: ...


_SIDE_STREAMS: Dict[int, torch.cuda.Stream] = {}


class RunState:
    """Process-wide switches read by the ops while a tape runs."""
    # True while an engine re-runs a forward pass only to restore the activations of an earlier autograd node (two
    # forwards, one backward: train_bmgan.py:170-180): BatchNorm normalises with the batch statistics again but must
    # not move its running statistics / num_batches_tracked a second time.
    freeze_running_stats = False
    # weight-gradient kernels were queued on the side stream and have not been joined into the main stream yet
    side_pending = False


def _side_stream(dev) -> torch.cuda.Stream:
    s = _SIDE_STREAMS.get(dev.index)
    if s is None:
        s = torch.cuda.Stream(device=dev)
        _SIDE_STREAMS[dev.index] = s
    return s


class ConvOp(Op):
    """z = conv(x) [+ bias] [-> activation]: raw output into its own contiguous buffer (bf16) or an fp32 tensor.

    Channel counts that are not multiples of 8 (the 9-channel BMGAN input, 1-channel outputs) are zero-padded in a
    staging copy of the weights; the padded activations' extra channels are zeros.
    """

    def __init__(self, x: Sl, weight: torch.nn.Parameter, bias: Optional[torch.nn.Parameter], ksize: int, stride: int,
                 pad: int, op: int = ops.OP_CONV, act: int = ops.ACT_NONE, slope: float = 0.2, y_fp32: bool = False,
                 use_bias: bool = True, need_dx: bool = True, need_dw: bool = True, name: str = "",
                 out: Optional["Sl"] = None, dy_from: Optional["Sl"] = None, cout_pad: Optional[int] = None,
                 out_index: Optional[Sequence[int]] = None, in_index: Optional[Sequence[int]] = None,
                 res: Optional["Sl"] = None):
        """``res``: a slice ADDED to the output by the kernel's epilogue (the residual sum of a ResnetBlock; needs
        ``plan.epi_ok[0]``; it may be the output slice itself, then an earlier op left the addend there).
        ``out_index`` / ``in_index`` (with ``cout_pad``): position of every true output / input channel of the weight inside
        the (wider, otherwise zero) channel range the kernels see -- attention heads of 8 or 16 channels padded to the
        kernel's 32-channel head: Conv3d / Linear layouts only."""
        self.x, self.weight, self.bias = x, weight, bias
        self.out_sl = out                      # write into a channel slice of an existing buffer instead of an own one
        # backward reads the gradient w.r.t. the output from THIS slice instead of z.g: the conv feeds a residual sum whose
        # result lives there (d(z + res) / dz = identity), so the sum needs no backward pass and z.g is never allocated
        self.dy_from = dy_from
        self.name = name
        self.opcode = op
        b = x.buf
        dev = b.t.device
        if op == ops.OP_CONVT:
            cin_w, cout_w = weight.shape[0], weight.shape[1]
        else:
            cout_w, cin_w = weight.shape[0], weight.shape[1]
        self.cin_w, self.cout_w = cin_w, cout_w
        self.cin, self.cout = x.c, (cout_w + 7) // 8 * 8
        if cout_w < 8:
            self.cout = 16                     # one-channel heads: smallest UMMA N with an unswizzled 32-byte row
        if cout_pad is not None:
            assert cout_pad >= self.cout and cout_pad % 8 == 0
            self.cout = cout_pad
        assert cin_w <= self.cin, (cin_w, self.cin)
        self.padded = (self.cin != cin_w) or (self.cout != cout_w) or out_index is not None or in_index is not None
        self._oidx = self._iidx = None
        if out_index is not None or in_index is not None:
            assert op != ops.OP_CONVT
            oi = list(out_index) if out_index is not None else list(range(cout_w))
            ii = list(in_index) if in_index is not None else list(range(cin_w))
            assert len(oi) == cout_w and len(ii) == cin_w and max(oi) < self.cout and max(ii) < self.cin
            self._oidx = torch.tensor(oi, dtype=torch.long, device=dev)
            self._iidx = torch.tensor(ii, dtype=torch.long, device=dev)
        self.use_bias = use_bias and bias is not None
        self.need_dx, self.need_dw = need_dx, need_dw
        ykw = {}
        if out is not None:
            assert out.c == self.cout and not y_fp32 and dy_from is None
            ykw = dict(y_cstride=out.buf.c, y_coff=out.off, dy_cstride=out.buf.c, dy_coff=out.off)
        if dy_from is not None:
            assert dy_from.c == self.cout and not y_fp32
            ykw = dict(dy_cstride=dy_from.buf.c, dy_coff=dy_from.off)
        self.plan = ops.ConvPlan(op, b.n, b.d, b.h, b.w, self.cin, self.cout, ksize, stride, pad,
                                 x_cstride=b.c, x_coff=x.off, dx_cstride=b.c, dx_coff=x.off, act=act, slope=slope,
                                 y_fp32=y_fp32, **ykw)
        od, oh, ow = self.plan.out_dims
        self.y_fp32 = y_fp32
        if out is not None:
            assert (out.buf.n, out.buf.d, out.buf.h, out.buf.w) == (b.n, od, oh, ow)
            self.z = None
        elif y_fp32:
            self.z = None
            self.zf = torch.zeros(b.n * od * oh * ow, self.cout, dtype=torch.float32, device=dev)
            self.zg = torch.zeros(b.n * od * oh * ow, self.cout, dtype=torch.bfloat16, device=dev)   # grad w.r.t. zf
        else:
            self.z = Buf(b.n, od, oh, ow, self.cout, dev, name + ".z")
        k3 = ksize ** 3
        if self.padded:
            shape = (self.cin, self.cout) if op == ops.OP_CONVT else (self.cout, self.cin)
            self.w_stage = torch.zeros(*shape, ksize, ksize, ksize, dtype=torch.float32, device=dev)
            self.dw_stage = torch.zeros_like(self.w_stage)
        self.bias_stage = None
        if self.use_bias and self.cout != cout_w:
            self.bias_stage = torch.zeros(self.cout, dtype=torch.float32, device=dev)
        self.dbias_stage = torch.zeros(self.cout, dtype=torch.float64, device=dev) if bias is not None else None
        self.tape_zeroes_dbias = False                 # Tape.finalize moved `dbias_stage` into its per-backward zero arena
        # the data gradient runs on a side stream concurrently with the weight gradient (a fork/join that CUDA-graph
        # capture keeps as a branch): layers whose kernels cannot fill 148 SMs (deep levels, transformer linears) overlap
        # fully, large ones overlap their ramp-up / tail.  Measured on configs[1]: 18.54 -> 17.82 ms per step.
        self.fork_bwd = not os.environ.get("PETSYN_NO_FORK")
        # "wgrad": the WEIGHT gradient (with its reduce / unpack launches) runs on the side stream and is only joined at the end
        # of the tape's backward (or where a gradient bucket closes), so it overlaps the critical chain data gradient ->
        # normalisation backward -> next data gradient, whose HBM-bound norm kernels use other resources than the
        # shared-memory-bound weight-gradient kernels.  "dgrad" (round 1): the data gradient forks, joined inside the op.
        self.fork_mode = os.environ.get("PETSYN_FORK_MODE", "wgrad")
        self.acc_dx = False
        self.colsum_done = False  # set by the NormActOp consuming z when it already summed dz over the rows (bias gradient)
        self.colsum_from: Optional["ConvOp"] = None   # another conv with the SAME output-gradient slice: its column sums are ours
        self.acc_dw = False      # add into grad_w / grad_b instead of overwriting (several backward calls per step)
        self._ver = None
        self._zero_b = None      # data_ptr of the bias-gradient slot that was cleared (biases with identically zero gradient)
        self.wg_scratch: Optional[torch.Tensor] = None  # Tape.finalize: ONE weight-gradient scratch for all convs of a tape
        self.grad_w: Optional[torch.Tensor] = None     # set by the owner: where dW / dbias go
        self.grad_b: Optional[torch.Tensor] = None
        self.flops = self.plan.flops_algorithmic
        # fused epilogues (Tape.finalize decides): statistics of the output for up to two consuming GroupNorms; the reduction
        # pass of the backward of the normalisation in front of this conv, done by the data-gradient kernel
        self.res = res
        if res is not None:
            assert self.plan.epi_ok[0] and res.c == self.cout and not y_fp32, "residual epilogue not available for this conv"
        self.stats_for: List[List[Tuple["NormActOp", int]]] = [[]]
        self.epi_norm: Optional["NormActOp"] = None
        self.absorbed = False    # the output is an addend that a later conv's epilogue folds into the same slice (`res`)

    # -------------------------------------------------------------------------------------------------
    # batched packing (Tape.repack): staleness check, operand allocation, staging of zero-padded weights
    def stale(self) -> bool:
        w = self.weight
        return (w._version, w.data_ptr(), self.need_dx) != self._ver

    def pack_source(self) -> torch.Tensor:
        return self.w_stage if self.padded else self.weight.detach()

    def alloc_packed(self) -> None:
        p, dev = self.plan, self.weight.device
        if p.w_fprop is None:
            p.w_fprop = torch.empty(p.packed_fprop_bytes, dtype=torch.uint8, device=dev)
        if self.need_dx and p.w_dgrad is None:
            p.w_dgrad = torch.empty(p.packed_dgrad_bytes, dtype=torch.uint8, device=dev)

    def _stage_weights(self) -> None:
        """fp32 master weights (and bias) -> the zero-padded staging copies the pack kernels read."""
        w = self.weight.detach()
        if self.padded:
            if self._oidx is not None:
                k = self.w_stage.shape[-1]
                self.w_stage[self._oidx[:, None], self._iidx[None, :]] = w.reshape(self.cout_w, self.cin_w, k, k, k)
            elif self.opcode == ops.OP_CONVT:
                self.w_stage[:self.cin_w, :self.cout_w].copy_(w)
            else:                                     # (an nn.Linear weight is 2-D: its kernel volume is 1)
                self.w_stage[:self.cout_w, :self.cin_w].copy_(w.reshape(self.cout_w, self.cin_w, *self.w_stage.shape[2:]))
        if self.bias_stage is not None:
            if self._oidx is not None:
                self.bias_stage[self._oidx] = self.bias.detach()
            else:
                self.bias_stage[:self.cout_w].copy_(self.bias.detach())

    def _dw_part(self) -> torch.Tensor:
        if self._oidx is not None:
            return self.dw_stage[self._oidx[:, None], self._iidx[None, :]]
        if self.opcode == ops.OP_CONVT:
            return self.dw_stage[:self.cin_w, :self.cout_w]
        return self.dw_stage[:self.cout_w, :self.cin_w]

    def _db_part(self) -> torch.Tensor:
        return self.dbias_stage[self._oidx] if self._oidx is not None else self.dbias_stage[:self.cout_w]

    def stage_padded(self) -> None:
        w = self.weight
        self._stage_weights()
        self._ver = (w._version, w.data_ptr(), self.need_dx)

    def repack(self, force: bool = False) -> None:
        w = self.weight
        ver = (w._version, w.data_ptr(), self.need_dx)
        if not force and ver == self._ver:
            return
        self._stage_weights()
        self.plan.pack(self.w_stage if self.padded else w.detach(), need_dgrad=self.need_dx)
        self._ver = ver

    def out(self) -> torch.Tensor:
        if self.out_sl is not None:
            return self.out_sl.buf.t
        return self.zf if self.y_fp32 else self.z.t

    def dy_slice(self) -> Optional["Sl"]:
        """Where backward reads d(loss)/d(output) from, as a slice (None for the fp32 heads)."""
        if self.dy_from is not None:
            return self.dy_from
        if self.out_sl is not None:
            return self.out_sl
        return None if self.y_fp32 else self.z.sl()

    def dout(self) -> torch.Tensor:
        if self.dy_from is not None:
            return self.dy_from.buf.g
        if self.out_sl is not None:
            return self.out_sl.buf.g
        return self.zg if self.y_fp32 else self.z.g

    def out_slice(self) -> Optional["Sl"]:
        """Where the forward output lives, as a slice (None for the fp32 heads)."""
        if self.out_sl is not None:
            return self.out_sl
        return None if self.y_fp32 else self.z.sl()

    def fwd(self, training: bool) -> None:
        bias = None
        if self.use_bias:
            bias = self.bias_stage if self.bias_stage is not None else self.bias.detach()
        tgts = self.stats_for[0]
        if self.res is not None or tgts:
            e = _cabi.ConvEpilogue()
            if self.res is not None:
                e.side, e.side_cstride, e.side_coff, e.add_side = ptr(self.res.buf.t), self.res.buf.c, self.res.off, 1
            for i, (q, off) in enumerate(tgts):
                if i == 0:
                    e.stats1, e.stats1_c, e.stats1_coff = ptr(q.sums), q.c, off
                else:
                    e.stats2, e.stats2_c, e.stats2_coff = ptr(q.sums), q.c, off
            self.plan.fprop_epi(self.x.buf.t, self.out(), bias, e)
            return
        self.plan.fprop(self.x.buf.t, self.out(), bias)

    def grad_writes(self):
        return [("dx", self.x)] if self.need_dx else []

    def _bwd_dx(self, dz: torch.Tensor) -> None:
        q = self.epi_norm
        if q is not None:
            assert not self.acc_dx
            e = _cabi.ConvEpilogue()
            e.side, e.side_cstride, e.side_coff = ptr(q.z.t), q.z.c, q.zs.off
            e.norm_scale, e.norm_shift, e.norm_mean, e.norm_rstd = ptr(q.scale), ptr(q.shift), ptr(q.mean), ptr(q.rstd)
            e.norm_act, e.norm_slope, e.bsums = q.act, q.slope, ptr(q.bsums)
            self.plan.dgrad_epi(dz, self.x.buf.g, e)
            return
        if self.acc_dx:
            check(lib.petsyn_conv_dgrad_accumulate(self.plan._h, ptr(dz), ptr(self.plan.w_dgrad), ptr(self.x.buf.g),
                                                   stream_ptr()), "conv_dgrad_accumulate")
        else:
            self.plan.dgrad(dz, self.x.buf.g)

    def _bwd_dw(self, dz: torch.Tensor) -> None:
        """Weight (and bias) gradient of this conv into grad_w / grad_b."""
        # bias gradient: column sums of dy (a by-product of the normalisation backward that wrote dy last, else a pass of
        # its own) sit in a float64 accumulator; the weight-gradient launch converts them into the fp32 slot
        bkw = {}
        if self.bias is not None:
            if self.use_bias:
                src = self.colsum_from
                if (src is not None and not self.padded and not src.padded and self.grad_b.is_contiguous()
                        and src.dbias_stage is not None):
                    # a ResnetBlock's conv2 and its 1x1 skip convolution read the same output gradient: one pass sums it
                    self.colsum_done = False
                    bkw = dict(dbias_acc=src.dbias_stage, dbias=self.grad_b)
                elif self.colsum_done:
                    self.colsum_done = False
                else:
                    dsl = self.dy_slice()
                    cs, co = (dsl.buf.c, dsl.off) if dsl is not None else (self.cout, 0)
                    check(lib.petsyn_colsum(ptr(dz), cs, co, ptr(self.dbias_stage), dz.shape[0], self.cout,
                                            stream_ptr()), "colsum")
                if not bkw and self.grad_b.is_contiguous() and not self.padded:
                    bkw = dict(dbias_acc=self.dbias_stage, dbias=self.grad_b)
            elif not self.acc_dw and self._zero_b != self.grad_b.data_ptr():
                # a bias in front of a non-affine InstanceNorm has exactly zero gradient and nothing ever writes its
                # slot: cleared once per binding instead of once per step
                self.grad_b.zero_()
                self._zero_b = self.grad_b.data_ptr()
        if self.padded:
            self.plan.wgrad(self.x.buf.t, dz, self.dw_stage, scratch=self.wg_scratch, **bkw)
            part = self._dw_part().reshape(self.grad_w.shape)
            if self.acc_dw:
                self.grad_w.add_(part)
            else:
                self.grad_w.copy_(part)
        else:
            self.plan.wgrad(self.x.buf.t, dz, self.grad_w, accumulate=self.acc_dw, scratch=self.wg_scratch, **bkw)
        if self.bias is not None and self.use_bias and not bkw:
            if self.acc_dw:
                self.grad_b.add_(self._db_part())
            else:
                self.grad_b.copy_(self._db_part())

    def bwd(self) -> None:
        dz = self.dout()
        fork = self.fork_bwd and self.need_dw and self.need_dx
        if self.fork_bwd and self.need_dw and self.fork_mode == "wgrad":
            # EVERY weight gradient of the tape goes to the side stream, also the ones without a data gradient to overlap
            # (conv_in): they share one scratch image, so they must stay serialised among themselves
            main, side = torch.cuda.current_stream(), _side_stream(dz.device)
            side.wait_stream(main)                       # dy (and the bias column sums) are final on the main stream here
            with torch.cuda.stream(side):
                self._bwd_dw(dz)
            RunState.side_pending = True                 # joined by Tape.backward / GradBucketer.on_ready
            if self.need_dx:
                self._bwd_dx(dz)
            return
        if fork:
            main, side = torch.cuda.current_stream(), _side_stream(dz.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self._bwd_dx(dz)
        if self.need_dw:
            self._bwd_dw(dz)
        if fork:
            main.wait_stream(side)
        elif self.need_dx:
            self._bwd_dx(dz)


def join_side_stream(dev=None) -> None:
    """Make the current stream wait for the weight-gradient work queued on the side stream (no-op when nothing is pending)."""
    if RunState.side_pending:
        d = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        torch.cuda.current_stream().wait_stream(_side_stream(d))
        RunState.side_pending = False


# debugging / A-B switch: run norm finalize and the backward group combine as their own launches (the first version)
SEPARATE_FINALIZE = bool(os.environ.get("PETSYN_SEPARATE_FINALIZE"))


class NormActOp(Op):
    """dst_i = act(norm(z)) [+ res] for one or two destinations (channel slices)."""

    def __init__(self, z, kind: str, act: int, dsts: Sequence[Sl], res: Optional[Sl] = None, slope: float = 0.2,
                 bn: Optional[torch.nn.BatchNorm3d] = None, eps: float = 1e-5, name: str = "",
                 slope_param: Optional[torch.nn.Parameter] = None, gn: Optional[torch.nn.GroupNorm] = None,
                 extra: Optional[Sl] = None, no_bwd: bool = False, act2: Optional[int] = None):
        """``z``: a whole buffer or a channel slice (Sl) of one.  ``act2``: activation of the second destination when it differs
        from the first's (UnetSkipConnectionBlock3d: LeakyReLU into the next convolution, ReLU into the skip half of a concat).  ``extra`` (backward): a gradient slice ADDED to dz -- the
        gradient reaching z through an identity skip connection.  ``no_bwd``: this op is a residual sum whose consumers'
        gradients have been redirected (ConvOp.dy_from, ``extra``, ResampleOp.dst_grad): nothing to do in backward."""
        assert kind in ("instance", "batch", "group", "none") and 1 <= len(dsts) <= 2
        self.zs: Sl = z if isinstance(z, Sl) else z.sl()
        z = self.zs.buf                         # the underlying buffer; self.zs.off / .c select the channels
        self.extra, self.no_bwd = extra, no_bwd
        self.act2 = act if act2 is None else act2
        self.gn = gn                            # nn.GroupNorm (kind == "group"): affine, num_groups, eps
        self.acc_dz = False
        self.fwd_batch_stats = True             # kind == "batch": the last forward normalised with batch statistics
        self.colsum_range = (0, 0)              # (first channel, count) of dz that belongs to colsum_conv; count 0 = all
        self.colsum_conv: Optional["ConvOp"] = None   # conv that produced z: its bias gradient is a by-product of bwd
        # statistics as a by-product: a "none" op may sum what it writes for the norm that consumes the destination ...
        self.stats_from_producers = False              # ... and that norm then skips its own statistics pass
        self.reduce_from_producer = False              # backward: (sum g, sum g zhat) come from the producing dgrad's epilogue
        self.tape_zeroes_sums = False                  # Tape.finalize moved `sums` into its per-step zero arena
        self.tape_zeroes_bsums = False                 # ... and `bsums` into its per-backward zero arena
        self.slope_param = slope_param          # nn.PReLU weight (one element): slope read from device memory
        self.grad_slope: Optional[torch.Tensor] = None
        self.z, self.kind, self.act, self.dsts, self.res, self.slope, self.bn, self.eps = z, kind, act, list(dsts), res, \
            slope, bn, eps
        self.name = name
        self.stats_for = [[] for _ in self.dsts]       # per destination: up to 2 x (consumer norm op, channel offset in its z)
        dev = z.t.device
        self.ns = z.n if kind in ("instance", "group") else 1
        self.rows = z.rows // self.ns
        c = self.c = self.zs.c
        if kind != "none":
            f = lambda m=1: torch.zeros(self.ns * c * m, dtype=torch.float32, device=dev)
            # reduction targets are DOUBLE accumulators (64-bit atomics of fp32 partials: exact, hence reproducible sums)
            self.sums, self.bsums = (torch.zeros(self.ns * c * 2, dtype=torch.float64, device=dev) for _ in range(2))
            self.scale, self.shift, self.mean, self.rstd = f(), f(), f(), f()
        self.dslope_stage = torch.zeros(1, dtype=torch.float64, device=dev) if slope_param is not None else None
        self.acc_res = False
        self.grad_gamma: Optional[torch.Tensor] = None
        self.grad_beta: Optional[torch.Tensor] = None
        self.acc_dw = False
        self._tmp_gb: Optional[torch.Tensor] = None

    def _desc(self, backward: bool) -> _cabi.NormActDesc:
        z, c = self.z, self.c
        d = _cabi.NormActDesc()
        d.z, d.rows, d.c, d.nsamples = ptr(z.t), self.rows, c, self.ns
        d.z_cstride, d.z_coff, d.dz_cstride, d.dz_coff = z.c, self.zs.off, z.c, self.zs.off
        d.per_sample_stats = 1 if self.kind in ("instance", "group") else 0
        if self.kind == "group":
            d.group_size = c // self.gn.num_groups
            if backward:
                d.gamma = ptr(self.gn.weight)
        d.dz_accumulate = int(self.acc_dz)
        if self.kind != "none":
            d.scale, d.shift = ptr(self.scale), ptr(self.shift)
            if not backward and self.kind in ("instance", "group") and not SEPARATE_FINALIZE:
                # finalize folded into the apply launch: sums in, scale / shift / mean / rstd out
                d.fin_sums, d.mean, d.rstd = ptr(self.sums), ptr(self.mean), ptr(self.rstd)
                if self.kind == "group":
                    d.fin_gamma, d.fin_beta = ptr(self.gn.weight), ptr(self.gn.bias)
                    d.fin_group_size, d.fin_eps = c // self.gn.num_groups, self.gn.eps
                else:
                    d.fin_group_size, d.fin_eps = 1, self.eps
            d.separate_group_combine = int(SEPARATE_FINALIZE)
            if backward:
                d.mean, d.rstd = ptr(self.mean), ptr(self.rstd)
                d.sums = ptr(self.bsums)
                d.sums_prezeroed = 1 if self.tape_zeroes_bsums else 0
                d.sums_precomputed = 1 if self.reduce_from_producer else 0
                if self.bn is not None and self.bn.weight is not None:
                    d.gamma = ptr(self.bn.weight)
        if not backward and any(self.stats_for):
            # per-sample launch so that the statistics land in the consumer's [sample][2][C] layout
            d.nsamples, d.rows = z.n, z.rows // z.n
            # one destination may feed two consuming normalisations (a skip tensor: the next block and the up path's concat
            # buffer); both destinations hold the same values, so the two descriptor slots are interchangeable
            tgts = [t for lst in self.stats_for for t in lst]
            assert len(tgts) <= 2
            for i, (q, off) in enumerate(tgts):
                if i == 0:
                    d.t1_stats, d.t1_stats_c, d.t1_stats_coff = ptr(q.sums), q.c, off
                else:
                    d.t2_stats, d.t2_stats_c, d.t2_stats_coff = ptr(q.sums), q.c, off
        src = (lambda s: s.buf.g) if backward else (lambda s: s.buf.t)
        s1 = self.dsts[0]
        d.t1, d.t1_cstride, d.t1_coff, d.act1 = ptr(src(s1)), s1.buf.c, s1.off, self.act
        if len(self.dsts) > 1:
            s2 = self.dsts[1]
            d.t2, d.t2_cstride, d.t2_coff, d.act2 = ptr(src(s2)), s2.buf.c, s2.off, self.act2
        d.slope = self.slope
        if self.slope_param is not None:
            d.slope_dev = ptr(self.slope_param)
            if backward and self.grad_slope is not None:
                d.dslope = ptr(self.dslope_stage)
        if self.res is not None:
            d.res, d.res_cstride, d.res_coff = ptr(src(self.res)), self.res.buf.c, self.res.off
            d.res_accumulate = int(self.acc_res)
        if backward:
            d.dz = ptr(z.g)
            if self.extra is not None:
                d.extra, d.extra_cstride, d.extra_coff = ptr(self.extra.buf.g), self.extra.buf.c, self.extra.off
            if self.colsum_conv is not None:
                d.dz_colsum = ptr(self.colsum_conv.dbias_stage)
                d.dz_colsum_coff, d.dz_colsum_c = self.colsum_range
            if self.grad_gamma is not None:
                if self.acc_dw:
                    if self._tmp_gb is None:
                        self._tmp_gb = torch.zeros(2, c, dtype=torch.float32, device=z.t.device)
                    d.dgamma, d.dbeta = ptr(self._tmp_gb[0]), ptr(self._tmp_gb[1])
                else:
                    d.dgamma, d.dbeta = ptr(self.grad_gamma), ptr(self.grad_beta)
        return d

    def fwd(self, training: bool) -> None:
        z, c = self.z, self.c
        stats = lambda ns, rows: check(lib.petsyn_norm_stats_slice(ptr(z.t), z.c, self.zs.off, ptr(self.sums), rows, c, ns,
                                                                   stream_ptr()), "norm_stats")
        if self.kind == "instance":
            if not self.stats_from_producers:
                if not self.tape_zeroes_sums:
                    self.sums.zero_()
                stats(z.n, self.rows)
            if SEPARATE_FINALIZE:        # default: folded into the apply launch below (fin_* fields of the descriptor)
                check(lib.petsyn_norm_finalize(ptr(self.sums), None, None, None, None, ptr(self.scale), ptr(self.shift),
                                               ptr(self.mean), ptr(self.rstd), self.rows, c, z.n, 1, self.eps, 0.0, 1,
                                               stream_ptr()), "norm_finalize")
        elif self.kind == "group":
            gn = self.gn
            if not self.stats_from_producers:
                if not self.tape_zeroes_sums:
                    self.sums.zero_()
                stats(z.n, self.rows)
            if SEPARATE_FINALIZE:
                check(lib.petsyn_norm_finalize(ptr(self.sums), ptr(gn.weight), ptr(gn.bias), None, None, ptr(self.scale),
                                               ptr(self.shift), ptr(self.mean), ptr(self.rstd), self.rows, c, z.n,
                                               c // gn.num_groups, gn.eps, 0.0, 1, stream_ptr()), "norm_finalize")
        elif self.kind == "batch":
            bn = self.bn
            if bn.momentum is None:
                raise NotImplementedError("BatchNorm3d(momentum=None) (cumulative moving average) is not implemented")
            use_batch = training or bn.running_mean is None
            freeze = RunState.freeze_running_stats
            if use_batch:
                if not self.tape_zeroes_sums:
                    self.sums.zero_()
                stats(1, z.rows)
                if bn.num_batches_tracked is not None and not freeze:
                    bn.num_batches_tracked.add_(1)
            self.fwd_batch_stats = use_batch
            rm, rv = (None, None) if (use_batch and freeze) else (ptr(bn.running_mean), ptr(bn.running_var))
            check(lib.petsyn_norm_finalize(ptr(self.sums), ptr(bn.weight), ptr(bn.bias), rm, rv, ptr(self.scale),
                                           ptr(self.shift), ptr(self.mean), ptr(self.rstd), z.rows, c, 1, 1, bn.eps,
                                           bn.momentum, int(use_batch), stream_ptr()), "norm_finalize")
        d = self._desc(False)
        check(lib.petsyn_normact_fwd(C.byref(d), stream_ptr()), "normact_fwd")

    def grad_writes(self):
        if self.no_bwd:
            return []
        w = [("dz", self.zs)]
        if self.res is not None:
            w.append(("dres", self.res))
        return w

    def bwd(self) -> None:
        if self.no_bwd:
            return
        if self.kind == "batch" and not self.fwd_batch_stats:
            # eval(): the layer is affine in its input (running statistics); the kernels implement the batch-statistics
            # backward only, which would silently add a mean / variance correction that does not exist here
            raise NotImplementedError("backward through BatchNorm3d in eval() mode is not implemented; call .train()")
        if self.grad_slope is not None:
            self.dslope_stage.zero_()
        if self.colsum_conv is not None:
            if not self.colsum_conv.tape_zeroes_dbias:
                self.colsum_conv.dbias_stage.zero_()
            self.colsum_conv.colsum_done = True
        d = self._desc(True)
        check(lib.petsyn_normact_bwd(C.byref(d), stream_ptr()), "normact_bwd")
        if self.grad_slope is not None:
            if self.acc_dw:
                self.grad_slope.add_(self.dslope_stage.view_as(self.grad_slope))
            else:
                self.grad_slope.copy_(self.dslope_stage.view_as(self.grad_slope))
        if self.acc_dw and self.grad_gamma is not None:
            self.grad_gamma.add_(self._tmp_gb[0])
            self.grad_beta.add_(self._tmp_gb[1])


class ResampleOp(Op):
    """dst = AvgPool3d(2,2)(src) (up=False) or nearest x2 upsampling (up=True) of a whole buffer."""

    def __init__(self, src: Sl, dst: Sl, up: bool, dst_grad: Optional[Sl] = None):
        """``dst_grad``: backward reads d(loss)/d(dst) from this slice instead of dst's own gradient buffer (dst is the
        identity branch of a residual sum whose result -- and therefore whose gradient -- lives there)."""
        self.src, self.dst, self.up = src, dst, up
        self.dst_grad = dst_grad
        self.acc_dx = False

    def _run(self, a: torch.Tensor, sa: Sl, b: torch.Tensor, sb: Sl, up: bool, scale: float, acc: bool) -> None:
        ob = sb.buf
        check(lib.petsyn_resample2(ptr(a), sa.buf.c, sa.off, ptr(b), ob.c, sb.off, ob.n, ob.d, ob.h, ob.w, sb.c, int(up),
                                   scale, int(acc), stream_ptr()), "resample2")

    def fwd(self, training: bool) -> None:
        self._run(self.src.buf.t, self.src, self.dst.buf.t, self.dst, self.up, 1.0 if self.up else 0.125, False)

    def grad_writes(self):
        return [("dx", self.src)]

    def bwd(self) -> None:      # transpose of the forward map
        dg = self.dst_grad if self.dst_grad is not None else self.dst
        self._run(dg.buf.g, dg, self.src.buf.g, self.src, not self.up, 1.0 if self.up else 0.125, self.acc_dx)


class LayerNormOp(Op):
    """y = nn.LayerNorm(c)(x) on a token stream (whole buffers)."""

    def __init__(self, x: Buf, y: Buf, ln: torch.nn.LayerNorm):
        self.x, self.y, self.ln = x, y, ln
        dev = x.t.device
        self.mean = torch.zeros(x.rows, dtype=torch.float32, device=dev)
        self.rstd = torch.zeros(x.rows, dtype=torch.float32, device=dev)
        self.grad_gamma = self.grad_beta = None
        self.acc_dx = False

    def fwd(self, training: bool) -> None:
        check(lib.petsyn_layernorm_fwd(ptr(self.x.t), ptr(self.ln.weight), ptr(self.ln.bias), ptr(self.y.t),
                                       ptr(self.mean), ptr(self.rstd), self.x.rows, self.x.c, self.ln.eps, stream_ptr()),
              "layernorm_fwd")

    def grad_writes(self):
        return [("dx", self.x.sl())]

    def bwd(self) -> None:
        check(lib.petsyn_layernorm_bwd(ptr(self.x.t), ptr(self.y.g), ptr(self.ln.weight), ptr(self.mean), ptr(self.rstd),
                                       ptr(self.x.g), ptr(self.grad_gamma), ptr(self.grad_beta), self.x.rows, self.x.c,
                                       int(self.acc_dx), stream_ptr()), "layernorm_bwd")


class GegluOp(Op):
    """out = x * gelu(gate) with h = (x | gate) (MONAI MLPBlock act="GEGLU")."""
    can_accumulate = False

    def __init__(self, h: Buf, out: Buf):
        assert h.c == 2 * out.c
        self.h, self.o = h, out

    def fwd(self, training: bool) -> None:
        check(lib.petsyn_geglu_fwd(ptr(self.h.t), ptr(self.o.t), self.h.rows, self.o.c, stream_ptr()), "geglu_fwd")

    def grad_writes(self):
        return [("dx", self.h.sl())]

    def bwd(self) -> None:
        check(lib.petsyn_geglu_bwd(ptr(self.h.t), ptr(self.o.g), ptr(self.h.g), self.h.rows, self.o.c, stream_ptr()),
              "geglu_bwd")


class AttentionOp(Op):
    """o = softmax(scale * q k^T) v per (sample, head); qkv = (q | k | v) columns of one buffer."""
    can_accumulate = False

    def __init__(self, qkv: Buf, out: Buf, heads: int, tokens_per_sample: int, true_head_dim: Optional[int] = None):
        """``true_head_dim``: the model's head width when the buffers hold heads zero-padded to the kernel's 32 channels (the
        softmax scale is 1 / sqrt(true width); zero channels change neither q k^T nor p v)."""
        self.qkv, self.o, self.heads, self.L = qkv, out, heads, tokens_per_sample
        self.n = qkv.rows // tokens_per_sample
        self.hd = out.c // heads
        dev = qkv.t.device
        self.lse = torch.zeros(self.n * heads * self.L, dtype=torch.float32, device=dev)
        self.delta = torch.zeros_like(self.lse)
        thd = true_head_dim or self.hd
        self.scale = 1.0 / (thd ** 0.5)
        self.flops = 4.0 * self.n * heads * self.L * self.L * thd

    def fwd(self, training: bool) -> None:
        check(lib.petsyn_attention_fwd(ptr(self.qkv.t), ptr(self.o.t), ptr(self.lse), self.n, self.L, self.heads, self.hd,
                                       self.scale, stream_ptr()), "attention_fwd")

    def grad_writes(self):
        return [("dx", self.qkv.sl())]

    def bwd(self) -> None:
        check(lib.petsyn_attention_bwd(ptr(self.qkv.t), ptr(self.o.t), ptr(self.o.g), ptr(self.lse), ptr(self.delta),
                                       ptr(self.qkv.g), self.n, self.L, self.heads, self.hd, self.scale, stream_ptr()),
              "attention_bwd")


class CovariateBiasOp(Op):
    """tokens += to_out(to_v(context)) broadcast over the tokens of each sample: what cross-attention over a length-1
    context computes (atten_unet_model.py:156-175).  In place; the gradient w.r.t. the tokens passes
    through unchanged, to_q / to_k receive exactly zero gradient."""

    def __init__(self, tokens: Buf, owner, to_v: torch.nn.Linear, to_out: torch.nn.Linear, tokens_per_sample: int):
        self.t, self.owner, self.to_v, self.to_out, self.L = tokens, owner, to_v, to_out, tokens_per_sample
        self.n = tokens.rows // tokens_per_sample
        dev = tokens.t.device
        c = tokens.c
        self.vbuf = torch.zeros(self.n, c, dtype=torch.float32, device=dev)
        self.bias = torch.zeros(self.n, c, dtype=torch.float32, device=dev)
        self.dbias = torch.zeros(self.n, c, dtype=torch.float32, device=dev)
        self.grad_wv = self.grad_wo = self.grad_bo = None

    def fwd(self, training: bool) -> None:
        ctx = self.owner.context
        check(lib.petsyn_covariate_bias_fwd(ptr(ctx), ptr(self.to_v.weight), ptr(self.to_out.weight),
                                            ptr(self.to_out.bias), ptr(self.vbuf), ptr(self.bias), ptr(self.t.t), self.n,
                                            ctx.shape[1], self.t.c, self.L, stream_ptr()), "covariate_bias_fwd")

    def bwd(self) -> None:
        ctx = self.owner.context
        check(lib.petsyn_covariate_bias_bwd(ptr(ctx), ptr(self.to_out.weight), ptr(self.vbuf), ptr(self.t.g),
                                            ptr(self.dbias), ptr(self.grad_wv), ptr(self.grad_wo), ptr(self.grad_bo),
                                            self.n, ctx.shape[1], self.t.c, self.L, stream_ptr()), "covariate_bias_bwd")


class DropoutOp(Op):
    """nn.Dropout(p) on a small token buffer, in place (the classifier head: Linear -> ReLU -> Dropout(0.1) -> Linear,
    atten_unet_model.py:1987).  The mask comes from torch's device generator (one rand launch on a [N, 512] tensor); identity
    in eval mode or with p == 0."""

    def __init__(self, x, module: torch.nn.Dropout):
        """``x``: a buffer or a channel slice of one (UnetSkipConnectionBlock3d's Dropout(0.5) acts on the up half of a
        concat buffer, unet_model.py:88)."""
        self.xs: Sl = x if isinstance(x, Sl) else x.sl()
        self.x, self.module = self.xs.buf, module
        self.mask: Optional[torch.Tensor] = None

    def _view(self, t: torch.Tensor) -> torch.Tensor:
        return t[:, self.xs.off:self.xs.off + self.xs.c]

    def fwd(self, training: bool) -> None:
        p = float(self.module.p)
        if not training or p == 0.0:
            self.mask = None
            return
        v = self._view(self.x.t)
        if p >= 1.0:
            self.mask = torch.zeros(v.shape, dtype=torch.bfloat16, device=v.device)
        else:
            self.mask = (torch.rand(v.shape, device=v.device) >= p).to(torch.bfloat16) / (1.0 - p)
        v.mul_(self.mask)

    def bwd(self) -> None:
        if self.mask is not None:
            self._view(self.x.g).mul_(self.mask)


class Tape:
    """Ordered op list with static gradient write/accumulate analysis."""

    def __init__(self):
        self.ops: List[Op] = []
        self._final = False

    def add(self, op: Op) -> Op:
        self.ops.append(op)
        return op

    def finalize(self) -> None:
        """Walk backward order once: the first writer of a gradient region overwrites, later writers accumulate.  A
        region may only be accumulated into if an earlier write covers it entirely."""
        init: Dict[int, List[Tuple[int, int]]] = {}
        for op in reversed(self.ops):
            for key, s in op.grad_writes():
                ranges = init.setdefault(id(s.buf), [])
                lo, hi = s.off, s.off + s.c
                covered = any(a <= lo and hi <= b for a, b in ranges)
                overlap = any(a < hi and lo < b for a, b in ranges)
                if overlap and not covered:
                    raise RuntimeError(f"gradient region {s.buf.name}[{lo}:{hi}] partially overlaps an earlier write")
                if covered and not getattr(op, "can_accumulate", True):
                    raise RuntimeError(f"{type(op).__name__} cannot accumulate into an already written gradient region")
                if key == "dx":
                    op.acc_dx = covered
                elif key == "dres":
                    op.acc_res = covered
                elif key == "dz":
                    op.acc_dz = covered
                if not covered:
                    ranges.append((lo, hi))
        # bias gradients as a by-product: the gradient w.r.t. a conv's output (its own z, or the slice ConvOp.dy_from points at)
        # is complete after its LAST writer in backward order; when that writer is a NormActOp's dz covering exactly that
        # slice, its apply pass also sums the final values over the rows (after any accumulation) = the bias gradient
        writes: List[Tuple[Op, str, Sl]] = []
        for op in reversed(self.ops):
            for key, sl in op.grad_writes():
                writes.append((op, key, sl))
        for op in self.ops:
            if isinstance(op, NormActOp):
                op.colsum_conv = None
        for cv in self.ops:
            if not (isinstance(cv, ConvOp) and cv.use_bias and cv.need_dw and cv.bias is not None):
                continue
            region = cv.dy_slice()
            if region is None:
                continue
            lo, hi = region.off, region.off + region.c
            hits = [(op, key, sl) for op, key, sl in writes
                    if sl.buf is region.buf and sl.off < hi and lo < sl.off + sl.c]
            if not hits:
                continue
            op, key, sl = hits[-1]
            if isinstance(op, NormActOp) and key == "dz" and sl.off <= lo and hi <= sl.off + sl.c and op.colsum_conv is None:
                op.colsum_conv = cv                    # the conv's channels may be a sub-range of the op's dz
                op.colsum_range = (lo - sl.off, region.c)
        # convs that read the SAME output-gradient slice (conv2 and the 1x1 skip convolution of a ResnetBlock) share one
        # column-sum pass: the owner is the conv whose sums are a normalisation's by-product, else the one that runs first in
        # backward (all weight gradients of a tape are serialised in backward order)
        same: Dict[Tuple[int, int, int], List[ConvOp]] = {}
        for cv in self.ops:
            if isinstance(cv, ConvOp):
                cv.colsum_from = None
                if cv.use_bias and cv.need_dw and cv.bias is not None and cv.dy_slice() is not None:
                    r = cv.dy_slice()
                    same.setdefault((id(r.buf), r.off, r.c), []).append(cv)
        if not os.environ.get("PETSYN_NO_COLSUM_SHARING"):
            fed = {id(op.colsum_conv) for op in self.ops if isinstance(op, NormActOp) and op.colsum_conv is not None}
            for lst in same.values():
                if len(lst) < 2:
                    continue
                by = [cv for cv in lst if id(cv) in fed]
                owner = by[0] if by else lst[-1]          # lst is in tape order: the last one runs first in backward
                for cv in lst:
                    if cv is not owner:
                        cv.colsum_from = owner
        # statistics as a by-product: a GroupNorm whose input (a buffer or a channel slice of one) is written, channel range
        # by channel range, only by un-normalised NormActOps (residual sums, copies into concat buffers) takes its sums from
        # those producers; a destination may serve two consuming normalisations
        produced: Dict[int, List[Tuple[NormActOp, int, Sl]]] = {}
        other_writers: Dict[int, List[Tuple[int, int]]] = {}

        def other(sl: Sl) -> None:
            other_writers.setdefault(id(sl.buf), []).append((sl.off, sl.off + sl.c))

        for op in self.ops:
            if isinstance(op, NormActOp):
                for i, sl in enumerate(op.dsts):
                    if op.kind == "none" and op.slope_param is None:
                        produced.setdefault(id(sl.buf), []).append((op, i, sl))
                    else:
                        other(sl)
            elif isinstance(op, ConvOp):
                osl = op.out_slice()
                if osl is None or op.absorbed:
                    continue                           # fp32 head / an addend that a later conv's epilogue overwrites in place
                if op.plan.epi_stats_ok and not os.environ.get("PETSYN_NO_EPI_STATS"):
                    op.stats_for = [[]]
                    produced.setdefault(id(osl.buf), []).append((op, 0, osl))   # its epilogue can sum what it stores
                else:
                    other(osl)
            elif isinstance(op, ResampleOp):
                other(op.dst)                          # src / dst_grad are only read
            else:
                for v in vars(op).values():            # any other op that holds the buffer may write it: be conservative
                    if isinstance(v, Buf):
                        other(v.sl())
                    elif isinstance(v, Sl):
                        other(v)
        shared = []
        for q in self.ops:
            if not (isinstance(q, NormActOp) and q.kind in ("group", "instance") and not os.environ.get("PETSYN_NO_STATS_FUSION")):
                continue
            lo, hi = q.zs.off, q.zs.off + q.zs.c
            if any(a < hi and lo < b for a, b in other_writers.get(id(q.z), [])):
                continue
            prods = [(p, i, sl) for p, i, sl in produced.get(id(q.z), []) if sl.off < hi and lo < sl.off + sl.c]
            if not prods:
                continue
            cover = sorted((sl.off, sl.off + sl.c) for _, _, sl in prods)
            if cover[0][0] != lo or cover[-1][1] != hi or any(a[1] != b[0] for a, b in zip(cover, cover[1:])):
                continue
            if any(sum(len(t) for t in p.stats_for) >= 2 for p, _, _ in prods):
                continue                               # an op carries at most two statistics targets
            for p, i, sl in prods:
                p.stats_for[i].append((q, sl.off - lo))
            q.stats_from_producers = True
            shared.append(q)
        # every small fp32 accumulator a step starts from zero lives in one of two arenas, cleared by ONE fill at the start
        # of forward / backward instead of a fill launch per op (~90 launches per AttenUNet step)
        def arena(pairs):
            if not pairs:
                return None
            dev = getattr(*pairs[0]).device
            sizes = [(getattr(h, a).numel() + 3) // 4 * 4 for h, a in pairs]            # 16-byte aligned views
            dtype = getattr(*pairs[0]).dtype
            assert all(getattr(h, a).dtype == dtype for h, a in pairs)
            buf = torch.zeros(sum(sizes), dtype=dtype, device=dev)
            off = 0
            for (h, a), n_ in zip(pairs, sizes):
                old_t = getattr(h, a)
                setattr(h, a, buf[off:off + old_t.numel()].view_as(old_t))
                off += n_
            return buf

        use = not os.environ.get("PETSYN_NO_ZERO_ARENA")
        normed = [op for op in self.ops if isinstance(op, NormActOp) and op.kind in ("instance", "group", "batch")]
        own = [op for op in normed if not op.stats_from_producers and use]
        for op in own:
            op.tape_zeroes_sums = True
        self._stats_arena = arena([(op, "sums") for op in shared + own])
        bwd_pairs = []
        if use:
            for op in self.ops:
                c = op.colsum_conv if isinstance(op, NormActOp) else None
                if c is not None and c.dbias_stage is not None and not c.tape_zeroes_dbias:
                    c.tape_zeroes_dbias = True
                    bwd_pairs.append((c, "dbias_stage"))
            for op in normed:
                op.tape_zeroes_bsums = True
                bwd_pairs.append((op, "bsums"))
        self._bwd_arena = arena(bwd_pairs)
        # backward reduction as a by-product: the data gradient of the conv that consumes a = act(norm(z)) is the ONLY writer of
        # a's gradient; its epilogue (petsyn_conv_dgrad_epi) reads z beside the tile it stores and accumulates (sum g, sum g zhat)
        # into the normalisation's backward sums, which then runs its apply pass only
        for op in self.ops:
            if isinstance(op, ConvOp):
                op.epi_norm = None
        # Measured (tools/conv_layer_bench.py --pass epi, B200): the sigmoid per element makes the epilogue warps the bottleneck of
        # the shared-memory-bound slab kernel -- 115 us fused against 56 us (data gradient) + 33 us (reduction pass) on the
        # full-resolution 16 -> 16 layer -- so this fusion is opt-in (PETSYN_EPI_BWD=1); the forward fusions are free and on
        if use and os.environ.get("PETSYN_EPI_BWD"):
            for q in normed:
                q.reduce_from_producer = False
                if not (q.kind in ("instance", "group") and len(q.dsts) == 1 and not q.no_bwd and q.slope_param is None
                        and q.ns == q.z.n):
                    continue
                a = q.dsts[0]
                ws = [(op, key, sl) for op, key, sl in writes if sl.buf is a.buf and sl.off < a.off + a.c and a.off < sl.off + sl.c]
                if len(ws) != 1:
                    continue
                op, key, sl = ws[0]
                if (isinstance(op, ConvOp) and key == "dx" and sl.off == a.off and sl.c == a.c and op.cin == q.c
                        and op.plan.epi_ok[1] and not op.acc_dx and op.epi_norm is None):
                    op.epi_norm = q
                    q.reduce_from_producer = True
        # the weight-gradient kernels write per-CTA / per-split partial images into a scratch that the unpack kernel sums in a
        # fixed order; all weight gradients of a tape run on its main stream one after the other, so they share ONE scratch
        convs = [op for op in self.ops if isinstance(op, ConvOp) and op.need_dw]
        if convs:
            nbytes = max(op.plan.wgrad_scratch_bytes for op in convs)
            shared = torch.empty(nbytes, dtype=torch.uint8, device=convs[0].weight.device)
            for op in convs:
                op.wg_scratch = shared
        self._final = True

    def repack(self) -> None:
        """Refresh the packed bf16 weight images of every conv whose parameter changed.  All of them go through ONE
        batched pack (``petsyn_pack_batch_*``: a launch per kernel-volume class, not per tensor); the batch is rebuilt
        when a weight tensor moves (``.to()``, flat-arena adoption)."""
        convs = [op for op in self.ops if isinstance(op, ConvOp)]
        for op in self.ops:
            if not isinstance(op, ConvOp):
                op.repack()
        stale = [op for op in convs if op.stale()]
        if not stale:
            return
        key = tuple((op.weight.data_ptr(), op.need_dx) for op in convs)
        if getattr(self, "_pack_key", None) != key:
            if getattr(self, "_pack_batch", None):
                lib.petsyn_pack_batch_destroy(self._pack_batch)
            n = len(convs)
            arr = lambda vals: (C.c_void_p * n)(*vals)
            for op in convs:
                op.alloc_packed()
            handle = C.c_void_p()
            check(lib.petsyn_pack_batch_create(
                n, arr([op.plan._h.value for op in convs]), arr([ptr(op.pack_source()) for op in convs]),
                arr([ptr(op.plan.w_fprop) for op in convs]),
                arr([ptr(op.plan.w_dgrad) if op.need_dx else None for op in convs]), C.byref(handle)), "pack_batch_create")
            self._pack_batch, self._pack_key = handle, key
        for op in convs:
            op.stage_padded()
        check(lib.petsyn_pack_batch_run(self._pack_batch, stream_ptr()), "pack_batch_run")

    # ``timers``: None, or a dict that receives {(op index, "fwd"|"bwd"): [(start, end) CUDA events, ...]} -- eager
    # profiling runs only (bench.py's roofline leg); events are recorded on the current stream around each op
    timers: Optional[Dict] = None

    def _timed(self, idx: int, which: str, fn) -> None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.timers.setdefault((idx, which), []).append((e0, e1))

    _stats_arena: Optional[torch.Tensor] = None
    _bwd_arena: Optional[torch.Tensor] = None

    def forward(self, training: bool) -> None:
        assert self._final
        self.repack()
        if self._stats_arena is not None:
            self._stats_arena.zero_()              # statistics sums: accumulated by the statistics passes / fused producers
        if self.timers is not None:
            for i, op in enumerate(self.ops):
                self._timed(i, "fwd", lambda: op.fwd(training))
            return
        for op in self.ops:
            op.fwd(training)

    def backward(self, on_op_done=None) -> None:
        if self._bwd_arena is not None:
            self._bwd_arena.zero_()                # bias-gradient column sums + the backward reductions of every norm
        for i in range(len(self.ops) - 1, -1, -1):
            op = self.ops[i]
            if self.timers is not None:
                self._timed(i, "bwd", op.bwd)
            else:
                op.bwd()
            if on_op_done is not None:
                on_op_done(op)
        join_side_stream()                                # weight gradients queued on the side stream

    def flops(self) -> float:
        return sum(getattr(op, "flops", 0.0) for op in self.ops)
