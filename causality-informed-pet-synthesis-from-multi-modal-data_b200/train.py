"""Train-step orchestration for the generators (the hot loop of ``unet/scripts/train_unet.py:129-168`` restricted to
the terms that exist offline: zero_grad -> forward -> nn.L1Loss -> backward -> Adam), plus the data-parallel gradient
exchange that ``DistributedDataParallel`` performs implicitly in the reference (train_unet.py:41,72).

Design (one process per GPU, torch.distributed for plumbing):
  * all fp32 master parameters live in ONE flat arena, all gradients in another, laid out in the order backward
    produces them; ``nn.Parameter.data`` / ``.grad`` are views, so ``state_dict``/checkpoints are unchanged;
  * the arena is cut into ~32 MB buckets at parameter boundaries; as soon as backward has enqueued the last kernel
    of a bucket, an event is recorded and ``all_reduce(AVG)`` of that bucket is launched on a side stream, so NCCL
    (NVLink 5 / NVSwitch) overlaps the rest of backward -- the same schedule DDP's reducer runs, without autograd;
  * Adam is one fused kernel over the whole arena.
BatchNorm uses per-rank batch statistics exactly like the reference (no SyncBatchNorm); the reference's DDP
``broadcast_buffers`` re-broadcast of running statistics from rank 0 is reproduced by ``sync_buffers()``.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import ops


class FlatArena:
    """Flat fp32 parameter + gradient storage with per-parameter views."""

    def __init__(self, params: List[torch.nn.Parameter], device):
        self.params = params
        self.offsets: List[int] = []
        off = 0
        for p in params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4          # keep every view 16-byte aligned
        self.numel = off
        self.p = torch.zeros(off, dtype=torch.float32, device=device)
        self.g = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad_views: Dict[int, torch.Tensor] = {}
        for p, o in zip(params, self.offsets):
            view = self.p[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            gv = self.g[o:o + p.numel()].view_as(p)
            p.grad = gv
            self.grad_views[id(p)] = gv


class GradBucketer:
    """Bucketed gradient all-reduce over a FlatArena, launched as buckets complete (device-agnostic: NCCL on CUDA
    with a side stream, gloo on CPU for the host-logic tests)."""

    def __init__(self, arena: FlatArena, bucket_mb: float = 32.0, process_group=None, comm_dtype=torch.float32):
        """``comm_dtype``: torch.float32 -- the reference's DistributedDataParallel (fp32 all-reduce of fp32 gradients) -- or
        torch.bfloat16: every bucket is cast to bf16 on the communication stream, averaged, and cast back into the fp32 arena
        (half the NVLink bytes; the average of bf16-rounded gradients instead of the fp32 average: an explicit opt-in)."""
        self.arena = arena
        self.pg = process_group
        self.comm_dtype = comm_dtype
        self.comm_buf = None
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.buckets: List[tuple] = []          # (start, end, id(last param))
        limit = max(1, int(bucket_mb * (1 << 20) / 4))
        start = 0
        last = arena.params[-1]
        for p, o in zip(arena.params, arena.offsets):
            end = o + (p.numel() + 3) // 4 * 4
            if end - start >= limit or p is last:
                self.buckets.append((start, end, id(p)))
                start = end
        self._bucket_of_last = {b[2]: i for i, b in enumerate(self.buckets)}
        self.cuda = arena.g.is_cuda
        self.comm_stream = torch.cuda.Stream(device=arena.g.device) if (self.cuda and self.world > 1) else None
        if comm_dtype != torch.float32 and self.cuda and self.world > 1:
            self.comm_buf = torch.empty(arena.g.numel(), dtype=comm_dtype, device=arena.g.device)
        self._pending: List = []
        self._avg = self.cuda          # NCCL supports ReduceOp.AVG; gloo does not

    def on_ready(self, p) -> None:
        """Call right after the kernels producing ``p.grad`` have been enqueued (gradient-production order)."""
        i = self._bucket_of_last.get(id(p))
        if i is None or self.world == 1:
            return
        start, end, _ = self.buckets[i]
        view = self.arena.g[start:end]
        op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
        if self.cuda:
            from .graph import join_side_stream
            join_side_stream(self.arena.g.device)      # weight gradients of this bucket may still run on the side stream
            ev = torch.cuda.Event()
            ev.record()                                # compute stream: bucket i is complete after this point
            self.comm_stream.wait_event(ev)
            with torch.cuda.stream(self.comm_stream):
                if self.comm_buf is not None:
                    cb = self.comm_buf[start:end]
                    cb.copy_(view)                     # fp32 -> bf16 on the communication stream
                    work = dist.all_reduce(cb, op=op, group=self.pg, async_op=True)
                    work.wait()                        # the communication stream waits for the collective ...
                    view.copy_(cb)                     # ... and casts the average back into the fp32 gradient arena
                    done = torch.cuda.Event()
                    done.record(self.comm_stream)
                    self._pending.append((None, view, done))
                    return
                work = dist.all_reduce(view, op=op, group=self.pg, async_op=True)
        else:
            work = dist.all_reduce(view, op=op, group=self.pg, async_op=True)
        self._pending.append((work, view, None))

    def wait_all(self) -> None:
        for work, view, done in self._pending:
            if done is not None:
                torch.cuda.current_stream().wait_event(done)
                continue
            work.wait()                                # CUDA: the current stream waits for the collective
            if not self._avg:
                view.div_(self.world)
        self._pending.clear()


class GraphSegments:
    """A training step captured as CUDA-graph SEGMENTS with eager actions (NCCL collectives) in between.

    Collectives stay outside the graphs: while a step function is being captured, ``cut(action)`` closes the open graph,
    remembers ``action`` and opens the next one; ``replay()`` then alternates ``graph.replay()`` and ``action()``.  The same
    step function therefore serves the eager path (actions run inline) and the captured one."""

    def __init__(self):
        self.items: List[tuple] = []
        self.pool = torch.cuda.graph_pool_handle()
        self._g = self._ctx = None

    def open(self) -> None:
        self._g = torch.cuda.CUDAGraph()
        self._ctx = torch.cuda.graph(self._g, pool=self.pool)
        self._ctx.__enter__()

    def cut(self, action) -> None:
        from .graph import join_side_stream
        join_side_stream()                           # a capture may only end with its forked weight-gradient branch joined
        self._ctx.__exit__(None, None, None)
        self.items.append((self._g, action))
        self.open()

    def finish(self) -> None:
        self._ctx.__exit__(None, None, None)
        self.items.append((self._g, None))
        self._g = self._ctx = None

    def replay(self) -> None:
        for g, action in self.items:
            g.replay()
            if action is not None:
                action()


def adam_state_dict(params: List[torch.nn.Parameter], exp_avg: List[torch.Tensor], exp_avg_sq: List[torch.Tensor],
                    step: int, lr: float, betas, eps: float) -> dict:
    """``torch.optim.Adam(params, lr).state_dict()`` for externally held moments: what the reference stores under
    ``'g_optimizer'`` (train_unet.py:297-302) and feeds back through ``g_optimizer.load_state_dict`` (:109).  ``params`` in
    ``model.parameters()`` order (the order ``Adam(unet.parameters())`` indexes its state by)."""
    groups = torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=lr, betas=tuple(betas), eps=eps).state_dict()[
        "param_groups"]
    groups[0]["params"] = list(range(len(params)))
    state = {}
    if step > 0:                                   # torch creates per-parameter state lazily at the first step
        for i, (m, v) in enumerate(zip(exp_avg, exp_avg_sq)):
            state[i] = {"step": torch.tensor(float(step)), "exp_avg": m.detach().clone(),
                        "exp_avg_sq": v.detach().clone()}
    return {"state": state, "param_groups": groups}


def read_adam_state_dict(sd: dict, exp_avg: List[torch.Tensor], exp_avg_sq: List[torch.Tensor]) -> int:
    """Inverse of ``adam_state_dict``: copies the moments into the given views and returns the step count."""
    idx = sd["param_groups"][0]["params"]
    if len(idx) != len(exp_avg):
        raise ValueError(f"optimizer state holds {len(idx)} parameters, the model has {len(exp_avg)}")
    step = 0
    for k, (m, v) in zip(idx, zip(exp_avg, exp_avg_sq)):
        st = sd["state"].get(k)
        if st is None:
            m.zero_(); v.zero_()
            continue
        if tuple(st["exp_avg"].shape) != tuple(m.shape):
            raise ValueError(f"optimizer state {k}: shape {tuple(st['exp_avg'].shape)} != parameter {tuple(m.shape)}")
        m.copy_(st["exp_avg"]); v.copy_(st["exp_avg_sq"])
        step = max(step, int(float(st["step"])))
    return step


class _CheckpointMixin:
    """The reference's checkpoint dictionary (train_unet.py:85-109 resume, :297-302 save):
    ``{'unet', 'discriminator', 'epoch', 'g_optimizer', 'eval_loss'}`` with ``state_dict()`` s as values.  Under DDP the
    reference's keys carry a ``module.`` prefix: ``load_checkpoint`` accepts both forms, ``checkpoint(ddp_prefix=...)`` writes
    either (default: prefixed when this trainer is data parallel), so checkpoints move in both directions."""

    def _moments(self):
        off = {id(p): o for p, o in zip(self.arena.params, self.arena.offsets)}
        ps = list(self.model.parameters())
        m = [self.m[off[id(p)]:off[id(p)] + p.numel()].view_as(p) for p in ps]
        v = [self.v[off[id(p)]:off[id(p)] + p.numel()].view_as(p) for p in ps]
        return ps, m, v

    def optimizer_state_dict(self) -> dict:
        ps, m, v = self._moments()
        return adam_state_dict(ps, m, v, int(self.step_dev.item()), self.lr, self.betas, self.eps)

    def load_optimizer_state_dict(self, sd: dict) -> None:
        g = sd["param_groups"][0]
        hyper = (float(g["lr"]), tuple(g["betas"]), float(g["eps"]))
        captured = getattr(self, "graph", None) is not None or getattr(self, "graphs", None) is not None
        if captured and hyper != (self.lr, tuple(self.betas), self.eps):
            raise RuntimeError("lr / betas / eps are baked into the captured step: load the optimizer state before capture()")
        _, m, v = self._moments()
        self.step_dev.fill_(read_adam_state_dict(sd, m, v))
        self.lr, self.betas, self.eps = hyper

    def checkpoint(self, epoch: int, eval_loss: float = float("nan"), discriminator=None, model_key: str = "unet",
                   ddp_prefix: Optional[bool] = None) -> dict:
        """``ddp_prefix``: write the state-dict keys as ``module.<key>`` -- what the reference saves and strictly re-loads when its
        networks are wrapped in DistributedDataParallel (train_unet.py:72-74,85-88).  Default: on when this trainer runs data
        parallel.  The reference's resume also indexes ``ckpt['discriminator']`` unconditionally: pass the discriminator
        (this trainer's own when it was built with one) for a checkpoint the reference script can resume from."""
        if ddp_prefix is None:
            ddp_prefix = getattr(self, "world", 1) > 1
        pre = (lambda sd: {"module." + k: v for k, v in sd.items()}) if ddp_prefix else (lambda sd: dict(sd))
        out = {model_key: pre(self.model.state_dict()), "epoch": epoch, "g_optimizer": self.optimizer_state_dict(),
               "eval_loss": eval_loss}
        if discriminator is None:
            discriminator = getattr(self, "disc", None)
        if discriminator is not None:
            out["discriminator"] = pre(discriminator.state_dict())
        return out

    def save_checkpoint(self, path: str, epoch: int, eval_loss: float = float("nan"), discriminator=None,
                        ddp_prefix: Optional[bool] = None) -> None:
        torch.save(self.checkpoint(epoch, eval_loss, discriminator, ddp_prefix=ddp_prefix), path)

    def load_checkpoint(self, ckpt, discriminator=None, model_key: str = "unet") -> int:
        """``ckpt``: a path or an already loaded dictionary.  Returns the epoch to resume from (train_unet.py:89)."""
        if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
            ckpt = torch.load(ckpt, map_location=self.dev, weights_only=False)
        strip = lambda sd: {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        self.model.load_state_dict(strip(ckpt[model_key]))       # parameters are views of the arena: copied in place
        if discriminator is not None and "discriminator" in ckpt:
            discriminator.load_state_dict(strip(ckpt["discriminator"]))
        if "g_optimizer" in ckpt:
            self.load_optimizer_state_dict(ckpt["g_optimizer"])
        self.eng.mark_weights_dirty()
        return int(ckpt["epoch"]) + 1


class Unet3dTrainer(_CheckpointMixin):
    """Fused training step for a petsyn ``UnetGenerator3d``; data-parallel when a process group is initialised."""

    def __init__(self, model, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, bucket_mb: float = 32.0,
                 process_group=None, example_input: Optional[torch.Tensor] = None):
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = next(model.parameters()).device
        self.dev = dev
        if example_input is None:
            raise ValueError("example_input (a tensor of the training shape) is required to lay out the arenas")
        if not model.default_family():
            raise NotImplementedError("the fused train_unet step covers the reference configuration (norm_layer=nn.BatchNorm3d, "
                                      "no dropout layer, unet/train_unet.py:59); other UnetGenerator3d families train through "
                                      "module.forward / loss.backward() and any torch optimiser")
        eng = model.engine_for(example_input)
        self.eng = eng
        order = eng.grad_order()
        assert {id(p) for p in order} == {id(p) for p in model.parameters()}
        self.arena = FlatArena(order, dev)
        self.m = torch.zeros_like(self.arena.p)
        self.v = torch.zeros_like(self.arena.p)
        self.step_count = 0
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.graphs = None
        self.opt_graph = None
        self.static_x = self.static_t = None
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dy = torch.empty(example_input.shape, dtype=torch.float32, device=dev)
        eng.mark_weights_dirty()
        self.bucketer = GradBucketer(self.arena, bucket_mb, process_group)
        if self.world > 1:
            self.sync_parameters()

    # ------------------------------------------------------------------------------------------------ collectives
    def sync_parameters(self) -> None:
        """DDP construction semantics: every rank starts from rank 0's parameters and buffers."""
        dist.broadcast(self.arena.p, src=0, group=self.pg)
        self.sync_buffers()
        self.eng.mark_weights_dirty()

    def sync_buffers(self) -> None:
        """DDP ``broadcast_buffers=True`` (train_unet.py:72): BatchNorm running statistics follow rank 0.  DDP re-broadcasts
        them before every forward; in training mode nothing reads them, and rank 0's own copy evolves from rank 0's batches
        alone either way, so broadcasting when they are about to be READ -- before evaluation on every rank -- leaves every rank
        with exactly the statistics DDP would have left there.  A collective: call it on ALL ranks (rank 0's own checkpoint,
        written by rank 0 alone as in train_unet.py:296-302, already holds the authoritative statistics)."""
        if self.world == 1:
            return
        bufs = [b for b in self.model.buffers() if b.dtype.is_floating_point]
        if not bufs:
            return
        flat = torch.cat([b.reshape(-1) for b in bufs])            # one collective instead of one per buffer
        dist.broadcast(flat, src=0, group=self.pg)
        off = 0
        for b in bufs:
            b.copy_(flat[off:off + b.numel()].view_as(b))
            off += b.numel()

    # ------------------------------------------------------------------------------------------------ step
    def _forward_and_loss(self, x: torch.Tensor, target: torch.Tensor, timers=None) -> None:
        y = self.eng.forward(x, save=True, timers=timers, clone_output=False)
        self.loss.zero_()
        ops.l1_loss_fwd_bwd(y, target, self.loss, self.dy)

    def _optimizer(self) -> None:
        self.step_dev.add_(1)
        ops.adam_step(self.arena.p, self.arena.g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, 0,
                      step_dev=self.step_dev)
        self.eng.mark_weights_dirty()

    def _step_impl(self, x: torch.Tensor, target: torch.Tensor, timers=None) -> torch.Tensor:
        self._forward_and_loss(x, target, timers)
        self.eng.backward(self.dy, out=self.arena.grad_views, on_ready=self.bucketer.on_ready, timers=timers)
        self.bucketer.wait_all()
        self._optimizer()
        return self.loss

    def step(self, x: torch.Tensor, target: torch.Tensor, timers=None) -> torch.Tensor:
        """One optimisation step (zero_grad -> fwd -> L1 -> bwd -> [all-reduce] -> Adam); returns the device loss
        tensor of this rank's micro-batch.  Replays the captured CUDA graph(s) when ``capture()`` has been called."""
        self.step_count += 1
        if self.graphs is None or timers is not None:
            return self._step_impl(x, target, timers)
        if x.data_ptr() != self.static_x.data_ptr():
            self.static_x.copy_(x, non_blocking=True)
        if target.data_ptr() != self.static_t.data_ptr():
            self.static_t.copy_(target, non_blocking=True)
        for graph, last_param in self.graphs:
            graph.replay()
            if last_param is not None:
                self.bucketer.on_ready(last_param)       # eager NCCL launch between graph segments
        if self.world > 1:
            self.bucketer.wait_all()
            self.opt_graph.replay()
        return self.loss

    @property
    def graph(self):
        return self.graphs

    def capture(self, warmup: int = 3) -> None:
        """Capture the training step into CUDA graphs (the step is ~100 short kernels; launching them one by one from
        Python costs more host time than the GPU needs to run them).

        Single GPU: one graph for the whole step.  Data parallel: the collectives stay outside the graphs -- the step
        is cut at every gradient-bucket boundary into [fwd + loss + backward-until-bucket-0] [..until bucket 1] ...
        [Adam]; after each segment the bucket's NCCL all-reduce is launched on the side stream, overlapping the next
        segment.  Parameters, optimiser state and BatchNorm buffers are restored after the warm-up steps capture
        needs, so capture() has no side effect on training state.  Write inputs into ``static_x`` / ``static_t`` to
        skip the device-to-device input copy."""
        snap = [t.clone() for t in (self.arena.p, self.m, self.v, self.step_dev)]
        bufs = [b.clone() for b in self.model.buffers()]
        count = self.step_count
        self.static_x = torch.zeros(self.dy.shape, dtype=torch.float32, device=self.dev)
        self.static_t = torch.zeros(self.dy.shape, dtype=torch.float32, device=self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(self.static_x, self.static_t)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        graphs = []
        pool = torch.cuda.graph_pool_handle()
        if self.world == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._step_impl(self.static_x, self.static_t)
            graphs.append((g, None))
            self.opt_graph = None
        else:
            closers = set(self.bucketer._bucket_of_last)
            it = self.eng.backward_iter(self.dy, self.arena.grad_views)
            g = torch.cuda.CUDAGraph()
            ctx = torch.cuda.graph(g, pool=pool)
            ctx.__enter__()
            self._forward_and_loss(self.static_x, self.static_t)
            for prm in it:
                if id(prm) in closers:
                    ctx.__exit__(None, None, None)
                    graphs.append((g, prm))
                    g = torch.cuda.CUDAGraph()
                    ctx = torch.cuda.graph(g, pool=pool)
                    ctx.__enter__()
            # the last parameter always closes the last bucket, so the open capture only holds trailing kernels (none)
            self._optimizer()
            ctx.__exit__(None, None, None)
            self.opt_graph = g
        for dst, src in zip((self.arena.p, self.m, self.v, self.step_dev), snap):
            dst.copy_(src)
        for dst, src in zip(self.model.buffers(), bufs):
            dst.copy_(src)
        self.step_count = count
        self.eng.mark_weights_dirty()
        self.graphs = graphs

    def grad_norm(self) -> float:
        out = torch.zeros(1, dtype=torch.float32, device=self.dev)
        ops.sumsq(self.arena.g, out)
        return float(out.item()) ** 0.5


class BmganTrainer:
    """One BMGAN "adversarial step" as written in ``bl_methods/BMGAN/train_bmgan.py:141-200``, minus the
    part that cannot exist offline (LPIPS needs downloaded weights):

      G phase (:141-161)  fake = G(t1, z); loss = LSGAN(D(fake), real) + lamda_l1 * L1(fake, pet); D frozen;
                          backward through D (data gradient only) into G; [bucketed all-reduce]; Adam on G.
      E phase (:163-180)  (when an encoder is given) G forward again under no_grad, mu/logvar = E(real), E(fake);
                          loss = mean(KL(real) + KL(fake)); Adam on E.
      D phase (:183-200)  G forward a third time (the reference recomputes it in every phase), then
                          LSGAN(D(fake), fake) and LSGAN(D(real), real): two backward calls whose weight gradients
                          accumulate.  The reference never calls ``d_optimizer.step()``: reproduced
                          "as written" -- D stays at its initialisation; set ``step_discriminator=True`` for the
                          evidently intended behaviour.
    Per-GPU batch 1 (train_bmgan.py:315); data parallel exactly like Unet3dTrainer (flat arenas, bucketed all-reduce of
    G's gradients overlapped with backward, per-rank normalisation statistics).
    """

    def __init__(self, gen, disc, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, lamda_l1: float = 20.0,
                 bucket_mb: float = 256.0, process_group=None, example_input: Optional[torch.Tensor] = None,
                 step_discriminator: bool = False, enc=None, grad_comm_dtype=torch.float32):
        if example_input is None:
            raise ValueError("example_input (a tensor of the training shape) is required to lay out the arenas")
        self.gen, self.disc = gen, disc
        self.lr, self.betas, self.eps, self.lamda_l1 = lr, betas, eps, lamda_l1
        self.step_discriminator = step_discriminator
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = example_input.device
        self.dev = dev
        self.geng = gen.engine_for(example_input)
        self.deng = disc.engine_for(example_input)
        self.garena = FlatArena(self.geng.grad_order(), dev)
        self.darena = FlatArena(list(self.deng.params), dev)
        self.gm, self.gv = torch.zeros_like(self.garena.p), torch.zeros_like(self.garena.p)
        self.dm, self.dv = torch.zeros_like(self.darena.p), torch.zeros_like(self.darena.p)
        self.enc = enc
        if enc is not None:
            self.eeng = enc.engine_for(example_input)
            self.earena = FlatArena(list(self.eeng.params), dev)
            self.em, self.ev = torch.zeros_like(self.earena.p), torch.zeros_like(self.earena.p)
            self.loss_kl = torch.zeros(1, dtype=torch.float32, device=dev)
            self.dlatent = torch.zeros(example_input.shape[0], 16, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        f = lambda: torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_adv, self.loss_l1, self.loss_d_fake, self.loss_d_real = f(), f(), f(), f()
        self.dy = torch.zeros(example_input.shape, dtype=torch.float32, device=dev)
        self.dlogits = torch.zeros_like(self.deng.logits)
        self.bucketer = GradBucketer(self.garena, bucket_mb, process_group, comm_dtype=grad_comm_dtype)
        self.graph = None
        self._cap: Optional[GraphSegments] = None      # set while capture() records the step
        self.static = None
        if self.world > 1:
            dist.broadcast(self.garena.p, src=0, group=self.pg)
            dist.broadcast(self.darena.p, src=0, group=self.pg)
            if enc is not None:
                dist.broadcast(self.earena.p, src=0, group=self.pg)
            for b in list(gen.buffers()) + list(disc.buffers()):
                if b.dtype.is_floating_point:
                    dist.broadcast(b, src=0, group=self.pg)
        self._dirty()

    def _dirty(self) -> None:
        for op in self.geng.tape.ops:
            if hasattr(op, "_ver"):
                op._ver = None

    def _collective(self, fn) -> None:
        """Run a collective now (eager step) or -- while the step is being captured -- end the open graph segment and let
        ``fn`` run between the segments at replay time."""
        if self._cap is not None:
            self._cap.cut(fn)
        else:
            fn()

    def _bucket_ready(self, p) -> None:
        if self.world > 1 and id(p) in self.bucketer._bucket_of_last:
            self._collective(lambda: self.bucketer.on_ready(p))

    def _step_impl(self, t1: torch.Tensor, pet: torch.Tensor, z: torch.Tensor) -> None:
        geng, deng = self.geng, self.deng
        # ---------------- G phase ----------------
        fake = geng.forward(t1, z)
        logits = deng.forward(fake)
        self.loss_adv.zero_()
        ops.mse_const_fwd_bwd(logits, 1.0, self.loss_adv, self.dlogits)
        dfake, _ = deng.backward(self.dlogits, need_dx=True, out=self.darena.grad_views, need_dw=False)
        self.loss_l1.zero_()
        ops.l1_loss_fwd_bwd(fake, pet, self.loss_l1, self.dy, grad_scale=self.lamda_l1)
        self.dy.add_(dfake)
        geng.backward(self.dy, out=self.garena.grad_views, on_ready=self._bucket_ready)
        if self.world > 1:
            self._collective(self.bucketer.wait_all)
        self.step_dev.add_(1)
        ops.adam_step(self.garena.p, self.garena.g, self.gm, self.gv, self.lr, self.betas[0], self.betas[1], self.eps, 0,
                      step_dev=self.step_dev)
        self._dirty()
        # ---------------- E phase ----------------
        if self.enc is not None:
            eeng = self.eeng
            fake = geng.forward(t1, z)                               # train_bmgan.py:168-169 (no_grad recompute)
            self.loss_kl.zero_()
            n = t1.shape[0]
            for i, vol in enumerate((pet, fake)):
                lat = eeng.forward(vol)                              # [n, 16] = [mu | logvar]
                ops.kl_fwd_bwd(lat, lat[:, 8:], self.loss_kl, self.dlatent, self.dlatent[:, 8:], n, 8, 16)
                eeng.backward(self.dlatent, out=self.earena.grad_views, accumulate=(i == 1))
            if self.world > 1:
                self._collective(lambda: dist.all_reduce(self.earena.g, op=dist.ReduceOp.AVG, group=self.pg))
            ops.adam_step(self.earena.p, self.earena.g, self.em, self.ev, self.lr, self.betas[0], self.betas[1], self.eps,
                          0, step_dev=self.step_dev)
            for op in eeng.tape.ops:
                if hasattr(op, "_ver"):
                    op._ver = None
        # ---------------- D phase ----------------
        fake = geng.forward(t1, z)                                   # recomputed again, as the reference does
        logits = deng.forward(fake)
        self.loss_d_fake.zero_()
        ops.mse_const_fwd_bwd(logits, 0.0, self.loss_d_fake, self.dlogits)
        deng.backward(self.dlogits, need_dx=False, out=self.darena.grad_views, need_dw=True,
                      accumulate=not self.step_discriminator)
        logits = deng.forward(pet)
        self.loss_d_real.zero_()
        ops.mse_const_fwd_bwd(logits, 1.0, self.loss_d_real, self.dlogits)
        deng.backward(self.dlogits, need_dx=False, out=self.darena.grad_views, need_dw=True, accumulate=True)
        if self.step_discriminator:
            if self.world > 1:
                self._collective(lambda: dist.all_reduce(self.darena.g, op=dist.ReduceOp.AVG, group=self.pg))
            ops.adam_step(self.darena.p, self.darena.g, self.dm, self.dv, self.lr, self.betas[0], self.betas[1], self.eps,
                          0, step_dev=self.step_dev)
            for op in self.deng.tape.ops:
                if hasattr(op, "_ver"):
                    op._ver = None

    # ------------------------------------------------------------------------------------------------ checkpoints
    def _nets(self):
        """(checkpoint key, optimizer key, module, engine, arena, m, v, optimiser steps taken) per network (train_bmgan.py:296-302)."""
        steps = int(self.step_dev.item())
        out = [("generator", "g_optimizer", self.gen, self.geng, self.garena, self.gm, self.gv, steps),
               ("discriminator", "d_optimizer", self.disc, self.deng, self.darena, self.dm, self.dv,
                steps if self.step_discriminator else 0)]
        if self.enc is not None:
            out.append(("encoder", "e_optimizer", self.enc, self.eeng, self.earena, self.em, self.ev, steps))
        return out

    @staticmethod
    def _moment_views(module, arena, m, v):
        off = {id(p): o for p, o in zip(arena.params, arena.offsets)}
        ps = list(module.parameters())
        return ps, [m[off[id(p)]:off[id(p)] + p.numel()].view_as(p) for p in ps], \
            [v[off[id(p)]:off[id(p)] + p.numel()].view_as(p) for p in ps]

    def checkpoint(self, epoch: int, ddp_prefix: Optional[bool] = None) -> dict:
        """The dictionary train_bmgan.py:296-302 saves every ``save_every`` epochs and :96-108 resumes from: ``{'generator',
        'discriminator', 'encoder', 'epoch', 'g_optimizer', 'd_optimizer', 'e_optimizer'}``, the optimizer entries being
        ``torch.optim.Adam.state_dict()`` s over ``module.parameters()``.  The reference never steps ``d_optimizer``
        (SURVEY 9 Q4), so its state is empty unless ``step_discriminator`` is set.  ``ddp_prefix`` as in
        ``Unet3dTrainer.checkpoint`` (train_bmgan.py:72-75 wraps the three networks in DistributedDataParallel)."""
        if ddp_prefix is None:
            ddp_prefix = self.world > 1
        pre = (lambda sd: {"module." + k: v for k, v in sd.items()}) if ddp_prefix else (lambda sd: dict(sd))
        out = {"epoch": epoch}
        for key, okey, module, _, arena, m, v, steps in self._nets():
            out[key] = pre(module.state_dict())
            ps, mv, vv = self._moment_views(module, arena, m, v)
            out[okey] = adam_state_dict(ps, mv, vv, steps, self.lr, self.betas, self.eps)
        return out

    def save_checkpoint(self, path: str, epoch: int, ddp_prefix: Optional[bool] = None) -> None:
        torch.save(self.checkpoint(epoch, ddp_prefix=ddp_prefix), path)

    def load_checkpoint(self, ckpt) -> int:
        """``ckpt``: a path or a loaded dictionary in the layout above (either key form; ``best.ckpt``, train_bmgan.py:282-288,
        carries no optimizer entries: the moments are then left as they are).  Returns the epoch to resume from (:100)."""
        if isinstance(ckpt, (str, bytes)) or hasattr(ckpt, "__fspath__"):
            ckpt = torch.load(ckpt, map_location=self.dev, weights_only=False)
        strip = lambda sd: {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
        steps = None
        for key, okey, module, eng, arena, m, v, _ in self._nets():
            module.load_state_dict(strip(ckpt[key]))                 # parameters are views of the arena: copied in place
            if okey in ckpt:
                g = ckpt[okey]["param_groups"][0]
                hyper = (float(g["lr"]), tuple(g["betas"]), float(g["eps"]))
                if self.graph is not None and hyper != (self.lr, tuple(self.betas), self.eps):
                    raise RuntimeError("lr / betas / eps are baked into the captured step: load the checkpoint before capture()")
                _, mv, vv = self._moment_views(module, arena, m, v)
                n = read_adam_state_dict(ckpt[okey], mv, vv)
                if okey == "g_optimizer":
                    steps = n
                    self.lr, self.betas, self.eps = hyper
        if steps is not None:
            self.step_dev.fill_(steps)
            self.step_count = steps
        self._dirty()
        for eng in (self.deng,) + ((self.eeng,) if self.enc is not None else ()):
            eng.mark_weights_dirty()
        return int(ckpt["epoch"]) + 1

    def step(self, t1: torch.Tensor, pet: torch.Tensor, z: torch.Tensor):
        """Returns the device tensors (adv, l1, d_fake, d_real) of this rank's micro-batch."""
        self.step_count += 1
        if self.graph is not None:
            for dst, src in zip(self.static, (t1, pet, z)):
                if dst.data_ptr() != src.data_ptr():
                    dst.copy_(src, non_blocking=True)
            self.graph.replay()
        else:
            self._step_impl(t1, pet, z)
        return self.loss_adv, self.loss_l1, self.loss_d_fake, self.loss_d_real

    def capture(self, warmup: int = 2) -> None:
        """The whole adversarial step as CUDA graph(s) (state is restored after the warm-up).  One GPU: a single graph.
        Data parallel: graph SEGMENTS cut wherever a collective happens -- after the last gradient of every generator bucket
        (its all-reduce is launched on the side stream and overlaps the next segment), before the generator's Adam (wait),
        around the encoder's / discriminator's gradient all-reduce."""
        state = [self.garena.p, self.gm, self.gv, self.darena.p, self.darena.g, self.dm, self.dv, self.step_dev]
        if self.enc is not None:
            state += [self.earena.p, self.em, self.ev]
        snap = [t.clone() for t in state]
        bufs = [b.clone() for b in list(self.gen.buffers()) + list(self.disc.buffers())]
        n = self.dy.shape[0]
        self.static = (torch.zeros_like(self.dy), torch.zeros_like(self.dy),
                       torch.zeros(n, self.gen.cfg["input_channel"] - 1, dtype=torch.float32, device=self.dev))
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(*self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = GraphSegments()
        g.open()
        self._cap = g
        try:
            self._step_impl(*self.static)
        finally:
            self._cap = None
        g.finish()
        for dst, src in zip(state, snap):
            dst.copy_(src)
        for dst, src in zip(list(self.gen.buffers()) + list(self.disc.buffers()), bufs):
            dst.copy_(src)
        # The captured step starts with a generator forward that does NOT repack (its weights were packed by the last
        # forward of the previous step); after restoring the pre-warm-up weights the packed images must be refreshed
        # eagerly once, or the first replay would run on the warm-up's weights.
        self._dirty()
        self.geng.tape.repack()
        if self.enc is not None:
            for op in self.eeng.tape.ops:
                if hasattr(op, "_ver"):
                    op._ver = None
            self.eeng.tape.repack()
        self.graph = g


class AttenUNetTrainer(_CheckpointMixin):
    """Fused training step for the covariate-conditioned generator, ``unet/scripts/train_unet.py:136-195``:

      G phase (:136-168)  D frozen; ``output_pet = unet(t1, condition)``; ``g_loss = L1 + adv_weight * LSGAN(D(output_pet)[-1],
                          real)`` (LPIPS has weight 0 in unet/config/training.json:55 and needs downloaded weights: dropped);
                          backward through D (data gradient only) into the generator; [bucketed all-reduce]; Adam(base_lr).
      D phase (:171-193)  (``adv_weight > 0`` and a discriminator given) the generator forward AGAIN under no_grad -- with the
                          weights the G phase just updated, as the reference does -- then ``LSGAN(D(fake), fake).backward()``
                          and ``LSGAN(D(real), real).backward()`` (two backward calls, gradients accumulate; the 0.5 *
                          adv_weight scaling of ``d_loss`` is only logged, never backpropagated), Adam(disc_lr) on D.  The
                          reference's DDP all-reduces D's gradients once per backward call; here once, after both.
    Without a discriminator (or ``adv_weight == 0``) the step is zero_grad -> unet -> L1 -> backward -> Adam, exactly the
    reference's ``else`` branches (:156-157,194-195).  Data-parallel like Unet3dTrainer; after ``capture()`` the step replays
    as CUDA graph(s)."""

    def __init__(self, model, lr: float = 5e-4, betas=(0.9, 0.999), eps: float = 1e-8, bucket_mb: float = 32.0,
                 process_group=None, example_input: Optional[torch.Tensor] = None, ssim_weight: float = 0.0,
                 discriminator=None, adv_weight: float = 0.0, disc_lr: float = 1e-4):
        """``ssim_weight`` > 0 adds ``ssim_weight * (1 - SSIM)`` (Gaussian window 5, sigma 0.5, data_range 1 -- the
        parameters of the reference's evaluation, output_predict.py:73) to the L1 reconstruction loss; the reference's own
        training loss is L1 (+ LPIPS / adversarial terms), so the default is 0.  ``discriminator``: a petsyn
        ``PatchDiscriminator(**training.json["discriminator"])``; ``adv_weight`` / ``disc_lr``: training.json:53-56."""
        if example_input is None:
            raise ValueError("example_input (a tensor of the training shape) is required to lay out the arenas")
        self.model = model
        self.lr, self.betas, self.eps = lr, betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        dev = example_input.device
        self.dev = dev
        self.eng = model.engine_for(example_input)
        order = self.eng.grad_order()
        assert {id(p) for p in order} == {id(p) for p in model.parameters()}
        self.arena = FlatArena(order, dev)
        self.m, self.v = torch.zeros_like(self.arena.p), torch.zeros_like(self.arena.p)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        self.loss = torch.zeros(1, dtype=torch.float32, device=dev)
        self.dy = torch.zeros(example_input.shape, dtype=torch.float32, device=dev)
        self.ssim_weight = float(ssim_weight)
        self.ssim = ops.SsimLoss(tuple(example_input.shape), dev) if self.ssim_weight > 0 else None
        self.ssim_value: Optional[torch.Tensor] = None      # mean SSIM of the last step (device scalar)
        self.bucketer = GradBucketer(self.arena, bucket_mb, process_group)
        self.graph = None
        self.segments = None
        self.d_graph = None
        self.static = None
        self.disc = discriminator if (discriminator is not None and adv_weight > 0) else None
        self.adv_weight, self.disc_lr = float(adv_weight), float(disc_lr)
        if self.disc is not None:
            self.deng = self.disc.engine_for(example_input)
            self.darena = FlatArena(list(self.deng.params), dev)
            self.dm, self.dv = torch.zeros_like(self.darena.p), torch.zeros_like(self.darena.p)
            self.d_step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            f = lambda: torch.zeros(1, dtype=torch.float32, device=dev)
            self.loss_adv, self.loss_d_fake, self.loss_d_real = f(), f(), f()
            self.dlogits = torch.zeros_like(self.deng.logits)
            self.deng.mark_weights_dirty()
        if self.world > 1:
            dist.broadcast(self.arena.p, src=0, group=self.pg)
            if self.disc is not None:
                dist.broadcast(self.darena.p, src=0, group=self.pg)
                for b in self.disc.buffers():
                    if b.dtype.is_floating_point:
                        dist.broadcast(b, src=0, group=self.pg)
        self.eng.mark_weights_dirty()

    def _forward_and_loss(self, x: torch.Tensor, context: torch.Tensor, target: torch.Tensor) -> None:
        y = self.eng.forward(x, context)
        self.loss.zero_()
        ops.l1_loss_fwd_bwd(y, target, self.loss, self.dy)
        if self.ssim is not None:
            self.ssim_value = self.ssim(y, target, self.dy, grad_scale=self.ssim_weight, accumulate=True)
        if self.disc is not None:                          # adv_weight * LSGAN(D(output_pet)[-1], real): train_unet.py:153-155
            logits = self.deng.forward(y)
            self.loss_adv.zero_()
            ops.mse_const_fwd_bwd(logits, 1.0, self.loss_adv, self.dlogits, grad_scale=self.adv_weight)
            dfake, _ = self.deng.backward(self.dlogits, need_dx=True, out=self.darena.grad_views, need_dw=False)
            self.dy.add_(dfake)

    def _d_phase_grads(self, x: torch.Tensor, context: torch.Tensor, target: torch.Tensor) -> None:
        """train_unet.py:171-184: the two backward calls of the discriminator phase (gradients into the D arena)."""
        fake = self.eng.forward(x, context)                # :175-176, no_grad recompute with the updated generator
        logits = self.deng.forward(fake)
        self.loss_d_fake.zero_()
        ops.mse_const_fwd_bwd(logits, 0.0, self.loss_d_fake, self.dlogits)
        self.deng.backward(self.dlogits, need_dx=False, out=self.darena.grad_views, need_dw=True, accumulate=False)
        logits = self.deng.forward(target)
        self.loss_d_real.zero_()
        ops.mse_const_fwd_bwd(logits, 1.0, self.loss_d_real, self.dlogits)
        self.deng.backward(self.dlogits, need_dx=False, out=self.darena.grad_views, need_dw=True, accumulate=True)

    def _d_optimizer(self) -> None:
        self.d_step_dev.add_(1)
        ops.adam_step(self.darena.p, self.darena.g, self.dm, self.dv, self.disc_lr, self.betas[0], self.betas[1], self.eps, 0,
                      step_dev=self.d_step_dev)
        self.deng.mark_weights_dirty()

    def _d_phase(self, x: torch.Tensor, context: torch.Tensor, target: torch.Tensor) -> None:
        self._d_phase_grads(x, context, target)
        if self.world > 1:
            dist.all_reduce(self.darena.g, op=dist.ReduceOp.AVG, group=self.pg)
        self._d_optimizer()

    def _optimizer(self) -> None:
        self.step_dev.add_(1)
        ops.adam_step(self.arena.p, self.arena.g, self.m, self.v, self.lr, self.betas[0], self.betas[1], self.eps, 0,
                      step_dev=self.step_dev)
        self.eng.mark_weights_dirty()

    def _step_impl(self, x: torch.Tensor, context: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        self._forward_and_loss(x, context, target)
        self.eng.backward(self.dy, out=self.arena.grad_views, on_ready=self.bucketer.on_ready)
        self.bucketer.wait_all()
        self._optimizer()
        if self.disc is not None:
            self._d_phase(x, context, target)
        return self.loss

    def step(self, x: torch.Tensor, context: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One step; returns the device L1 loss of this rank's micro-batch (``loss_adv`` / ``loss_d_fake`` / ``loss_d_real``
        hold the adversarial terms when a discriminator is trained along)."""
        self.step_count += 1
        ctx = context.reshape(x.shape[0], -1)
        if self.graph is None:
            return self._step_impl(x, ctx.contiguous().float(), target)
        for dst, src in zip(self.static, (x, ctx, target)):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        if self.segments is None:
            self.graph.replay()
            return self.loss
        for graph, last_param in self.segments:
            graph.replay()
            self.bucketer.on_ready(last_param)       # eager NCCL launch between graph segments (side stream)
        self.bucketer.wait_all()
        self.graph.replay()                          # Adam [+ the D phase's forward / backward calls]
        if self.disc is not None:
            dist.all_reduce(self.darena.g, op=dist.ReduceOp.AVG, group=self.pg)
            self.d_graph.replay()                    # Adam on D
        return self.loss

    def capture(self, warmup: int = 2) -> None:
        """Capture the step into CUDA graphs.  Single GPU: one graph.  Data parallel: the collectives stay outside the
        graphs -- the step is cut at every gradient-bucket boundary into [fwd + loss + backward-until-bucket-0]
        [..until bucket 1] ... [Adam]; each bucket's all-reduce is launched eagerly on the side stream right after
        its segment and overlaps the next one (same scheme as Unet3dTrainer.capture)."""
        state = [self.arena.p, self.m, self.v, self.step_dev]
        bufs = []
        if self.disc is not None:
            state += [self.darena.p, self.dm, self.dv, self.d_step_dev]
            bufs = [b for b in self.disc.buffers()]
        buf_snap = [b.clone() for b in bufs]
        snap = [t.clone() for t in state]
        n = self.dy.shape[0]
        cdim = self.model.cfg["cross_attention_dim"]
        self.static = (torch.zeros_like(self.dy), torch.zeros(n, cdim, dtype=torch.float32, device=self.dev),
                       torch.zeros_like(self.dy))
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_impl(*self.static)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(self.dev)
        pool = torch.cuda.graph_pool_handle()
        if self.world == 1:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._step_impl(*self.static)
            self.segments = None
        else:
            closers = set(self.bucketer._bucket_of_last)
            segments = []
            cur = {"g": torch.cuda.CUDAGraph()}
            cur["ctx"] = torch.cuda.graph(cur["g"], pool=pool)
            cur["ctx"].__enter__()

            def cut(prm) -> None:                    # called right after the op that completes prm's gradient
                if id(prm) not in closers:
                    return
                from .graph import join_side_stream
                join_side_stream()                   # a capture may only end with its forked weight-gradient branch joined
                cur["ctx"].__exit__(None, None, None)
                segments.append((cur["g"], prm))
                cur["g"] = torch.cuda.CUDAGraph()
                cur["ctx"] = torch.cuda.graph(cur["g"], pool=pool)
                cur["ctx"].__enter__()

            self._forward_and_loss(*self.static)
            self.eng.backward(self.dy, out=self.arena.grad_views, on_ready=cut)
            self._optimizer()                        # the open capture holds only what follows the last bucket: Adam
            if self.disc is not None:
                self._d_phase_grads(*self.static)    # ... and the D phase up to its gradients (their all-reduce is eager)
            cur["ctx"].__exit__(None, None, None)
            g = cur["g"]
            self.segments = segments
            if self.disc is not None:
                self.d_graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.d_graph, pool=pool):
                    self._d_optimizer()
        for dst, src in zip(state, snap):
            dst.copy_(src)
        for dst, src in zip(bufs, buf_snap):
            dst.copy_(src)
        self.eng.mark_weights_dirty()        # without a discriminator the captured step begins with the repack
        if self.disc is not None:
            # with one, every step ENDS with a generator forward (the D phase's recompute) that leaves the packed weights
            # current, so the captured step does not begin with a repack: refresh them once for the restored weights
            self.eng.tape.repack()
            self.deng.mark_weights_dirty()   # D is re-packed inside the graph (its Adam is the last thing a step does)
        self.graph = g
