"""CPU tests of the op tape's planning (no kernel runs): which GroupNorms take their statistics from their producers, and
which per-step accumulators move into the forward / backward zero arenas (graph.Tape.finalize)."""
import torch


def _bufs(G, dev, *chs):
    return [G.Buf(2, 4, 4, 4, c, dev, f"b{i}") for i, c in enumerate(chs)]


def test_groupnorm_takes_statistics_from_residual_and_concat_producers(petsyn):
    from petsyn_b200 import graph as G, ops
    dev = torch.device("cpu")
    x, r, cat, a, b = _bufs(G, dev, 16, 16, 32, 32, 32)
    t = G.Tape()
    p1 = t.add(G.NormActOp(x, "none", ops.ACT_NONE, [cat.sl(0, 16)], res=r.sl()))      # residual sum into the concat buffer
    p2 = t.add(G.NormActOp(r, "none", ops.ACT_NONE, [cat.sl(16, 16)]))                 # skip copy into its other half
    q = t.add(G.NormActOp(cat, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 32)))
    q2 = t.add(G.NormActOp(a, "group", ops.ACT_SILU, [b.sl()], gn=torch.nn.GroupNorm(16, 32)))   # produced by a norm: own pass
    t.finalize()
    assert q.stats_from_producers and not q2.stats_from_producers
    assert p1.stats_for == [[(q, 0)]] and p2.stats_for == [[(q, 16)]]
    # the per-sample launch of the producers lands in the consumer's [sample][2][C] layout
    d = p2._desc(False)
    assert d.nsamples == 2 and d.t1_stats == q.sums.data_ptr() and d.t1_stats_c == 32 and d.t1_stats_coff == 16
    # all statistics sums live in ONE arena cleared once per forward; the backward reductions in another
    assert q2.tape_zeroes_sums and not q.tape_zeroes_sums            # q's sums are zeroed by the same fill (shared part)
    assert t._stats_arena.dtype == torch.float64           # 64-bit accumulators: exact (order-free) sums of fp32 partials
    lo, hi = t._stats_arena.data_ptr(), t._stats_arena.data_ptr() + 8 * t._stats_arena.numel()
    assert all(lo <= op.sums.data_ptr() < hi for op in (q, q2))
    assert q.tape_zeroes_bsums and q2.tape_zeroes_bsums and q._desc(True).sums_prezeroed == 1
    blo, bhi = t._bwd_arena.data_ptr(), t._bwd_arena.data_ptr() + 8 * t._bwd_arena.numel()
    assert all(blo <= op.bsums.data_ptr() < bhi and op.bsums.data_ptr() % 16 == 0 for op in (q, q2))
    # forward descriptor: finalize folded into the apply launch
    f = q._desc(False)
    assert f.fin_sums == q.sums.data_ptr() and f.fin_group_size == 2 and f.mean == q.mean.data_ptr()
    assert abs(f.fin_eps - 1e-5) < 1e-12


def test_partial_or_foreign_writers_keep_the_separate_statistics_pass(petsyn):
    from petsyn_b200 import graph as G, ops
    dev = torch.device("cpu")
    # (1) the producers cover only half of the channels
    x, cat, a = _bufs(G, dev, 16, 32, 32)
    t = G.Tape()
    p = t.add(G.NormActOp(x, "none", ops.ACT_NONE, [cat.sl(0, 16)]))
    q = t.add(G.NormActOp(cat, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 32)))
    t.finalize()
    assert not q.stats_from_producers and p.stats_for == [[]]
    # (2) a normalising op (not a plain sum / copy) also writes the buffer
    x, y, cat, a = _bufs(G, dev, 16, 16, 32, 32)
    t = G.Tape()
    t.add(G.NormActOp(x, "none", ops.ACT_NONE, [cat.sl(0, 16)]))
    t.add(G.NormActOp(y, "group", ops.ACT_SILU, [cat.sl(16, 16)], gn=torch.nn.GroupNorm(16, 16)))
    q = t.add(G.NormActOp(cat, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 32)))
    t.finalize()
    assert not q.stats_from_producers
    # (3) one destination feeds at most TWO consumers' statistics (a skip tensor: the next block and the up path)
    x, h, a, b, c = _bufs(G, dev, 16, 16, 16, 16, 16)
    t = G.Tape()
    p = t.add(G.NormActOp(x, "none", ops.ACT_NONE, [h.sl()]))
    q1 = t.add(G.NormActOp(h, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 16)))
    q2 = t.add(G.NormActOp(h, "group", ops.ACT_NONE, [b.sl()], gn=torch.nn.GroupNorm(16, 16)))
    q3 = t.add(G.NormActOp(h, "group", ops.ACT_NONE, [c.sl()], gn=torch.nn.GroupNorm(16, 16)))
    t.finalize()
    assert q1.stats_from_producers and q2.stats_from_producers and not q3.stats_from_producers
    assert p.stats_for == [[(q1, 0), (q2, 0)]]
    d = p._desc(False)
    assert d.t1_stats == q1.sums.data_ptr() and d.t2_stats == q2.sums.data_ptr() and not d.t2


def test_block_output_living_in_a_concat_slot(petsyn):
    """The AttenUNet down path: a ResnetBlock's output exists only as a channel slice of the up path's concat buffer.  The
    next block's GroupNorm reads that slice (statistics from the residual sum that wrote it, shared with the GroupNorm over
    the whole concat buffer); the residual sum has no backward pass: the conv feeding it reads its output gradient from the
    slot's gradient slice, and the identity skip's gradient joins norm1's dz as the ``extra`` addend."""
    from petsyn_b200 import graph as G, ops
    dev = torch.device("cpu")
    z, up_h, cat, a1, a_cat, prev = _bufs(G, dev, 16, 16, 32, 16, 32, 16)
    slot = cat.sl(16, 16)
    t = G.Tape()
    emit = t.add(G.NormActOp(z, "none", ops.ACT_NONE, [slot], res=prev.sl(), no_bwd=True))       # out = conv2(...) + x
    fill = t.add(G.NormActOp(up_h, "none", ops.ACT_NONE, [cat.sl(0, 16)]))                       # the up path's half
    nxt = t.add(G.NormActOp(slot, "group", ops.ACT_SILU, [a1.sl()], gn=torch.nn.GroupNorm(16, 16), extra=slot))
    whole = t.add(G.NormActOp(cat, "group", ops.ACT_SILU, [a_cat.sl()], gn=torch.nn.GroupNorm(16, 32)))
    t.finalize()
    assert nxt.stats_from_producers and whole.stats_from_producers
    assert emit.stats_for == [[(nxt, 0), (whole, 16)]] and fill.stats_for == [[(whole, 0)]]
    assert emit.grad_writes() == []                                   # no backward pass for the residual sum
    # backward order: `whole` writes the slot's gradient first (overwrite), `nxt` accumulates into the same slice
    assert not whole.acc_dz and nxt.acc_dz
    d = nxt._desc(True)
    assert (d.z_cstride, d.z_coff, d.dz_cstride, d.dz_coff) == (32, 16, 32, 16) and d.c == 16
    assert d.extra == cat.g.data_ptr() and (d.extra_cstride, d.extra_coff) == (32, 16) and d.dz_accumulate == 1


def test_engine_cache_is_least_recently_used(monkeypatch):
    """Drop-in modules keep one engine (all activation buffers of a shape) per input shape: bounded, least recently used out."""
    import petsyn
    monkeypatch.setenv("PETSYN_MAX_ENGINES", "2")
    c = petsyn.ops.EngineCache()
    c["a"], c["b"] = 1, 2
    assert c.get("a") == 1                     # refreshes "a"
    c["c"] = 3
    assert list(c) == ["a", "c"] and c.get("b") is None
    m = petsyn.UnetGenerator3d(1, 1, 4, ngf=8)
    assert isinstance(m._engines, petsyn.ops.EngineCache)
