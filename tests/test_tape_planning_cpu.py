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
    assert p1.stats_for == [(q, 0)] and p2.stats_for == [(q, 16)]
    # the per-sample launch of the producers lands in the consumer's [sample][2][C] layout
    d = p2._desc(False)
    assert d.nsamples == 2 and d.t1_stats == q.sums.data_ptr() and d.t1_stats_c == 32 and d.t1_stats_coff == 16
    # all statistics sums live in ONE arena cleared once per forward; the backward reductions in another
    assert q2.tape_zeroes_sums and not q.tape_zeroes_sums            # q's sums are zeroed by the same fill (shared part)
    lo, hi = t._stats_arena.data_ptr(), t._stats_arena.data_ptr() + 4 * t._stats_arena.numel()
    assert all(lo <= op.sums.data_ptr() < hi for op in (q, q2))
    assert q.tape_zeroes_bsums and q2.tape_zeroes_bsums and q._desc(True).sums_prezeroed == 1
    blo, bhi = t._bwd_arena.data_ptr(), t._bwd_arena.data_ptr() + 4 * t._bwd_arena.numel()
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
    assert not q.stats_from_producers and p.stats_for == [None]
    # (2) a normalising op (not a plain sum / copy) also writes the buffer
    x, y, cat, a = _bufs(G, dev, 16, 16, 32, 32)
    t = G.Tape()
    t.add(G.NormActOp(x, "none", ops.ACT_NONE, [cat.sl(0, 16)]))
    t.add(G.NormActOp(y, "group", ops.ACT_SILU, [cat.sl(16, 16)], gn=torch.nn.GroupNorm(16, 16)))
    q = t.add(G.NormActOp(cat, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 32)))
    t.finalize()
    assert not q.stats_from_producers
    # (3) one destination feeds at most one consumer's statistics
    x, h, a, b = _bufs(G, dev, 16, 16, 16, 16)
    t = G.Tape()
    p = t.add(G.NormActOp(x, "none", ops.ACT_NONE, [h.sl()]))
    q1 = t.add(G.NormActOp(h, "group", ops.ACT_SILU, [a.sl()], gn=torch.nn.GroupNorm(16, 16)))
    q2 = t.add(G.NormActOp(h, "group", ops.ACT_NONE, [b.sl()], gn=torch.nn.GroupNorm(16, 16)))
    t.finalize()
    assert q1.stats_from_producers and not q2.stats_from_producers and p.stats_for == [(q1, 0)]
