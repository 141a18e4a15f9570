"""Host logic of the data-parallel gradient exchange (bucket layout, launch order, averaging) on CPU with gloo,
world_size 2.  The CUDA/NCCL path uses the same GradBucketer with a side stream."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import petsyn  # noqa: F401
    from petsyn_b200.train import FlatArena, GradBucketer
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        shapes = [(3, 5), (1000,), (7, 11, 2), (4096,), (33,)]
        params = [torch.nn.Parameter(torch.randn(*s)) for s in shapes]
        before = [p.detach().clone() for p in params]
        arena = FlatArena(params, torch.device("cpu"))
        # views keep values, are 16-byte aligned and alias the arena
        for p, b, o in zip(params, before, arena.offsets):
            assert torch.equal(p.detach(), b) and o % 4 == 0
            assert p.data_ptr() == arena.p.data_ptr() + 4 * o and p.grad.data_ptr() == arena.g.data_ptr() + 4 * o
        bk = GradBucketer(arena, bucket_mb=0.004)            # ~1000 floats per bucket -> several buckets
        assert len(bk.buckets) >= 3
        assert bk.buckets[0][0] == 0 and bk.buckets[-1][1] == arena.numel
        for (s0, e0, _), (s1, e1, _) in zip(bk.buckets, bk.buckets[1:]):
            assert e0 == s1                                    # contiguous, gap-free cover of the arena
        # rank-dependent gradients, reduced bucket by bucket in production order
        for i, p in enumerate(params):
            p.grad.fill_(float(rank + 1) * (i + 1))
            bk.on_ready(p)
        bk.wait_all()
        for i, p in enumerate(params):
            expect = (i + 1) * sum(r + 1 for r in range(world)) / world
            assert torch.allclose(p.grad, torch.full_like(p.grad, expect)), (i, p.grad.flatten()[:3])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_bucketed_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results


def test_flat_arena_single_process():
    import petsyn  # noqa: F401
    from petsyn_b200.train import FlatArena, GradBucketer
    params = [torch.nn.Parameter(torch.ones(5)), torch.nn.Parameter(torch.ones(2, 3))]
    arena = FlatArena(params, torch.device("cpu"))
    assert arena.numel == 8 + 8
    bk = GradBucketer(arena, bucket_mb=32)
    assert len(bk.buckets) == 1 and bk.world == 1
    bk.on_ready(params[-1])      # world 1: no collective, nothing pending
    bk.wait_all()
