#!/usr/bin/env python
"""Time the GroupNorm + SiLU kernels (statistics, apply, backward reduce / apply) on one AttenUNet-sized tensor.

    python tools/norm_bench.py [--c 16] [--iters 10]
CUDA events around each C-ABI call, L2 flushed between launches; prints achieved GB/s against algorithmic bytes.
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import petsyn  # noqa: E402,F401
from petsyn_b200 import graph as G, ops  # noqa: E402
from petsyn_b200._cabi import check, lib, ptr, stream_ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--shape", type=int, nargs=4, default=[2, 96, 128, 96])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    n, d, h, w = args.shape
    c = args.c
    z = G.Buf(n, d, h, w, c, dev, "z")
    a = G.Buf(n, d, h, w, c, dev, "a")
    z.t.copy_(torch.randn_like(z.t, dtype=torch.float32))
    a.g.copy_(torch.randn_like(a.t, dtype=torch.float32))
    gn = torch.nn.GroupNorm(16, c).to(dev)
    op = G.NormActOp(z, "group", ops.ACT_SILU, [a.sl()], gn=gn)
    op.grad_gamma = torch.zeros(c, device=dev)
    op.grad_beta = torch.zeros(c, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    nbytes = z.t.numel() * 2

    def timed(fn):
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    op.fwd(True); op.bwd(); torch.cuda.synchronize()
    out = {"c": c, "tensor_mb": nbytes / 1e6}
    t = timed(lambda: check(lib.petsyn_norm_stats(ptr(z.t), ptr(op.sums), op.rows, c, n, stream_ptr())))
    out["stats"] = {"us": t * 1e3, "gbs": nbytes / t / 1e6}
    d_f = op._desc(False)
    t = timed(lambda: check(lib.petsyn_normact_fwd(C.byref(d_f), stream_ptr())))
    out["fwd_apply"] = {"us": t * 1e3, "gbs": 2 * nbytes / t / 1e6}
    d_b = op._desc(True)
    t = timed(lambda: check(lib.petsyn_normact_bwd(C.byref(d_b), stream_ptr())))
    out["bwd_reduce+apply"] = {"us": t * 1e3, "gbs": 5 * nbytes / t / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
