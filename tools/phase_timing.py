#!/usr/bin/env python
"""In-graph phase times of the default training step: forward + loss, backward, optimiser, each captured as its own CUDA graph
and replayed (CUDA events) -- unlike an ncu launch list these are warm-cache, back-to-back times.

    python tools/phase_timing.py [--batch 2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import petsyn  # noqa: E402
from petsyn_b200.train import AttenUNetTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = petsyn.AttenUNet(**bench.ATTEN_CFG)
    bench.redraw_parameters_(model.named_parameters(), seed=777)
    model = model.to(dev).train()
    batch = tuple(t.to(dev) for t in bench.atten_batch((96, 128, 96), 777, args.batch))
    tr = AttenUNetTrainer(model, lr=5e-4, example_input=batch[0])
    for _ in range(3):
        tr.step(*batch)
    torch.cuda.synchronize()
    phases = {
        "forward+loss": lambda: tr._forward_and_loss(*batch),
        "backward": lambda: tr.eng.backward(tr.dy, out=tr.arena.grad_views, on_ready=tr.bucketer.on_ready),
        "optimizer": lambda: tr._optimizer(),
        "step": lambda: tr._step_impl(*batch),
    }
    out = {"batch": args.batch}
    pool = torch.cuda.graph_pool_handle()
    for name, fn in phases.items():
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, pool=pool):
            fn()
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.iters):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / args.iters
    print(json.dumps(out))


if __name__ == "__main__":
    main()
