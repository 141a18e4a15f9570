#!/usr/bin/env python
"""Per-op CUDA-event times of one AttenUNet training step (eager), sorted: which layer costs what.

    python tools/atten_op_profile.py [--top 40]
"""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
import petsyn  # noqa: E402
from petsyn_b200 import graph as G  # noqa: E402
from petsyn_b200.train import AttenUNetTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--top", type=int, default=40)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    model = petsyn.AttenUNet(**bench.ATTEN_CFG)
    bench.redraw_parameters_(model.named_parameters(), seed=777)
    model = model.to(dev).train()
    batch = tuple(t.to(dev) for t in bench.atten_batch((96, 128, 96), 777, 2))
    trainer = AttenUNetTrainer(model, lr=5e-4, example_input=batch[0])
    for _ in range(2):
        trainer.step(*batch)
    tape = trainer.eng.tape
    tape.timers = {}
    for _ in range(4):
        trainer.step(*batch)
    torch.cuda.synchronize()
    rows = []
    for (idx, which), ev in tape.timers.items():
        op = tape.ops[idx]
        ms = statistics.mean(a.elapsed_time(b) for a, b in ev[1:])
        desc = type(op).__name__
        if isinstance(op, G.ConvOp):
            b = op.x.buf
            desc += f" {op.name} {op.cin}->{op.cout} k{op.plan.desc.ksize} op{op.opcode} @{b.d}x{b.h}x{b.w} path{op.plan.kernel_path}"
        elif isinstance(op, G.NormActOp):
            desc += f" {op.name} {op.kind} C={op.z.c} rows={op.z.rows} act={op.act} dsts={len(op.dsts)} res={op.res is not None}"
        rows.append((ms, which, desc))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total {tot:.3f} ms over {len(rows)} op passes")
    for ms, which, desc in rows[:args.top]:
        print(f"{ms * 1e3:9.1f} us  {which}  {desc}")


if __name__ == "__main__":
    main()
