#!/usr/bin/env python
"""Time one convolution layer of the hot path in isolation (CUDA events, L2 flushed between launches).

    python tools/conv_layer_bench.py --layer up1 [--pass fprop|dgrad|wgrad|all] [--iters 20] [--once]

Layers are the UnetGenerator3d(1,1,4) convs at 96x112x96, batch 1.  ``--once`` runs each pass a few
times without timing -- the form used under ``ncu --set full`` (profiles/).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import petsyn  # noqa: E402

ops = petsyn.ops
LAYERS = {
    # name: (op, n, d, h, w, cin, cout, k, s, p)   -- dims of the stored input
    "down1": (ops.OP_CONV, 1, 48, 56, 48, 128, 256, 4, 2, 1),
    "down2": (ops.OP_CONV, 1, 24, 28, 24, 256, 512, 4, 2, 1),
    "down3": (ops.OP_CONV, 1, 12, 14, 12, 512, 512, 4, 2, 1),
    "up3": (ops.OP_UPCONV, 1, 6, 7, 6, 512, 512, 3, 1, 1),
    "up2": (ops.OP_UPCONV, 1, 12, 14, 12, 1024, 256, 3, 1, 1),
    "up1": (ops.OP_UPCONV, 1, 24, 28, 24, 512, 128, 3, 1, 1),
    # AttenUNet (training.json) full- and half-resolution layers at 96x128x96, batch 2
    "a16_16": (ops.OP_CONV, 2, 96, 128, 96, 16, 16, 3, 1, 1),
    "a32_32": (ops.OP_CONV, 2, 96, 128, 96, 32, 32, 3, 1, 1),
    "a48_16": (ops.OP_CONV, 2, 96, 128, 96, 48, 16, 3, 1, 1),
    "a32_16": (ops.OP_CONV, 2, 96, 128, 96, 32, 16, 3, 1, 1),
    "h32_32": (ops.OP_CONV, 2, 48, 64, 48, 32, 32, 3, 1, 1),
    "h64_32": (ops.OP_CONV, 2, 48, 64, 48, 64, 32, 3, 1, 1),
    "h64_64": (ops.OP_CONV, 2, 48, 64, 48, 64, 64, 3, 1, 1),
    "q64_64": (ops.OP_CONV, 2, 24, 32, 24, 64, 64, 3, 1, 1),        # quarter resolution (level 2)
    "q128_64": (ops.OP_CONV, 2, 24, 32, 24, 128, 64, 3, 1, 1),
    "e128_128": (ops.OP_CONV, 2, 12, 16, 12, 128, 128, 3, 1, 1),    # eighth resolution (level 3): split-K + finish
    # BMGAN dense_unet_generator (configs[2]) at 96x128x96, batch 1: input-layer and first dense-block convolutions
    "b64_64": (ops.OP_CONV, 1, 96, 128, 96, 64, 64, 3, 1, 1),
    "b192_128": (ops.OP_CONV, 1, 48, 64, 48, 192, 128, 3, 1, 1),
    "b512_512": (ops.OP_CONV, 1, 12, 16, 12, 512, 512, 3, 1, 1),    # split-K + finish
    "b256_256": (ops.OP_CONV, 1, 24, 32, 24, 256, 256, 3, 1, 1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--layer", default="up1", choices=sorted(LAYERS))
    ap.add_argument("--pass", dest="which", default="all",
                    choices=["fprop", "dgrad", "wgrad", "all", "epi"])   # epi: the fused-epilogue variants beside the plain passes
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    op, n, d, h, w, cin, cout, k, s, p = LAYERS[args.layer]
    n = args.batch or n
    dev = torch.device("cuda:0")
    plan = ops.ConvPlan(op, n, d, h, w, cin, cout, k, s, p)
    g = torch.Generator().manual_seed(0)
    wt = (torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).to(dev)
    plan.pack(wt)
    x = torch.randn(n, d, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
    od, oh, ow = plan.out_dims
    y = torch.empty(n, od, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    dy = torch.randn(n, od, oh, ow, cout, generator=g).to(dev).to(torch.bfloat16)
    dx = torch.empty_like(x)
    dw = torch.empty_like(wt)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2
    passes = {"fprop": lambda: plan.fprop(x, y), "dgrad": lambda: plan.dgrad(dy, dx),
              "wgrad": lambda: plan.wgrad(x, dy, dw)}
    if args.which == "epi":
        from petsyn_b200._cabi import ConvEpilogue, check, lib, ptr, stream_ptr
        assert plan.epi_stats_ok, "no fused epilogue for this layer's forward pass"
        res = torch.randn(n, od, oh, ow, cout, generator=g).to(dev).to(torch.bfloat16)
        st = torch.zeros(n, 2, cout, dtype=torch.float64, device=dev)
        e_st, e_rs, e_r = ConvEpilogue(), ConvEpilogue(), ConvEpilogue()
        e_st.stats1, e_st.stats1_c = ptr(st), cout
        e_rs.side, e_rs.side_cstride, e_rs.add_side, e_rs.stats1, e_rs.stats1_c = ptr(res), cout, 1, ptr(st), cout
        e_r.side, e_r.side_cstride, e_r.add_side = ptr(res), cout, 1
        if not plan.epi_ok[0]:         # gather-form kernel (or its split-K finish pass): statistics targets only
            passes = {"fprop": passes["fprop"], "fprop+stats": lambda: plan.fprop_epi(x, y, None, e_st),
                      "stats_pass": lambda: check(lib.petsyn_norm_stats(ptr(y), ptr(st), od * oh * ow, cout, n, stream_ptr()))}
        else:
            passes = {"fprop": passes["fprop"], "fprop+stats": lambda: plan.fprop_epi(x, y, None, e_st),
                      "fprop+res": lambda: plan.fprop_epi(x, y, None, e_r),
                      "fprop+res+stats": lambda: plan.fprop_epi(x, y, None, e_rs),
                      "stats_pass": lambda: check(lib.petsyn_norm_stats(ptr(y), ptr(st), od * oh * ow, cout, n, stream_ptr())),
                      "dgrad": passes["dgrad"]}
        if plan.epi_ok[1]:
            z = torch.randn(n, d, h, w, cin, generator=g).to(dev).to(torch.bfloat16)
            f = lambda: torch.rand(n, cin, generator=g).to(dev) + 0.5
            sc, sh, mu, rs = f(), f(), f(), f()
            bs = torch.zeros(n, 2, cin, dtype=torch.float64, device=dev)
            e_n = ConvEpilogue()
            e_n.side, e_n.side_cstride = ptr(z), cin
            e_n.norm_scale, e_n.norm_shift, e_n.norm_mean, e_n.norm_rstd = ptr(sc), ptr(sh), ptr(mu), ptr(rs)
            e_n.norm_act, e_n.bsums = ops.ACT_SILU, ptr(bs)
            passes["dgrad+normreduce"] = lambda: plan.dgrad_epi(dy, dx, e_n)
    todo = list(passes) if args.which in ("all", "epi") else [args.which]
    out = {"layer": args.layer, "batch": n, "flops_algorithmic": plan.flops_algorithmic,
           "flops_executed": plan.flops_executed}
    for name in todo:
        fn = passes[name]
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if args.once:
            continue
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        io_bytes = 2.0 * n * d * h * w * cin + 2.0 * n * od * oh * ow * cout     # bf16 read-once + write-once
        out[name] = {"ms_median": med, "ms_min": ts[0], "tflops_executed": plan.flops_executed / med / 1e9,
                     "tflops_algorithmic": plan.flops_algorithmic / med / 1e9, "hbm_gbs_min": io_bytes / med / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
