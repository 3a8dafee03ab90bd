#!/usr/bin/env python
"""Per-layer table: cuDNN (channels_last bf16, and fp32 with TF32 as the reference runs it) against the tcgen05 implicit-GEMM
kernels, forward / data gradient / weight gradient, for every distinct convolution of FastFlow3DUNet at the benchmark's
sizes (BASELINE configs[1]: B = 16, 512 x 512).  VERDICT r01 item 6.  Measurement tool (torch's own convolutions appear here
as the comparator only).

    python tools/conv_layer_table.py [--batch 16] [--grid 512] [--out gpurun_out/conv_layer_table.txt]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deflow_b200 import conv as tc  # noqa: E402

# name, [cin per source], cout, k, stride, input resolution divisor (grid / div), count per step (layers x frames)
LAYERS = [
    ("enc1.0", [32], 64, 3, 2, 1, 2), ("enc1.1-3", [64], 64, 3, 1, 2, 6),
    ("enc2.0", [64], 128, 3, 2, 2, 2), ("enc2.1-5", [128], 128, 3, 1, 4, 10),
    ("enc3.0", [128], 256, 3, 2, 4, 2), ("enc3.1-5", [256], 256, 3, 1, 8, 10),
    ("dec1.u1", [256, 256], 256, 1, 1, 8, 1), ("dec1.u3", [128, 128], 256, 1, 1, 4, 1),
    ("dec1.u4", [256, 256], 256, 3, 1, 4, 1), ("dec1.u5", [256], 256, 3, 1, 4, 1),
    ("dec2.u1", [256], 128, 1, 1, 4, 1), ("dec2.u3", [64, 64], 128, 1, 1, 2, 1),
    ("dec2.u4", [128, 128], 128, 3, 1, 2, 1), ("dec2.u5", [128], 128, 3, 1, 2, 1),
    ("dec3.u1", [128], 64, 1, 1, 2, 1), ("dec3.u3", [32, 32], 64, 1, 1, 1, 1),
    ("dec3.u4", [64, 64], 64, 3, 1, 1, 1), ("dec3.u5+dec4", [64], 64, 3, 1, 1, 2),
]


def timeit(fn, iters=11, warm=3):
    """Median of `iters` timed launches (a mean lets one allocator / clock hiccup of a few ms into a 0.1 ms row)."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()                                   # inputs out of L2 between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True       # give cuDNN its best algorithm
    rows = []
    hdr = (f"{'layer':14s} {'x/step':>6s} {'GFLOP':>7s} | {'cuDNN bf16 fwd/dgrad/wgrad ms':>30s} | {'cuDNN TF32 fwd/dgrad/wgrad ms':>30s} | "
           f"{'tcgen05 fwd/dgrad/wgrad ms':>27s} | ours/cuDNN-bf16 (fwd, dgrad, wgrad)")
    print(hdr, flush=True)
    tot = {"cb": 0.0, "ct": 0.0, "us": 0.0}
    for name, cins, cout, k, stride, div, count in LAYERS:
        B, H = args.batch, args.grid // div
        cin = sum(cins)
        Ho = (H + 2 * (k // 2) - k) // stride + 1
        gflop = 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9
        w = (torch.randn(cout, cin, k, k, device=dev) * 0.05)
        bias = torch.zeros(cout, device=dev)
        xs = [torch.randn(B, H, H, c, device=dev).to(torch.bfloat16) for c in cins]
        gy = torch.randn(B, Ho, Ho, cout, device=dev).to(torch.bfloat16)
        res = {}
        for tag, dt in (("cb", torch.bfloat16), ("ct", torch.float32)):
            torch.backends.cudnn.allow_tf32 = True
            xc = torch.cat(xs, 3).permute(0, 3, 1, 2).to(dt).contiguous(memory_format=torch.channels_last)
            wc = w.to(dt).contiguous(memory_format=torch.channels_last)
            gc = gy.permute(0, 3, 1, 2).to(dt).contiguous(memory_format=torch.channels_last)
            bc = bias.to(dt)
            p = k // 2
            f = timeit(lambda: torch.nn.functional.conv2d(xc, wc, bc, stride, p))
            conv_bwd = torch.ops.aten.convolution_backward
            dg = timeit(lambda: conv_bwd(gc, xc, wc, [cout], [stride, stride], [p, p], [1, 1], False, [0, 0], 1, [True, False, False]))
            wg = timeit(lambda: conv_bwd(gc, xc, wc, [cout], [stride, stride], [p, p], [1, 1], False, [0, 0], 1, [False, True, False]))
            res[tag] = (f, dg, wg)
        wf, wd = tc.pack_weights(w, True, False)
        stats = torch.zeros((2, cout), dtype=torch.float64, device=dev) if name.startswith("enc") else None
        f = timeit(lambda: tc.conv2d_forward(xs, wf, bias, cout, k, stride, stats))
        if len(cins) == 2 and k == 1 and cin <= 256:
            dg = timeit(lambda: tc.conv2d_dgrad_two(gy, wd, H, H, cins[0], cins[1], cin, k, colsum=False))   # as the step calls it
        else:
            def dgrad_all():
                off = 0
                for c in cins:
                    tc.conv2d_dgrad(gy, wd, H, H, c, cin, off, k, stride)
                    off += c
            dg = timeit(dgrad_all)
        wg = timeit(lambda: tc.conv2d_wgrad(xs, gy, k, stride))
        res["us"] = (f, dg, wg)
        for t in tot:
            tot[t] += count * sum(res[t])
        fmt = lambda r: f"{r[0]:9.3f} {r[1]:9.3f} {r[2]:9.3f}"  # noqa: E731
        ratio = " ".join(f"{res['us'][i] / res['cb'][i]:6.2f}" for i in range(3))
        line = f"{name:14s} {count:6d} {gflop:7.1f} | {fmt(res['cb']):>30s} | {fmt(res['ct']):>30s} | {fmt(res['us']):>27s} | {ratio}"
        print(line, flush=True)
        rows.append(line)
    summ = (f"sum over one training step (count x (fwd + dgrad + wgrad)): cuDNN bf16 {tot['cb']:.2f} ms, cuDNN TF32 {tot['ct']:.2f} ms, "
            f"tcgen05 {tot['us']:.2f} ms   [B={args.batch}, {args.grid}^2, median of 11 launches, L2 flushed between launches, cudnn.benchmark=True, "
            f"cuDNN {torch.backends.cudnn.version()}; cuDNN rows exclude the channel concatenation and BatchNorm statistics "
            f"that the tcgen05 rows include]")
    print(summ, flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write("\n".join([hdr] + rows + [summ]) + "\n")


if __name__ == "__main__":
    main()
