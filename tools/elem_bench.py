"""Microbenchmark (diagnostic): the HBM-bound UNet passes at the layer sizes of BASELINE configs[1] (16 frames of a
512x512 grid per encoder call), GB/s on the bytes each pass must move.  usage: python tools/elem_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from deflow_b200 import conv  # noqa: E402

dev = "cuda"


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # > L2: every timed call starts cold
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    print(f"{'pass':28s} {'shape':22s} {'ms':>8s} {'GB/s':>8s}")
    for (n, H, W, C) in [(16, 256, 256, 64), (16, 128, 128, 128), (16, 64, 64, 256), (16, 512, 512, 64)]:
        x = torch.randn(n, H, W, C, device=dev).to(torch.bfloat16)
        gy = torch.randn(n, H, W, C, device=dev).to(torch.bfloat16)
        bn = torch.rand(4, C, device=dev) + 0.5
        gg, gb, gbias = (torch.zeros(C, device=dev) for _ in range(3))
        nbytes = x.numel() * 2
        rows = [("bn_gelu_apply", lambda: conv.bn_gelu_apply(x, bn), 2 * nbytes),
                ("bn_gelu_backward(2 passes)", lambda: conv.bn_gelu_backward(x, gy, bn, True, gg, gb, gbias), 5 * nbytes),
                ("channel_sum", lambda: conv.channel_sum(gy), nbytes)]
        if H <= 256:
            rows.append(("upsample2x", lambda: conv.upsample2x(x), 5 * nbytes))
            g4 = torch.randn(n, 2 * H, 2 * W, C, device=dev).to(torch.bfloat16)
            rows.append(("upsample2x_bwd", lambda: conv.upsample2x(g4, backward=True), 5 * nbytes))
        for name, fn, by in rows:
            ms = timeit(fn)
            print(f"{name:28s} {str((n, H, W, C)):22s} {ms:8.4f} {by / ms / 1e6:8.0f}")


if __name__ == "__main__":
    main()
