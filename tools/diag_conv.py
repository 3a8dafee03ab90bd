"""GPU diagnostic for the tensor-core convolution kernels: prints max/mean error of forward, dgrad and wgrad
against torch fp32 convolutions of the same bf16-rounded operands.  Usage: python tools/diag_conv.py CASE_INDEX|all"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deflow_b200 import conv  # noqa: E402

# (name, n, H, W, cins, cout, k, stride)
CASES = [
    ("1x1 64->64", 2, 32, 32, [64], 64, 1, 1),
    ("3x3 64->64", 2, 32, 32, [64], 64, 3, 1),
    ("3x3 64->128 s2", 2, 32, 32, [64], 128, 3, 2),
    ("3x3 32->64 s2 (KC=32)", 2, 64, 64, [32], 64, 3, 2),
    ("3x3 128->128", 1, 32, 24, [128], 128, 3, 1),
    ("3x3 256->256", 1, 16, 16, [256], 256, 3, 1),
    ("3x3 cat(64,64)->64", 2, 32, 32, [64, 64], 64, 3, 1),
    ("1x1 cat(256,256)->256", 1, 16, 16, [256, 256], 256, 1, 1),
    ("3x3 64->64 odd size", 1, 20, 12, [64], 64, 3, 1),
    ("3x3 128->256 s2", 1, 32, 32, [128], 256, 3, 2),
]


def run(i):
    name, n, H, W, cins, cout, k, s = CASES[i]
    torch.manual_seed(i)
    dev = "cuda"
    xs = [torch.randn(n, H, W, c, device=dev).to(torch.bfloat16) for c in cins]
    ct = sum(cins)
    w = torch.randn(cout, ct, k, k, device=dev) / (ct * k * k) ** 0.5
    b = torch.randn(cout, device=dev)
    wf, wd = conv.pack_weights(w)
    wr = w.to(torch.bfloat16).float()
    xcat = torch.cat([x.float() for x in xs], 3).permute(0, 3, 1, 2).requires_grad_(True)
    wr.requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = F.conv2d(xcat, wr, b, stride=s, padding=k // 2)
    stats = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    t0 = time.time()
    y = conv.conv2d_forward(xs, wf, b, cout, k, s, stats)
    torch.cuda.synchronize()
    refn = ref.permute(0, 2, 3, 1)
    e = (y.float() - refn).abs()
    print(f"[{i}] {name}: fwd max err {float(e.max()):.4g} mean {float(e.mean()):.4g} (ref absmax {float(refn.abs().max()):.3g}) {time.time()-t0:.2f}s", flush=True)
    es = (stats[0] - refn.double().sum((0, 1, 2))).abs().max() / max(1.0, float(refn.double().sum((0, 1, 2)).abs().max()))
    eq = (stats[1] - refn.double().square().sum((0, 1, 2))).abs().max() / float(refn.double().square().sum((0, 1, 2)).abs().max())
    print(f"     stats rel err sum {float(es):.3g} sumsq {float(eq):.3g}", flush=True)
    gy = torch.randn_like(refn).to(torch.bfloat16)
    ref.backward(gy.float().permute(0, 3, 1, 2))
    off = 0
    for x, c in zip(xs, cins):
        gx = conv.conv2d_dgrad(gy.contiguous(), wd, H, W, c, ct, off, k, s)
        torch.cuda.synchronize()
        rg = xcat.grad.permute(0, 2, 3, 1)[..., off:off + c]
        e = (gx.float() - rg).abs()
        print(f"     dgrad[{off}:{off+c}] max err {float(e.max()):.4g} mean {float(e.mean()):.4g} (ref absmax {float(rg.abs().max()):.3g})", flush=True)
        off += c
    gw = conv.conv2d_wgrad(xs, gy.contiguous(), k, s)
    torch.cuda.synchronize()
    e = (gw - wr.grad).abs()
    print(f"     wgrad max err {float(e.max()):.4g} mean {float(e.mean()):.4g} (ref absmax {float(wr.grad.abs().max()):.3g})", flush=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for i in (range(len(CASES)) if which == "all" else [int(which)]):
        run(i)
