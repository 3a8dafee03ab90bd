"""GPU diagnostic for the fused GRU decoder kernels: forward (inference), forward (training), backward -- against the
fp32 torch path; prints errors stage by stage so that a hang or a wrong stage is easy to locate."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deflow_b200 as d  # noqa: E402
from deflow_b200 import ops, synth  # noqa: E402


def main(stage):
    torch.manual_seed(3)
    DEV = "cuda"
    B, H, W = 2, 64, 64
    rg = [-6.4, -6.4, -3, 6.4, 6.4, 3]
    pts = synth.make_batch(2 * B, 1500, seed=31)["pc0"].clone()
    pts[..., :2] /= 8.0
    idx = ops.pillar_index(pts.half().float().to(DEV), [0.2, 0.2, 6], rg)
    n0 = idx.pt_off(B)
    head = d.ConvGRUDecoder(num_iters=4).to(DEV)
    head.apply(d.weights_init)
    img = (torch.randn(2 * B, H, W, 32, device=DEV) * 0.5).to(torch.bfloat16)
    unet = (torch.randn(B, H, W, 64, device=DEV) * 0.5).to(torch.bfloat16)
    gflow = torch.randn(n0, 3, device=DEV)
    print("points", n0, flush=True)
    head.compute_dtype = torch.float32
    head.use_library = True
    i32, u32 = img.float().requires_grad_(True), unet.float().requires_grad_(True)
    ref = head.forward_flat(i32, u32, idx, B, n0)
    ref.backward(gflow)
    ref_g = {k: p.grad.clone() for k, p in head.named_parameters()}
    head.zero_grad()
    head.use_library = False
    head.compute_dtype = torch.bfloat16
    for mode in (["fused"] if stage != "both" else ["unfused", "fused"]):
        os.environ["DFB_GRU"] = mode
        with torch.no_grad():
            f = head.forward_flat(img, unet, idx, B, n0)
        torch.cuda.synchronize()
        print(f"[{mode}] inference forward: max err {float((f - ref).abs().max()):.4g} (|ref| max {float(ref.abs().max()):.3g})", flush=True)
        if stage == "fwd":
            continue
        i2, u2 = img.clone().requires_grad_(True), unet.clone().requires_grad_(True)
        f = head.forward_flat(i2, u2, idx, B, n0)
        torch.cuda.synchronize()
        print(f"[{mode}] training forward: max err {float((f - ref).abs().max()):.4g}", flush=True)
        f.backward(gflow)
        torch.cuda.synchronize()
        rel = lambda a, b: float((a - b).abs().max()) / max(1e-9, float(b.abs().max()))  # noqa: E731
        print(f"[{mode}] grad img rel err {rel(i2.grad.float(), i32.grad):.4g}  grad unet rel err {rel(u2.grad.float(), u32.grad):.4g}", flush=True)
        for k, p in head.named_parameters():
            print(f"    {k}: rel err {rel(p.grad, ref_g[k]):.4g}", flush=True)
        head.zero_grad()


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "all")
