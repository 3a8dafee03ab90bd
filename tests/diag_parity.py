"""Diagnostic: parity mode (split-precision tensor-core path) against the cuDNN/cuBLAS fp32 comparator on a golden fixture."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import deflow_b200 as d  # noqa: E402
from oracle import deflow_oracle as orc  # noqa: E402
from helpers import load_fixture, batch_to  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
name = sys.argv[1] if len(sys.argv) > 1 else "deflow_cfg1_20k"
fx, batch, cfg = load_fixture(name)
out = {}
for prec in ("fp32_library", "fp32"):
    m = d.DeFlow(cfg["voxel_size"], cfg["range"], cfg["grid"], cfg["decoder"], 4, precision=prec)
    m.load_state_dict(orc.random_state(cfg["seed_state"], cfg["decoder"]), strict=True)
    m = m.to("cuda").train(cfg["training"])
    with torch.no_grad():
        res = m(batch_to(batch, "cuda"))
    out[prec] = res
    e = np.abs(res["flow"][0].float().cpu().numpy() - fx["flow_0"]).max()
    print(f"{name} {prec}: flow error vs golden {e:.3e}")
a, b = out["fp32_library"]["_dfb"], out["fp32"]["_dfb"]
for k in ("image", "unet", "flow_flat"):
    x, y = a[k].float(), b[k].float()
    print(k, tuple(x.shape), "max abs diff", float((x - y).abs().max()), "ref max", float(x.abs().max()))
