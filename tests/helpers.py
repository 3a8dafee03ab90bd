"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    fx = {k: z[k] for k in z.files}
    bsz = fx["pose0"].shape[0]
    batch = {
        "pc0": torch.from_numpy(fx["pc0"].astype(np.float32)),
        "pc1": torch.from_numpy(fx["pc1"].astype(np.float32)),
        "pose0": [torch.from_numpy(fx["pose0"][b]) for b in range(bsz)],
        "pose1": [torch.from_numpy(fx["pose1"][b]) for b in range(bsz)],
        "flow": torch.from_numpy(fx["flow_gt"]),
        "flow_category_indices": torch.from_numpy(fx["classes"]),
    }
    cfg = {
        "voxel_size": [float(v) for v in fx["cfg_voxel_size"]],
        "range": [float(v) for v in fx["cfg_range"]],
        "grid": [int(v) for v in fx["cfg_grid"]],
        "decoder": str(fx["decoder"]),
        "loss": str(fx["loss_name"]),
        "training": bool(fx["training"]),
        "seed_state": int(fx["seed_state"]),
    }
    return fx, batch, cfg


def batch_to(batch, device):
    out = {}
    for k, v in batch.items():
        out[k] = [t.to(device) for t in v] if isinstance(v, list) else v.to(device)
    return out
