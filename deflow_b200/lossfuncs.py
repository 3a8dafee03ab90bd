"""Loss functions with the reference's call contract (OSF/src/lossfuncs.py): ``fn(res_dict) -> {'loss': t}``.

``training_step_loss`` is the arithmetic of ``ModelWrapper.training_step`` (OSF/src/trainer.py:116-152):
gt = flow[b][idx] - pose_flow[b][idx], per-sample losses SUMMED over the batch -- in one fused launch
sequence for the whole batch when handed a ``deflow_b200.DeFlow`` result.
"""
from __future__ import annotations

import torch

from . import ops


def _single(res_dict, name):
    est, gt = res_dict["est_flow"], res_dict["gt_flow"]
    n = est.shape[0]
    idx = ops.PillarIndex()
    idx.F = 1
    idx._host = [n, 0, 0, n, 0, 0]
    idx.counts = torch.tensor(idx._host, dtype=torch.int32, device=est.device)
    idx.pt_idx = torch.arange(n, dtype=torch.int64, device=est.device)
    cls = res_dict.get("gt_classes")
    if name == "ff3dLoss":
        cls = cls.to(torch.uint8).unsqueeze(0)
    else:
        cls = None
    gt = gt.detach().float().unsqueeze(0)
    return {"loss": ops.flow_loss(est, gt, torch.zeros_like(gt), cls, idx, 1, name)}


def deflowLoss(res_dict):
    """OSF/src/lossfuncs.py:102-125."""
    return _single(res_dict, "deflowLoss")


def ff3dLoss(res_dict):
    """OSF/src/lossfuncs.py:148-157."""
    return _single(res_dict, "ff3dLoss")


def zeroflowLoss(res_dict):
    """OSF/src/lossfuncs.py:128-145 (FastFlow3DDistillationLoss of ZeroFlow)."""
    return _single(res_dict, "zeroflowLoss")


def training_step_loss(batch, res, loss_fn: str = "deflowLoss") -> torch.Tensor:
    """Sum over the batch of loss_fn({'est_flow': res.flow[b], 'gt_flow': flow[b][idx] - pose_flow[b][idx],
    'gt_classes': classes[b][idx]})  (OSF/src/trainer.py:120-142)."""
    h = res["_dfb"]
    cls = batch.get("flow_category_indices") if loss_fn == "ff3dLoss" else None
    return ops.flow_loss(h["flow_flat"], batch["flow"], h["pose_flow"], cls, h["index"], h["B"], loss_fn)
