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


TRUNCATED_DIST = 4   # OSF/src/lossfuncs.py:19


def seflowLoss(res_dict, timer=None):
    """Self-supervised SeFlow loss (OSF/src/lossfuncs.py:22-100): chamfer, dynamic chamfer, static-flow and cluster terms.

    The nearest-neighbour searches -- the native op of this loss (chamfer3D.cu) -- run on csrc/chamfer.cu.  The cluster term
    (Eq. 8-11; a Python loop over the labels with an argsort, a nonzero and three boolean-index kernels per cluster in the
    reference, lossfuncs.py:63-86) is evaluated for all clusters at once with segment reductions: per cluster the point with
    the largest nearest-neighbour distance whose neighbour in pc1 is dynamic (amax over the cluster, then the lowest point
    index attaining it), its displacement as the cluster's rigid flow, and the mean deviation of the estimated flow from it
    over all points of clusters that have such a point."""
    from .chamfer3D import nnChamferDis
    cham = nnChamferDis()
    l0, l1 = res_dict["pc0_labels"].long(), res_dict["pc1_labels"].long()
    pc0, pc1, est = res_dict["pc0"], res_dict["pc1"], res_dict["est_flow"]
    dev = est.device
    zero = torch.zeros((), device=dev)
    warped = pc0 + est
    dyn0, dyn1 = l0 > 0, l1 > 0
    have_dyn = int(dyn0.sum()) > 256 and int(dyn1.sum()) > 256          # lossfuncs.py:35-38 (one host read)
    e0, e1, _, _ = cham.disid_res(warped, pc1)
    r0, r1, nn0, _ = cham.disid_res(pc0, pc1)
    out = {"chamfer_dis": e0[e0 <= TRUNCATED_DIST].mean() + e1[e1 <= TRUNCATED_DIST].mean()}
    out["dynamic_chamfer_dis"] = zero + cham(warped[dyn0], pc1[dyn1], truncate_dist=TRUNCATED_DIST) if have_dyn else zero
    # Eq. 6: static points should not move (a NaN mean when there is no static point, as in the reference)
    static = ~dyn0 & (l0 == 0)
    out["static_flow_loss"] = zero + torch.linalg.vector_norm(est[static], dim=-1).mean() if bool((l0 == 0).any()) else zero
    moved = zero
    if have_dyn:
        nn0 = nn0.long()
        labels, seg = torch.unique(l0, return_inverse=True)             # seg[p] = cluster slot of point p
        K = labels.shape[0]
        cand = dyn0 & (l1[nn0] > 0)                                     # the neighbour in pc1 is dynamic too (Eq. 8)
        key = torch.where(cand, r0.detach(), torch.full_like(r0, -1.0))
        best = torch.full((K,), -1.0, device=dev).scatter_reduce(0, seg, key, "amax", include_self=True)
        hit = cand & (key == best[seg])
        n = l0.shape[0]
        pick = torch.full((K,), n, device=dev, dtype=torch.long).scatter_reduce(
            0, seg, torch.where(hit, torch.arange(n, device=dev), torch.full((n,), n, device=dev)), "amin", include_self=True)
        has = pick < n                                                  # clusters with a usable point
        safe = pick.clamp(max=n - 1)
        rigid = pc1[nn0[safe]] - pc0[safe]                              # Eq. 9: that point's displacement, per cluster
        use = has[seg] & dyn0
        if bool(use.any()):
            moved = torch.linalg.vector_norm(est[use] - rigid[seg[use]].detach(), dim=-1).mean()      # Eq. 10-11
        else:
            moved = r0[r0 <= TRUNCATED_DIST].mean() + r1[r1 <= TRUNCATED_DIST].mean()                 # lossfuncs.py:90-91
    out["cluster_based_pc0pc1"] = moved
    return out


def training_step_loss(batch, res, loss_fn: str = "deflowLoss") -> torch.Tensor:
    """Sum over the batch of loss_fn({'est_flow': res.flow[b], 'gt_flow': flow[b][idx] - pose_flow[b][idx],
    'gt_classes': classes[b][idx]})  (OSF/src/trainer.py:120-142)."""
    h = res["_dfb"]
    cls = batch.get("flow_category_indices") if loss_fn == "ff3dLoss" else None
    return ops.flow_loss(h["flow_flat"], batch["flow"], h["pose_flow"], cls, h["index"], h["B"], loss_fn)
