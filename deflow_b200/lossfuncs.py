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
    The nearest-neighbour searches -- the native op of this loss (chamfer3D.cu) -- run on csrc/chamfer.cu; the per-cluster
    bookkeeping is the reference's torch-level glue (data-dependent label loop), kept as it is there."""
    from .chamfer3D import nnChamferDis
    cham = nnChamferDis()
    pc0_label, pc1_label = res_dict["pc0_labels"], res_dict["pc1_labels"]
    pc0, pc1, est_flow = res_dict["pc0"], res_dict["pc1"], res_dict["est_flow"]
    pseudo_pc1from0 = pc0 + est_flow
    unique_labels = torch.unique(pc0_label)
    pc0_dynamic = pc0[pc0_label > 0]
    pc1_dynamic = pc1[pc1_label > 0]
    have_dynamic_cluster = (pc0_dynamic.shape[0] > 256) & (pc1_dynamic.shape[0] > 256)
    est_dist0, est_dist1, _, _ = cham.disid_res(pseudo_pc1from0, pc1)
    raw_dist0, raw_dist1, raw_idx0, _ = cham.disid_res(pc0, pc1)
    chamfer_dis = torch.mean(est_dist0[est_dist0 <= TRUNCATED_DIST]) + torch.mean(est_dist1[est_dist1 <= TRUNCATED_DIST])
    dynamic_chamfer_dis = torch.tensor(0.0, device=est_flow.device)
    if have_dynamic_cluster:
        dynamic_chamfer_dis = dynamic_chamfer_dis + cham(pseudo_pc1from0[pc0_label > 0], pc1_dynamic, truncate_dist=TRUNCATED_DIST)
    static_cluster_loss = torch.tensor(0.0, device=est_flow.device)
    moved_cluster_loss = torch.tensor(0.0, device=est_flow.device)
    moved_cluster_norms = []
    raw_idx0 = raw_idx0.long()
    for label in unique_labels:
        mask = pc0_label == label
        if label == 0:
            static_cluster_loss = static_cluster_loss + torch.linalg.vector_norm(est_flow[mask, :], dim=-1).mean()   # Eq. 6
        elif label > 0 and have_dynamic_cluster:
            cluster_id_flow = est_flow[mask, :]
            cluster_nnd = raw_dist0[mask]
            if cluster_nnd.shape[0] <= 0:
                continue
            sorted_idxs = torch.argsort(cluster_nnd, descending=True)                                                # Eq. 8
            nearby_label = pc1_label[raw_idx0[mask][sorted_idxs]]
            non_zero_valid_indices = torch.nonzero(nearby_label > 0)
            if non_zero_valid_indices.shape[0] <= 0:
                continue
            max_idx = sorted_idxs[non_zero_valid_indices.squeeze(1)[0]]
            max_flow = pc1[raw_idx0[mask][max_idx]] - pc0[mask][max_idx]                                             # Eq. 9
            moved_cluster_norms.append(torch.linalg.vector_norm(cluster_id_flow - max_flow, dim=-1))                 # Eq. 10
    if moved_cluster_norms:
        moved_cluster_loss = torch.cat(moved_cluster_norms).mean()                                                   # Eq. 11
    elif have_dynamic_cluster:
        moved_cluster_loss = torch.mean(raw_dist0[raw_dist0 <= TRUNCATED_DIST]) + torch.mean(raw_dist1[raw_dist1 <= TRUNCATED_DIST])
    return {"chamfer_dis": chamfer_dis, "dynamic_chamfer_dis": dynamic_chamfer_dis,
            "static_flow_loss": static_cluster_loss, "cluster_based_pc0pc1": moved_cluster_loss}


def training_step_loss(batch, res, loss_fn: str = "deflowLoss") -> torch.Tensor:
    """Sum over the batch of loss_fn({'est_flow': res.flow[b], 'gt_flow': flow[b][idx] - pose_flow[b][idx],
    'gt_classes': classes[b][idx]})  (OSF/src/trainer.py:120-142)."""
    h = res["_dfb"]
    cls = batch.get("flow_category_indices") if loss_fn == "ff3dLoss" else None
    return ops.flow_loss(h["flow_flat"], batch["flow"], h["pose_flow"], cls, h["index"], h["B"], loss_fn)
