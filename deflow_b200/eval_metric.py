"""Evaluation metrics of the validation step, on the device.

The reference evaluates every validation frame on the host: ``evaluate_leaderboard`` (EPE three-way + dynamic IoU),
``evaluate_leaderboard_v2`` (bucketed class x speed EPE) and ``evaluate_ssf`` (range-wise EPE) build boolean masks in
torch, copy every per-point tensor to the CPU and reduce in float64 numpy (OSF/src/utils/eval_metric.py:28-106,
OSF/src/utils/av2_eval.py:460-553, 839-915; called from ``ModelWrapper.train_validation_step_`` / ``eval_only_step_``,
OSF/src/trainer.py:154-171, 224-266).  Here one kernel pass per frame accumulates all three families into an
814-double record (csrc/eval_metric.cu); ``OfficialMetrics`` keeps one record per frame on the device and reads them
back once, in ``normalize()``.  Same accumulation rules and the same result fields as the reference's
``OfficialMetrics`` (eval_metric.py:235-345): ``epe_3way``, ``bucketed``, ``epe_ssf``.

Class tables: ``av2`` (av2==0.2.1, OSF/environment.yaml:35) defines the 30 annotation categories;
CATEGORY_TO_INDEX = {NONE: 0, category_i: i + 1} (av2_eval.py:32-35).
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import EvalTables, check
from .ops import _stream

ANNOTATION_CATEGORIES = [
    "ANIMAL", "ARTICULATED_BUS", "BICYCLE", "BICYCLIST", "BOLLARD", "BOX_TRUCK", "BUS", "CONSTRUCTION_BARREL",
    "CONSTRUCTION_CONE", "DOG", "LARGE_VEHICLE", "MESSAGE_BOARD_TRAILER", "MOBILE_PEDESTRIAN_CROSSING_SIGN", "MOTORCYCLE",
    "MOTORCYCLIST", "OFFICIAL_SIGNALER", "PEDESTRIAN", "RAILED_VEHICLE", "REGULAR_VEHICLE", "SCHOOL_BUS", "SIGN",
    "STOP_SIGN", "STROLLER", "TRAFFIC_LIGHT_TRAILER", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER", "WHEELCHAIR",
    "WHEELED_DEVICE", "WHEELED_RIDER"]
CATEGORY_TO_INDEX = {"NONE": 0, **{k: i + 1 for i, k in enumerate(ANNOTATION_CATEGORIES)}}
BUCKETED_METACATAGORIES = {          # av2_eval.py:47-75
    "BACKGROUND": ["NONE"],
    "CAR": ["REGULAR_VEHICLE"],
    "PEDESTRIAN": ["PEDESTRIAN", "STROLLER", "WHEELCHAIR", "OFFICIAL_SIGNALER"],
    "WHEELED_VRU": ["BICYCLE", "BICYCLIST", "MOTORCYCLE", "MOTORCYCLIST", "WHEELED_DEVICE", "WHEELED_RIDER"],
    "OTHER_VEHICLES": ["BOX_TRUCK", "LARGE_VEHICLE", "RAILED_VEHICLE", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER",
                       "ARTICULATED_BUS", "BUS", "SCHOOL_BUS"]}
CLASS_NAMES = ["BACKGROUND", "CAR", "OTHER_VEHICLES", "PEDESTRIAN", "WHEELED_VRU"]       # eval_metric.py:262
SPEED_SPLITS = np.concatenate([np.linspace(0, 2.0, 51), [np.inf]])                       # av2_eval.py:848
DISTANCE_SPLIT = [0, 35, 50, 75, 100, np.inf]                                             # av2_eval.py:892
ACC = 814
_V2, _SSF = 19, 784


def default_tables() -> EvalTables:
    t = EvalTables()
    for i in range(256):
        t.fg_bg[i] = 0 if i == 0 else (1 if i <= 30 else 255)      # FOREGROUND_BACKGROUND_BREAKDOWN, av2_eval.py:217-229
        t.meta[i] = 255
    for name, cats in BUCKETED_METACATAGORIES.items():
        for c in cats:
            t.meta[CATEGORY_TO_INDEX[c]] = CLASS_NAMES.index(name)
    for i, v in enumerate(SPEED_SPLITS):
        t.speed_splits[i] = float(v)
    for i, v in enumerate(DISTANCE_SPLIT):
        t.dist_splits[i] = float(v)
    t.n_speed, t.n_dist = len(SPEED_SPLITS) - 1, len(DISTANCE_SPLIT) - 1
    return t


_TABLES = None


def accumulate_frame(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids, acc: Optional[torch.Tensor] = None):
    """One frame -> its 814-double metric record on the device (added to ``acc`` when given).  Argument meaning as
    ``evaluate_leaderboard(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids)`` (eval_metric.py:28)."""
    global _TABLES
    if not est_flow.is_cuda:
        raise RuntimeError("deflow_b200.eval_metric runs on CUDA only; there is no CPU path")
    if _TABLES is None:
        _TABLES = default_tables()
    n = est_flow.shape[0]
    f32 = lambda t: t.detach().to(torch.float32).contiguous()  # noqa: E731
    est, rig, gt, pc = f32(est_flow), f32(rigid_flow), f32(gt_flow), f32(pc0)
    val = is_valid.detach().to(torch.uint8).contiguous()
    ids = pts_ids.detach().to(torch.uint8).contiguous()
    assert est.shape == (n, 3) and rig.shape == (n, 3) and gt.shape == (n, 3) and pc.shape[0] == n and pc.shape[1] >= 3
    if acc is None:
        acc = torch.zeros(ACC, dtype=torch.float64, device=est.device)
    check(_lib.lib().dfb_eval_accumulate(est.data_ptr(), rig.data_ptr(), pc.data_ptr(), pc.shape[1], gt.data_ptr(),
                                         val.data_ptr(), ids.data_ptr(), n, C.byref(_TABLES), acc.data_ptr(), _stream(est)),
          "eval_accumulate")
    return acc


# ---------------------------------------------------------------------------------------------- host-side finishing
def _three_way(a: np.ndarray) -> Dict[str, float]:
    """compute_metrics' tail (av2_eval.py:540-553) from one frame's record."""
    cnt, s = a[0:8], a[8:16]

    def cepe(idx):
        c = cnt[idx].sum()
        return float(s[idx].sum() / (c + 1e-8)) if c != 0 else 0.0
    tp, fp, fn = a[16:19]
    return {"EPE_BS": cepe([2, 3]), "EPE_FD": cepe([4, 5]), "EPE_FS": cepe([6, 7]), "IoU": float(tp / (tp + fp + fn + 1e-6))}


def evaluate_leaderboard(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> Dict[str, float]:
    """eval_metric.py:28-54 (reads 19 doubles back; use OfficialMetrics.step_frame to avoid the per-frame sync)."""
    return _three_way(accumulate_frame(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids).cpu().numpy())


def evaluate_leaderboard_v2(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> List[tuple]:
    """eval_metric.py:57-78 -> [(class, avg_epe, avg_speed, (lo, hi), count)] for the non-empty buckets."""
    a = accumulate_frame(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids).cpu().numpy()
    m = a[_V2:_SSF].reshape(5, 51, 3)
    out = []
    for name in BUCKETED_METACATAGORIES:            # the reference's iteration order
        r = CLASS_NAMES.index(name)
        for b in range(51):
            c = m[r, b, 0]
            if c == 0 and not (name == "BACKGROUND" and b == 0):
                continue
            if name == "BACKGROUND" and b > 0:
                continue
            thr = (0.0, 0.04) if name == "BACKGROUND" else (SPEED_SPLITS[b], SPEED_SPLITS[b + 1])
            out.append((name, m[r, b, 1] / c if c else float("nan"), m[r, b, 2] / c if c else float("nan"), thr, int(c)))
    return out


def evaluate_ssf(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids) -> List[tuple]:
    """eval_metric.py:81-106 -> [(motion, avg_epe, avg_distance, (lo, hi), count)] for the non-empty cells."""
    a = accumulate_frame(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids).cpu().numpy()
    m = a[_SSF:].reshape(5, 2, 3)
    out = []
    for r in range(5):
        for motion, k in (("Dynamic", 1), ("Static", 0)):
            c = m[r, k, 0]
            if c:
                out.append((motion, m[r, k, 1] / c, m[r, k, 2] / c, (DISTANCE_SPLIT[r], DISTANCE_SPLIT[r + 1]), int(c)))
    return out


class OfficialMetrics:
    """eval_metric.OfficialMetrics (eval_metric.py:235-345) fed from the device: ``step_frame`` launches one kernel and
    keeps the frame's record on the GPU; nothing is read back until ``normalize()``."""

    def __init__(self, device="cuda", capacity: int = 256):
        self.device = torch.device(device)
        self.rows = torch.zeros((capacity, ACC), dtype=torch.float64, device=self.device)
        self.n_frames = 0
        self.num_occupied_voxels: List[int] = []
        self.epe_3way: Dict = {}
        self.bucketed: Dict = {}
        self.epe_ssf: Dict = {}
        self.norm_flag = False

    def step_frame(self, est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids, num_occupied_voxels=-1):
        if self.n_frames == self.rows.shape[0]:
            self.rows = torch.cat([self.rows, torch.zeros_like(self.rows)], 0)
        accumulate_frame(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids, self.rows[self.n_frames])
        self.n_frames += 1
        self.num_occupied_voxels.append(num_occupied_voxels)

    def step_batch(self, batch, res_dict):
        """``ModelWrapper.train_validation_step_`` (OSF/src/trainer.py:154-171) for a whole batch: final flow = pose flow +
        estimate on the valid points of every sample."""
        for b, gt_flow in enumerate(batch["flow"]):
            idx = res_dict["pc0_valid_point_idxes"][b]
            pose_flow = res_dict["pose_flow"][b][idx]
            self.step_frame(pose_flow + res_dict["flow"][b], pose_flow, batch["pc0"][b][idx], gt_flow[idx],
                            batch["flow_is_valid"][b][idx], batch["flow_category_indices"][b][idx])

    def normalize(self):
        a = self.rows[:self.n_frames].cpu().numpy()          # the one device -> host copy
        # EPE three-way: per-frame values, mean over frames (eval_metric.py:303-306)
        per = [_three_way(r) for r in a]
        self.epe_3way = {k: float(np.mean([p[k] for p in per])) if per else float("nan") for k in ("EPE_FD", "EPE_BS", "EPE_FS", "IoU")}
        self.epe_3way["Three-way"] = float(np.mean([self.epe_3way["EPE_FD"], self.epe_3way["EPE_BS"], self.epe_3way["EPE_FS"]]))
        # bucketed: count-weighted running averages == pooled sums / pooled counts (BucketResultMatrix.accumulate_value)
        m = a[:, _V2:_SSF].sum(0).reshape(5, 51, 3)
        with np.errstate(divide="ignore", invalid="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            epe = np.where(m[:, :, 0] > 0, m[:, :, 1] / m[:, :, 0], np.nan)
            rng = np.where(m[:, :, 0] > 0, m[:, :, 2] / m[:, :, 0], np.nan)
            err = epe.copy()
            err[:, 1:] = err[:, 1:] / rng[:, 1:]                                   # get_normalized_error_matrix
            dyn = np.nanmean(err[:, 1:], axis=1)                                   # get_overall_class_errors
            self.bucketed = {c: {"Static": float(err[i, 0]), "Dynamic": float(dyn[i])} for i, c in enumerate(CLASS_NAMES)}
            self.bucketed["Mean"] = {"Static": float(np.nanmean(err[:, 0])), "Dynamic": float(np.nanmean(dyn))}
            # range-wise (distanceMatrix, eval_metric.py:321-345)
            s = a[:, _SSF:].sum(0).reshape(5, 2, 3)
            self.epe_ssf = {}
            for r in range(5):
                lo, hi = DISTANCE_SPLIT[r], DISTANCE_SPLIT[r + 1]
                key = f"{int(lo)}-{int(hi)}" if hi != np.inf else f"{int(lo)}-inf"
                ent = {}
                for motion, k in (("Static", 0), ("Dynamic", 1)):
                    c = s[r, k, 0]
                    ent[motion] = float(s[r, k, 1] / c) if c else float("nan")
                    ent["#" + motion] = int(c)
                    ent["dist_" + motion] = float(s[r, k, 2] / c) if c else float("nan")
                self.epe_ssf[key] = ent
            self.epe_ssf["Mean"] = {mo: float(np.nanmean([self.epe_ssf[k][mo] for k in list(self.epe_ssf)])) for mo in ("Static", "Dynamic")}
        self.norm_flag = True
        return self
