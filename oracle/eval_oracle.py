"""CPU restatement of the reference's evaluation metrics.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows OpenSceneFlow/src/utils/eval_metric.py:28-100 (evaluate_leaderboard / _v2 / evaluate_ssf: masks and thresholds in
fp32 torch, then float64 numpy), src/utils/av2_eval.py:450-553 (compute_metrics: EPE 3-way + dynamic IoU), :839-870
(compute_bucketed_epe), :872-915 (compute_ssf_metrics) and the accumulation / normalisation of OfficialMetrics
(eval_metric.py:235-330).

Third-party dependency absent from /root/reference: ``av2==0.2.1`` (OSF/environment.yaml:35) supplies
``av2.datasets.sensor.constants.AnnotationCategories``; CATEGORY_TO_INDEX = {"NONE": 0, category_i: i + 1} (av2_eval.py:32-35)
depends only on that enum's order, restated below from the published av2-api source (30 categories, alphabetical).
Pinned against the reference's own functions (imported from /root/reference or oracle/_ref/osf with stub ``av2`` modules
that carry exactly this enum) in tests/test_eval_oracle.py."""
import numpy as np
import torch

ANNOTATION_CATEGORIES = [
    "ANIMAL", "ARTICULATED_BUS", "BICYCLE", "BICYCLIST", "BOLLARD", "BOX_TRUCK", "BUS", "CONSTRUCTION_BARREL",
    "CONSTRUCTION_CONE", "DOG", "LARGE_VEHICLE", "MESSAGE_BOARD_TRAILER", "MOBILE_PEDESTRIAN_CROSSING_SIGN", "MOTORCYCLE",
    "MOTORCYCLIST", "OFFICIAL_SIGNALER", "PEDESTRIAN", "RAILED_VEHICLE", "REGULAR_VEHICLE", "SCHOOL_BUS", "SIGN",
    "STOP_SIGN", "STROLLER", "TRAFFIC_LIGHT_TRAILER", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER", "WHEELCHAIR",
    "WHEELED_DEVICE", "WHEELED_RIDER"]
CATEGORY_TO_INDEX = {"NONE": 0, **{k: i + 1 for i, k in enumerate(ANNOTATION_CATEGORIES)}}          # av2_eval.py:32-35
BUCKETED_METACATAGORIES = {                                                                          # av2_eval.py:47-75
    "BACKGROUND": ["NONE"],
    "CAR": ["REGULAR_VEHICLE"],
    "PEDESTRIAN": ["PEDESTRIAN", "STROLLER", "WHEELCHAIR", "OFFICIAL_SIGNALER"],
    "WHEELED_VRU": ["BICYCLE", "BICYCLIST", "MOTORCYCLE", "MOTORCYCLIST", "WHEELED_DEVICE", "WHEELED_RIDER"],
    "OTHER_VEHICLES": ["BOX_TRUCK", "LARGE_VEHICLE", "RAILED_VEHICLE", "TRUCK", "TRUCK_CAB", "VEHICULAR_TRAILER",
                       "ARTICULATED_BUS", "BUS", "SCHOOL_BUS"]}
MATRIX_CLASSES = ["BACKGROUND", "CAR", "OTHER_VEHICLES", "PEDESTRIAN", "WHEELED_VRU"]                # eval_metric.py:262
FOREGROUND = list(range(1, 31))          # av2_eval.py:218-229: the four category enums cover all 30 annotation classes
SPEED_SPLITS = np.concatenate([np.linspace(0, 2.0, 51), [np.inf]])                                  # av2_eval.py:848
DISTANCE_SPLIT = [0, 35, 50, 75, 100, np.inf]                                                        # av2_eval.py:892
CLOSE = 35.0
EPS = 1e-6


def _no_nan(est, rigid, pc0, gt):
    return ~est.isnan().any(1) & ~rigid.isnan().any(1) & ~pc0[:, :3].isnan().any(1) & ~gt.isnan().any(1)


def evaluate_leaderboard(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids):
    """eval_metric.py:28-54 + av2_eval.py:460-553 -> {'EPE_BS','EPE_FD','EPE_FS','IoU'} plus the raw subset table."""
    gt_dyn = torch.linalg.vector_norm(gt_flow - rigid_flow, dim=-1) >= 0.05
    m = _no_nan(est_flow, rigid_flow, pc0, gt_flow)
    est, rigid, pc, gt, gt_dyn, valid, ids = est_flow[m], rigid_flow[m], pc0[m], gt_flow[m], gt_dyn[m], is_valid[m], pts_ids[m]
    est_dyn = torch.linalg.vector_norm(est - rigid, dim=-1) >= 0.05
    close = torch.all(torch.abs(pc[:, :2]) <= CLOSE, dim=1)
    v = valid.numpy().astype(bool)
    pred = est.numpy().astype(float)[v]
    gts = gt.numpy().astype(float)[v]
    pd_, gd, cl, ci = est_dyn.numpy()[v], gt_dyn.numpy()[v], close.numpy()[v], ids.numpy().astype(int)[v]
    count, epe, tp, fp, fn = [], [], 0, 0, 0
    for cats in ([0], FOREGROUND):
        cm = np.isin(ci, cats)
        for mm in (gd, ~gd):
            for dm in (cl, ~cl):
                mask = cm & mm & dm
                n = int(mask.sum())
                count.append(n)
                epe.append(np.linalg.norm(pred[mask] - gts[mask], axis=-1).mean() if n else np.nan)
                tp += int((pd_[mask] & gd[mask]).sum()); fp += int((pd_[mask] & ~gd[mask]).sum()); fn += int((~pd_[mask] & gd[mask]).sum())

    def cepe(idx):
        s, c = 0.0, 0
        for i in idx:
            if count[i]:
                s += epe[i] * count[i]; c += count[i]
        return s / (c + 1e-8) if c else 0.0
    return {"EPE_BS": cepe([2, 3]), "EPE_FD": cepe([4, 5]), "EPE_FS": cepe([6, 7]), "IoU": tp / (tp + fp + fn + EPS),
            "_count": count, "_epe": epe, "_tp_fp_fn": (tp, fp, fn)}


def evaluate_leaderboard_v2(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids):
    """eval_metric.py:57-78 + av2_eval.py:839-870 -> list of (class, avg_epe, avg_speed, (lo, hi), count)."""
    dmask = torch.linalg.vector_norm(pc0[:, :2], dim=-1) <= CLOSE
    m = _no_nan(est_flow, rigid_flow, pc0, gt_flow) & dmask
    rigid = rigid_flow[m]
    est = (est_flow[m] - rigid).numpy().astype(float)
    gt = (gt_flow[m] - rigid).numpy().astype(float)
    valid, ids = is_valid[m].numpy().astype(bool), pts_ids[m].numpy().astype(np.uint8)
    speeds = np.linalg.norm(gt, axis=-1)
    err = np.linalg.norm(est - gt, axis=-1)
    out = []
    for name, cats in BUCKETED_METACATAGORIES.items():
        cm = np.isin(ids, np.array([CATEGORY_TO_INDEX[c] for c in cats]))
        if name == "BACKGROUND":
            mask = cm & valid
            out.append((name, err[mask].mean() if mask.any() else np.nan, speeds[mask].mean() if mask.any() else np.nan,
                        (0.0, 0.04), int(mask.sum())))
            continue
        for lo, hi in zip(SPEED_SPLITS, SPEED_SPLITS[1:]):
            mask = cm & (speeds >= lo) & (speeds < hi) & valid
            if mask.sum() == 0:
                continue
            out.append((name, err[mask].mean(), speeds[mask].mean(), (lo, hi), int(mask.sum())))
    return out


def evaluate_ssf(est_flow, rigid_flow, pc0, gt_flow, is_valid, pts_ids):
    """eval_metric.py:81-106 + av2_eval.py:872-915 -> list of (motion, avg_epe, avg_distance, (lo, hi), count)."""
    dist = torch.linalg.vector_norm(pc0[:, :3], dim=-1)
    m = _no_nan(est_flow, rigid_flow, pc0, gt_flow)
    rigid = rigid_flow[m]
    est = (est_flow[m] - rigid).numpy().astype(float)
    gt = (gt_flow[m] - rigid).numpy().astype(float)
    valid, dist = is_valid[m].numpy().astype(bool), dist[m].numpy().astype(float)
    speeds = np.linalg.norm(gt, axis=-1) * 10
    out = []
    for lo, hi in zip(DISTANCE_SPLIT, DISTANCE_SPLIT[1:]):
        mask = (dist >= lo) & (dist < hi) & valid
        sp, e, g, dd = speeds[mask], est[mask], gt[mask], dist[mask]
        dyn = sp >= 1.4
        for motion, mm in (("Dynamic", dyn), ("Static", ~dyn)):
            if mm.sum() == 0:
                continue
            out.append((motion, np.linalg.norm(e - g, axis=-1)[mm].mean(), dd[mm].mean(), (lo, hi), int(mm.sum())))
    return out


class OfficialMetrics:
    """Accumulation + normalisation of eval_metric.OfficialMetrics (eval_metric.py:235-345), numbers only."""

    def __init__(self):
        self.epe_3way = {k: [] for k in ("EPE_FD", "EPE_BS", "EPE_FS", "IoU")}
        nb = len(SPEED_SPLITS) - 1
        self.b_epe = np.full((5, nb), np.nan); self.b_rng = np.full((5, nb), np.nan); self.b_cnt = np.zeros((5, nb), np.int64)
        nd = len(DISTANCE_SPLIT) - 1
        self.d_epe = np.full((2, nd), np.nan); self.d_rng = np.full((2, nd), np.nan); self.d_cnt = np.zeros((2, nd), np.int64)

    @staticmethod
    def _acc(E, R, C, i, j, epe, rng, cnt):          # BucketResultMatrix.accumulate_value, eval_metric.py:142-172
        if np.isnan(E[i, j]):
            E[i, j], R[i, j], C[i, j] = epe, rng, cnt
            return
        E[i, j] = np.average([E[i, j], epe], weights=[C[i, j], cnt])
        R[i, j] = np.average([R[i, j], rng], weights=[C[i, j], cnt])
        C[i, j] += cnt

    def step(self, v1, v2, ssf):
        for k in self.epe_3way:
            self.epe_3way[k].append(v1[k])
        splits = list(zip(SPEED_SPLITS, SPEED_SPLITS[1:]))
        for name, epe, rng, thr, cnt in v2:
            self._acc(self.b_epe, self.b_rng, self.b_cnt, MATRIX_CLASSES.index(name), splits.index(thr), epe, rng, cnt)
        ds = list(zip(DISTANCE_SPLIT, DISTANCE_SPLIT[1:]))
        for name, epe, rng, thr, cnt in ssf:
            self._acc(self.d_epe, self.d_rng, self.d_cnt, ["Static", "Dynamic"].index(name), ds.index(thr), epe, rng, cnt)

    def normalize(self):
        out = {k: float(np.mean(v)) for k, v in self.epe_3way.items()}
        out["Three-way"] = float(np.mean([out["EPE_FD"], out["EPE_BS"], out["EPE_FS"]]))
        err = self.b_epe.copy()
        err[:, 1:] = err[:, 1:] / self.b_rng[:, 1:]
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", category=RuntimeWarning)
            dyn = np.nanmean(err[:, 1:], axis=1)
            bucketed = {c: {"Static": err[i, 0], "Dynamic": dyn[i]} for i, c in enumerate(MATRIX_CLASSES)}
            bucketed["Mean"] = {"Static": np.nanmean(err[:, 0]), "Dynamic": np.nanmean(dyn)}
        return {"epe_3way": out, "bucketed": bucketed,
                "ssf": {"epe": self.d_epe.copy(), "dist": self.d_rng.copy(), "count": self.d_cnt.copy()}}


def make_frame(n, seed, nan_frac=0.01):
    """A synthetic evaluation frame: points, rigid flow, gt flow with a mix of static / slow / fast objects, classes over the
    whole 0..30 range (+ an out-of-range id), a few invalid and NaN rows."""
    g = torch.Generator().manual_seed(seed)
    pc0 = torch.randn(n, 3, generator=g) * torch.tensor([30.0, 30.0, 1.5])
    rigid = 0.05 * torch.randn(n, 3, generator=g) + torch.tensor([0.8, 0.0, 0.0])
    cls = torch.randint(0, 31, (n,), generator=g, dtype=torch.uint8)
    cls[torch.rand(n, generator=g) < 0.5] = 0
    cls[torch.rand(n, generator=g) < 0.01] = 77
    moving = (cls > 0) | (torch.rand(n, generator=g) < 0.05)          # a few mislabelled / moving background points too
    speed = torch.rand(n, 1, generator=g) ** 3 * 2.5 * moving.unsqueeze(1)
    dirn = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=1)
    gt = rigid + speed * dirn
    est = gt + 0.05 * torch.randn(n, 3, generator=g)
    valid = torch.rand(n, generator=g) < 0.9
    bad = torch.rand(n, generator=g) < nan_frac
    est[bad] = float("nan")
    gt[torch.rand(n, generator=g) < nan_frac / 2] = float("nan")
    return est, rigid, pc0, gt, valid, cls
