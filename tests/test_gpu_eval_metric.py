"""On-device evaluation metrics (deflow_b200/eval_metric.py, csrc/eval_metric.cu) against the oracle restatement of
OSF/src/utils/eval_metric.py + av2_eval.py (itself pinned against the reference's functions, tests/test_eval_oracle.py):
counts exact, float64 means to 1e-9."""
import numpy as np
import pytest
import torch

import deflow_b200 as d
from deflow_b200 import eval_metric as em
from oracle import eval_oracle as eo

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dev(f):
    return [t.to(DEV) for t in f]


@pytest.mark.parametrize("n,seed", [(20000, 1), (3000, 2), (50, 3), (200000, 4), (1, 5)])
def test_frame_metrics_equal_oracle(n, seed):
    f = eo.make_frame(n, seed)
    g = _dev(f)
    o1, o2, o3 = eo.evaluate_leaderboard(*f), eo.evaluate_leaderboard_v2(*f), eo.evaluate_ssf(*f)
    acc = em.accumulate_frame(*g).cpu().numpy()
    assert [int(c) for c in acc[0:8]] == o1["_count"]                   # subset counts: exact
    assert tuple(int(c) for c in acc[16:19]) == o1["_tp_fp_fn"]
    r1 = em.evaluate_leaderboard(*g)
    for k in ("EPE_BS", "EPE_FD", "EPE_FS", "IoU"):
        assert abs(r1[k] - o1[k]) <= 1e-9 * max(1.0, abs(o1[k])), k
    r2 = em.evaluate_leaderboard_v2(*g)
    if o2[0][4] == 0:                    # a frame without valid background points: the reference's NaN entry
        assert r2[0][4] == 0
        r2, o2 = r2[1:], o2[1:]
    assert [(a[0], tuple(a[3]), a[4]) for a in r2] == [(b[0], tuple(b[3]), b[4]) for b in o2]
    for a, b in zip(r2, o2):
        assert abs(a[1] - b[1]) <= 1e-9 and abs(a[2] - b[2]) <= 1e-9
    r3 = em.evaluate_ssf(*g)
    assert [(a[0], tuple(a[3]), a[4]) for a in r3] == [(b[0], tuple(b[3]), b[4]) for b in o3]
    for a, b in zip(r3, o3):
        assert abs(a[1] - b[1]) <= 1e-9 and abs(a[2] - b[2]) <= 1e-8 * max(1.0, b[2])


def test_official_metrics_accumulation_equals_oracle():
    orc = eo.OfficialMetrics()
    dev = em.OfficialMetrics(DEV, capacity=2)          # capacity 2: the record buffer has to grow
    for seed in range(5):
        f = eo.make_frame(8000, 20 + seed)
        orc.step(eo.evaluate_leaderboard(*f), eo.evaluate_leaderboard_v2(*f), eo.evaluate_ssf(*f))
        dev.step_frame(*_dev(f))
    want = orc.normalize()
    dev.normalize()
    for k, v in want["epe_3way"].items():
        assert abs(dev.epe_3way[k] - v) <= 1e-9, k
    for c, dd in want["bucketed"].items():
        for kk in ("Static", "Dynamic"):
            a, b = dev.bucketed[c][kk], dd[kk]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9, (c, kk)
    keys = ["0-35", "35-50", "50-75", "75-100", "100-inf"]
    for r, key in enumerate(keys):
        for i, motion in enumerate(["Static", "Dynamic"]):
            assert dev.epe_ssf[key]["#" + motion] == int(want["ssf"]["count"][i, r])
            a, b = dev.epe_ssf[key][motion], want["ssf"]["epe"][i, r]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-9


def test_validation_step_from_model_outputs():
    """train_validation_step_ (OSF/src/trainer.py:154-171): metrics straight from DeFlow's result dict, one launch per sample,
    no per-point host copies; equals the oracle fed with the same tensors."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import load_fixture, batch_to
    from oracle import deflow_oracle as orc
    fx, batch, cfg = load_fixture("deflow_small_gru_eval")
    m = d.DeFlow(cfg["voxel_size"], cfg["range"], cfg["grid"], cfg["decoder"], 4, precision="fp32")
    m.load_state_dict(orc.random_state(cfg["seed_state"], cfg["decoder"]), strict=True)
    m = m.to(DEV).eval()
    gb = batch_to(batch, DEV)
    gb["flow_is_valid"] = torch.ones(gb["flow"].shape[:2], dtype=torch.bool, device=DEV)
    with torch.no_grad():
        res = m(gb)
    met = em.OfficialMetrics(DEV)
    met.step_batch(gb, res)
    met.normalize()
    ref = eo.OfficialMetrics()
    for b in range(len(batch["pose0"])):
        i = res["pc0_valid_point_idxes"][b].cpu()
        pf = res["pose_flow"][b].cpu()[i]
        args = (pf + res["flow"][b].cpu(), pf, batch["pc0"][b][i], batch["flow"][b][i], torch.ones(len(i), dtype=torch.bool),
                batch["flow_category_indices"][b][i])
        ref.step(eo.evaluate_leaderboard(*args), eo.evaluate_leaderboard_v2(*args), eo.evaluate_ssf(*args))
    want = ref.normalize()
    for k, v in want["epe_3way"].items():
        assert abs(met.epe_3way[k] - v) <= 1e-6 * max(1.0, abs(v)), k
