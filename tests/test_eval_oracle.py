"""The evaluation-metric oracle (oracle/eval_oracle.py) against the reference's OWN functions
(OSF/src/utils/eval_metric.py, av2_eval.py imported unmodified with stub av2 modules).  CPU only."""
import numpy as np
import pytest

from oracle import eval_oracle as eo
from oracle import ref_modules


@pytest.fixture(scope="module")
def em():
    if ref_modules.root() is None:
        pytest.skip("reference sources not available")
    return ref_modules.load_eval_metric()


@pytest.mark.parametrize("n,seed", [(20000, 1), (3000, 2), (50, 3)])
def test_frame_metrics_equal_reference(em, n, seed):
    f = eo.make_frame(n, seed)
    r1 = em.evaluate_leaderboard(*f)
    o1 = eo.evaluate_leaderboard(*f)
    for k in ("EPE_BS", "EPE_FD", "EPE_FS", "IoU"):
        assert abs(r1[k] - o1[k]) <= 1e-12 * max(1.0, abs(r1[k])), k
    r2 = em.evaluate_leaderboard_v2(*f)
    o2 = eo.evaluate_leaderboard_v2(*f)
    assert len(r2) == len(o2)
    for a, b in zip(r2, o2):
        assert (a.name, tuple(a.thresholds_range), int(a.count)) == (b[0], tuple(b[3]), b[4])
        if b[4]:
            assert abs(a.avg_epe - b[1]) <= 1e-12 and abs(a.avg_range - b[2]) <= 1e-12
    r3 = em.evaluate_ssf(*f)
    o3 = eo.evaluate_ssf(*f)
    assert len(r3) == len(o3)
    for a, b in zip(r3, o3):
        assert (a.name, tuple(a.thresholds_range), int(a.count)) == (b[0], tuple(b[3]), b[4])
        assert abs(a.avg_epe - b[1]) <= 1e-12 and abs(a.avg_range - b[2]) <= 1e-12


def test_accumulated_and_normalised_metrics_equal_reference(em):
    ref, orc = em.OfficialMetrics(), eo.OfficialMetrics()
    for seed in range(4):
        f = eo.make_frame(8000, 10 + seed)
        ref.step(em.evaluate_leaderboard(*f), em.evaluate_leaderboard_v2(*f), em.evaluate_ssf(*f))
        orc.step(eo.evaluate_leaderboard(*f), eo.evaluate_leaderboard_v2(*f), eo.evaluate_ssf(*f))
    ref.normalize()
    out = orc.normalize()
    for k, v in out["epe_3way"].items():
        assert abs(ref.epe_3way[k] - v) <= 1e-12, k
    for c, d in out["bucketed"].items():
        for kk in ("Static", "Dynamic"):
            a, b = ref.bucketed[c][kk], d[kk]
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-12, (c, kk)   # BACKGROUND has no dynamic buckets
    for i, motion in enumerate(["Static", "Dynamic"]):
        epe, dis, cnt = ref.distanceMatrix.get_class_entries(motion)
        np.testing.assert_allclose(out["ssf"]["epe"][i], epe, rtol=1e-12, equal_nan=True)
        np.testing.assert_array_equal(out["ssf"]["count"][i], cnt)


def test_category_tables():
    assert len(eo.ANNOTATION_CATEGORIES) == 30 and eo.CATEGORY_TO_INDEX["REGULAR_VEHICLE"] == 19
    assert eo.CATEGORY_TO_INDEX["PEDESTRIAN"] == 17 and eo.CATEGORY_TO_INDEX["WHEELED_RIDER"] == 30
    assert abs(eo.SPEED_SPLITS[1] - 0.04) == 0 and len(eo.SPEED_SPLITS) == 52
