"""oracle/mmcv_ext_oracle.py against outputs of the REFERENCE's own CUDA kernels recorded on a B200
(tests/golden/ext_gpu_ref.npz, made by tests/golden/make_ext_golden_gpu.py).  CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_ext_golden_gpu as gen  # noqa: E402
from oracle import mmcv_ext_oracle as ext  # noqa: E402

FIX = os.path.join(HERE, "golden", "ext_gpu_ref.npz")


@pytest.fixture(scope="module")
def fx():
    if not os.path.exists(FIX):
        pytest.skip("ext_gpu_ref.npz not recorded yet")
    z = np.load(FIX)
    return {k: z[k] for k in z.files}


def test_voxelize_oracle_matches_reference_cuda(fx):
    for i, (vs, rg) in enumerate(gen.VOX_CASES):
        pts = gen.vox_points(20000, 100 + i, rg)
        assert np.array_equal(ext.dynamic_voxelize_forward(pts, vs, rg), fx[f"vox{i}_coors"])


def test_scatter_oracle_matches_reference_cuda(fx):
    for i, (n, c, span, seed) in enumerate(gen.SCATTER_CASES):
        if i not in gen.RECORDED:
            continue
        coors, feats, gseed = gen.scatter_case(n, c, span, seed)
        for red in ("sum", "mean", "max"):
            vf, vc, cmap, cnt = ext.dynamic_point_to_voxel_forward(feats, coors, red)
            assert np.array_equal(vc, fx[f"sc{i}_coors"]) and np.array_equal(cmap, fx[f"sc{i}_map"])
            assert np.array_equal(cnt, fx[f"sc{i}_count"])
            np.testing.assert_allclose(vf, fx[f"sc{i}_{red}_feats"], rtol=1e-6, atol=1e-7)
            gv = np.random.default_rng(gseed).normal(size=vf.shape).astype(np.float32)
            g = ext.dynamic_point_to_voxel_backward(gv, feats, fx[f"sc{i}_{red}_feats"], cmap, cnt, red)
            np.testing.assert_allclose(g, fx[f"sc{i}_{red}_grad"], rtol=1e-6, atol=1e-7)
