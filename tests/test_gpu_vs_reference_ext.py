"""The CUDA path against the REFERENCE's own CUDA kernels (oracle/_ref/mmcv_ref_ext.so, compiled from /root/reference
by oracle/build_ref.py) on the same inputs, live on the GPU: integer results bit-exact."""
import numpy as np
import pytest
import torch

from deflow_b200 import ops, synth
from oracle import build_ref
import importlib.util, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_ext_golden_gpu as gen

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def ref():
    m = build_ref.load_ref()
    if m is None:
        pytest.skip("oracle/_ref/mmcv_ref_ext.so not built (python oracle/build_ref.py in the build container)")
    return m


@pytest.mark.parametrize("case", range(len(gen.VOX_CASES)))
def test_voxelize_equals_reference_kernel(ref, case):
    vs, rg = gen.VOX_CASES[case]
    for n, seed in ((20000, 100 + case), (200000, 7), (1, 3)):
        pts = torch.from_numpy(gen.vox_points(max(n, 3), seed, rg)[:n]).to(DEV)
        a = torch.zeros((n, 3), dtype=torch.int32, device=DEV)
        b = torch.zeros((n, 3), dtype=torch.int32, device=DEV)
        ref.dynamic_voxelize_forward(pts, torch.tensor(vs, dtype=torch.float32), torch.tensor(rg, dtype=torch.float32), a, 3)
        ops.dynamic_voxelize_forward(pts, torch.tensor(vs, dtype=torch.float32), torch.tensor(rg, dtype=torch.float32), b, 3)
        assert torch.equal(a, b)


@pytest.mark.parametrize("case", range(len(gen.SCATTER_CASES)))
@pytest.mark.parametrize("red", ["sum", "mean", "max"])
def test_scatter_equals_reference_kernel(ref, case, red):
    n, c, span, seed = gen.SCATTER_CASES[case]
    coors, feats, gseed = gen.scatter_case(n, c, span, seed)
    tf, tc = torch.from_numpy(feats).to(DEV), torch.from_numpy(coors).to(DEV)
    rf, rc, rmap, rcnt = ref.dynamic_point_to_voxel_forward(tf, tc, red)
    vf, vc, cmap, cnt = ops.dynamic_point_to_voxel_forward(tf, tc, red)
    assert torch.equal(vc, rc) and torch.equal(cmap, rmap) and torch.equal(cnt, rcnt)   # bit-exact integers
    # features are multiples of 1/8: sums are exact in fp32, so sum / max must be bit-identical too
    if red in ("sum", "max"):
        assert torch.equal(vf, rf)
    else:
        np.testing.assert_allclose(vf.cpu().numpy(), rf.cpu().numpy(), rtol=1e-6, atol=1e-7)
    gv = torch.from_numpy(np.random.default_rng(gseed).normal(size=tuple(rf.shape)).astype(np.float32)).to(DEV)
    g0, g1 = torch.zeros((n, c), device=DEV), torch.full((n, c), 3.0, device=DEV)
    ref.dynamic_point_to_voxel_backward(g0, gv, tf, rf, rmap, rcnt, red)
    ops.dynamic_point_to_voxel_backward(g1, gv, tf, vf, cmap, cnt, red)
    np.testing.assert_allclose(g1.cpu().numpy(), g0.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_batched_pillar_index_equals_reference_kernels_on_synthetic_av2(ref):
    """DeFlow's own call sequence (DynamicVoxelizer + DynamicScatter per frame) with the reference kernels vs the
    batched pillar index, on AV2-shaped frames at BASELINE configs[1] size."""
    vs, rg = [0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]
    batch = synth.make_batch(2, 80000, seed=5)
    pts = torch.cat([batch["pc0"][:, :80000], batch["pc1"][:, :80000]], 0).contiguous().to(DEV)
    idx = ops.pillar_index(pts, vs, rg)
    for f in range(pts.shape[0]):
        p = pts[f]
        keep = ~torch.isnan(p).any(1)
        pv = p[keep].contiguous()
        coors = torch.zeros((pv.shape[0], 3), dtype=torch.int32, device=DEV)
        ref.dynamic_voxelize_forward(pv, torch.tensor(vs, dtype=torch.float32), torch.tensor(rg, dtype=torch.float32), coors, 3)
        ok = (coors != -1).all(1)
        coors, pv = coors[ok].contiguous(), pv[ok].contiguous()
        vf, vc, cmap, cnt = ref.dynamic_point_to_voxel_forward(pv, coors, "mean")
        a, b, q0, q1 = idx.pt_off(f), idx.pt_off(f + 1), idx.pil_off(f), idx.pil_off(f + 1)
        assert torch.equal(idx.pt_coor[a:b], coors) and torch.equal(idx.pt_xyz[a:b], pv)
        assert torch.equal(idx.pil_coor[q0:q1], vc) and torch.equal(idx.pil_cnt[q0:q1], cnt)
        assert torch.equal(idx.pt_pillar[a:b] - q0, cmap)
