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


@pytest.mark.parametrize("n,max_points,max_voxels,seed", [(30000, 10, 20000, 1), (30000, 3, 500, 2), (5000, 35, 20000, 3), (1, 5, 5, 4)])
def test_hard_voxelize_equals_reference_kernel_and_oracle(ref, n, max_points, max_voxels, seed):
    """hard_voxelize_forward (voxelization_cuda.cu:8-148): bit-exact against the reference's own kernels (deterministic path)
    and the numpy oracle -- voxel order of first appearance, per-voxel point order, max_points / max_voxels truncation."""
    from oracle import mmcv_ext_oracle as ext
    vs, rg = [0.4, 0.4, 6], [-20.0, -20.0, -3, 20.0, 20.0, 3]
    rng = np.random.default_rng(seed)
    pts = np.concatenate([rng.normal(size=(n, 3)).astype(np.float32) * np.array([9, 9, 1.5], np.float32),
                          rng.random(size=(n, 1)).astype(np.float32)], 1)        # xyz + intensity
    tp = torch.from_numpy(pts).to(DEV)
    out = {}
    for name, fn in (("ref", ref.hard_voxelize_forward), ("dfb", ops.hard_voxelize_forward)):
        voxels = torch.zeros((max_voxels, max_points, 4), device=DEV)
        coors = torch.zeros((max_voxels, 3), dtype=torch.int32, device=DEV)
        npv = torch.zeros((max_voxels,), dtype=torch.int32, device=DEV)
        vnum = torch.zeros((), dtype=torch.long)
        fn(tp, torch.tensor(vs, dtype=torch.float32), torch.tensor(rg, dtype=torch.float32), voxels, coors, npv, vnum,
           max_points, max_voxels, 3, True)
        m = int(vnum)
        out[name] = (voxels[:m].cpu(), coors[:m].cpu(), npv[:m].cpu())
    ov, oc, on = ext.hard_voxelize_forward(pts, vs, rg, max_points, max_voxels)
    for a, b in zip(out["dfb"], out["ref"]):
        assert torch.equal(a, b)
    assert np.array_equal(out["dfb"][0].numpy(), ov) and np.array_equal(out["dfb"][1].numpy(), oc)
    assert np.array_equal(out["dfb"][2].numpy(), on)
    # the module surface (voxelize.py:115-189)
    import deflow_b200 as d
    v, c, k = d.Voxelization(vs, rg, max_points, max_voxels)(tp)
    assert torch.equal(v.cpu(), out["ref"][0]) and torch.equal(c.cpu(), out["ref"][1]) and torch.equal(k.cpu(), out["ref"][2])
