"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures made by
the reference's own modules.  Integer / index results must be bit-exact; floating point within the
tolerance written next to each assert (north_star: per-point flow within 1e-3 abs fp32)."""
import numpy as np
import pytest
import torch

import deflow_b200 as d
from deflow_b200 import ops, synth
from oracle import deflow_oracle as orc
from oracle import mmcv_ext_oracle as ext
from helpers import load_fixture, batch_to

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS, RG = [0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]
SMALL_RG = [-6.4, -6.4, -3, 6.4, 6.4, 3]


def _boundary_points(n, rng, rg):
    """Random points plus many exactly on / next to voxel boundaries and the range limits."""
    p = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32) * np.array([rg[3], rg[4], 4.0], np.float32)
    k = n // 4
    g = np.round(p[:k, :2] / 0.2) * 0.2
    p[:k, :2] = g.astype(np.float32)
    p[k:2 * k, 0] = np.nextafter(p[k:2 * k, 0], np.float32(np.inf))
    special = [[rg[0], rg[1], rg[2]], [rg[3], rg[4], rg[5]],
               [np.nextafter(np.float32(rg[3]), np.float32(-np.inf)), 0, 0],
               [0, 0, np.nextafter(np.float32(rg[5]), np.float32(-np.inf))]]
    for i, sp in enumerate(special[:n]):
        p[i] = sp
    return p


# ----------------------------------------------------------------------------- mmcv._ext drop-ins
@pytest.mark.parametrize("n,nf", [(1, 3), (777, 3), (20000, 4), (200000, 3)])
def test_dynamic_voxelize_bit_exact(n, nf):
    rng = np.random.default_rng(n)
    pts = np.concatenate([_boundary_points(n, rng, RG), rng.normal(size=(n, nf - 3)).astype(np.float32)], 1) if nf > 3 \
        else _boundary_points(n, rng, RG)
    for vs, rg in ((VS, RG), ([0.1, 0.1, 6], RG), ([0.3, 0.25, 0.2], [-10, -7, -3, 11.1, 8.3, 3.4])):
        ref = ext.dynamic_voxelize_forward(pts, vs, rg)
        coors = torch.zeros((n, 3), dtype=torch.int32, device=DEV)
        ops.dynamic_voxelize_forward(torch.from_numpy(pts).to(DEV), torch.tensor(vs), torch.tensor(rg), coors, 3)
        assert np.array_equal(coors.cpu().numpy(), ref)


def test_dynamic_voxelize_empty_and_module():
    v = d.Voxelization(VS, RG, max_num_points=-1)
    out = v(torch.zeros((0, 3), device=DEV))
    assert out.shape == (0, 3) and out.dtype == torch.int32
    pts = torch.tensor([[0.05, 0.05, 0.0], [100.0, 0, 0], [0, 100.0, 0], [0, 0, 100.0]], device=DEV)
    assert v(pts).cpu().tolist() == [[0, 256, 256], [-1, 0, 0], [-1, -1, 0], [-1, -1, -1]]


@pytest.mark.parametrize("reduce", ["sum", "mean", "max"])
@pytest.mark.parametrize("n,c,span", [(1, 3, 4), (513, 3, 6), (5000, 32, 20), (60000, 32, 300), (3000, 7, 3)])
def test_dynamic_scatter_forward_backward(reduce, n, c, span):
    rng = np.random.default_rng(n * 7 + c)
    coors = np.stack([rng.integers(0, 2, n), rng.integers(0, span, n), rng.integers(0, span, n)], 1).astype(np.int32)
    bad = rng.random(n) < 0.1
    coors[bad, rng.integers(0, 3, int(bad.sum()))] = -1  # rows with any negative component are invalid
    feats = rng.normal(size=(n, c)).astype(np.float32)
    if reduce == "max":
        feats = np.round(feats * 4) / 4  # ties: the smallest point index must take the gradient
    r_feats, r_coors, r_map, r_cnt = ext.dynamic_point_to_voxel_forward(feats, coors, reduce)
    tf, tc = torch.from_numpy(feats).to(DEV), torch.from_numpy(coors).to(DEV)
    vf, vc, cmap, cnt = ops.dynamic_point_to_voxel_forward(tf, tc, reduce)
    assert np.array_equal(vc.cpu().numpy(), r_coors)       # sorted unique coords: exact
    assert np.array_equal(cmap.cpu().numpy(), r_map)       # point2voxel_map: exact
    assert np.array_equal(cnt.cpu().numpy(), r_cnt)        # counts: exact
    # fp32 accumulation in an unspecified order on both sides (the reference adds with atomics, scatter_points_cuda_kernel.cuh:
    # 91-112): the error of a sum grows with the number of addends -- ~170 points per pillar in the (3000, 7, 3) case
    atol = 2e-6 + (2e-7 * float(r_cnt.max()) if reduce == "sum" and r_cnt.size else 0.0)
    np.testing.assert_allclose(vf.cpu().numpy(), r_feats, rtol=1e-5, atol=atol)
    gv = rng.normal(size=r_feats.shape).astype(np.float32)
    r_grad = ext.dynamic_point_to_voxel_backward(gv, feats, vf.cpu().numpy(), r_map, r_cnt, reduce)
    g = torch.full((n, c), 7.0, device=DEV)
    ops.dynamic_point_to_voxel_backward(g, torch.from_numpy(gv).to(DEV), tf, vf, cmap, cnt, reduce)
    np.testing.assert_allclose(g.cpu().numpy(), r_grad, rtol=1e-6, atol=1e-7)


def test_dynamic_scatter_module_autograd_and_edge_cases():
    sc = d.DynamicScatter(VS, RG, average_points=True)
    f = torch.randn(100, 5, device=DEV, requires_grad=True)
    c = torch.randint(0, 4, (100, 3), device=DEV, dtype=torch.int32)
    vf, vc = sc(f, c)
    vf.square().sum().backward()
    _, rc, rmap, rcnt = ext.dynamic_point_to_voxel_forward(f.detach().cpu().numpy(), c.cpu().numpy(), "mean")
    assert np.array_equal(vc.cpu().numpy(), rc)
    ref = ext.dynamic_point_to_voxel_backward(2 * vf.detach().cpu().numpy(), f.detach().cpu().numpy(),
                                              vf.detach().cpu().numpy(), rmap, rcnt, "mean")
    np.testing.assert_allclose(f.grad.cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    # n == 0 returns clones and empty int32 (scatter_points_cuda.cu:15-18)
    out = ops.dynamic_point_to_voxel_forward(torch.zeros((0, 3), device=DEV), torch.zeros((0, 3), dtype=torch.int32, device=DEV), "mean")
    assert out[0].shape == (0, 3) and out[2].numel() == 0 and out[3].dtype == torch.int32
    # all rows invalid -> zero voxels, map all -1
    out = ops.dynamic_point_to_voxel_forward(torch.ones((5, 3), device=DEV), torch.full((5, 3), -1, dtype=torch.int32, device=DEV), "mean")
    assert out[0].shape[0] == 0 and out[2].cpu().tolist() == [-1] * 5
    with pytest.raises(RuntimeError):
        ops.dynamic_point_to_voxel_forward(torch.ones((5, 3), device=DEV), torch.zeros((5, 3), dtype=torch.int32, device=DEV), "median")
    # heavy collisions: every point in one pillar (max 746 points / pillar on the real sweep)
    f = torch.randn(4000, 32, device=DEV)
    vf, vc, m, cnt = ops.dynamic_point_to_voxel_forward(f, torch.full((4000, 3), 3, dtype=torch.int32, device=DEV), "mean")
    assert cnt.cpu().tolist() == [4000]
    np.testing.assert_allclose(vf.cpu().numpy()[0], f.double().mean(0).cpu().numpy(), atol=1e-5)


# ----------------------------------------------------------------------------- batched pillar index
def _oracle_frames(pts_cpu, vs, rg):
    out = []
    for f in range(pts_cpu.shape[0]):
        info = orc.voxelize_frame(pts_cpu[f], vs, rg)
        vc, cmap, cnt = ext.unique_pillars(info["voxel_coords"].numpy())
        out.append((info, vc, cmap, cnt))
    return out


@pytest.mark.parametrize("F,n,vs,rg", [(1, 1, VS, RG), (4, 3000, VS, RG), (2, 20000, VS, RG), (3, 5000, [0.2, 0.2, 6], SMALL_RG),
                                       (2, 30000, [0.1, 0.1, 6], RG)])
def test_pillar_index_bit_exact(F, n, vs, rg):
    rng = np.random.default_rng(F * 1000 + n)
    pts = np.stack([_boundary_points(n, rng, rg) for _ in range(F)], 0)
    if n > 10:
        pts[0, n // 2:] = np.nan            # NaN padding (collate_fn_pad)
        pts[-1, 5] = [np.nan, 0.0, 0.0]     # a row with a single NaN is dropped as a whole
    tp = torch.from_numpy(pts)
    idx = ops.pillar_index(tp.to(DEV), vs, rg)
    ref = _oracle_frames(tp, vs, rg)
    for f, (info, vc, cmap, cnt) in enumerate(ref):
        a, b = idx.pt_off(f), idx.pt_off(f + 1)
        q0, q1 = idx.pil_off(f), idx.pil_off(f + 1)
        assert b - a == info["points"].shape[0] == idx.n_valid(f)
        assert q1 - q0 == vc.shape[0] == idx.n_pillars(f)
        assert torch.equal(idx.pt_idx[a:b].cpu(), info["point_idxes"])
        assert torch.equal(idx.pt_coor[a:b].cpu(), info["voxel_coords"])
        assert torch.equal(idx.pt_xyz[a:b].cpu(), info["points"])
        assert torch.equal(idx.pt_offs[a:b].cpu(), info["point_offsets"])          # same fp32 op order: exact
        assert np.array_equal(idx.pt_pillar[a:b].cpu().numpy() - q0, cmap)        # point2voxel_map
        assert np.array_equal(idx.pil_coor[q0:q1].cpu().numpy(), vc)              # unique_dim order
        assert np.array_equal(idx.pil_cnt[q0:q1].cpu().numpy(), cnt)
        gx = idx.grid[0]
        assert np.array_equal(idx.pil_pix[q0:q1].cpu().numpy(), f * idx.grid[0] * idx.grid[1] + vc[:, 1] * gx + vc[:, 2])
    # CSR: every pillar lists exactly its own points
    n_tot, m_tot = idx.pt_off(F), idx.pil_off(F)
    start = idx.pil_start[:m_tot + 1].cpu().numpy()
    srt = idx.sorted_pt[:n_tot].cpu().numpy()
    pil = idx.pt_pillar[:n_tot].cpu().numpy()
    assert start[0] == 0 and start[-1] == n_tot
    assert np.array_equal(np.diff(start), idx.pil_cnt[:m_tot].cpu().numpy())
    assert np.array_equal(np.sort(srt), np.arange(n_tot))
    assert np.array_equal(pil[srt], np.repeat(np.arange(m_tot), np.diff(start)))


def test_three_d_voxels_flow4d_configuration():
    """the second consumer of the encoder -- Flow4D voxelises in 3-D (voxel 0.2 m cubes, z in [-3.2, 3.2]: 32
    layers, OSF/src/models/basic/flow4d_module.py:31-115) with the same DynamicVoxelizer / DynamicPillarFeatureNet classes.
    Index integers bit-exact against the oracle; per-voxel features of the per-sample call within fp32 tolerance."""
    vs3, rg3 = [0.2, 0.2, 0.2], [-12.8, -12.8, -3.2, 12.8, 12.8, 3.2]
    rng = np.random.default_rng(77)
    F, n = 2, 6000
    pts = rng.normal(size=(F, n, 3)).astype(np.float32) * np.array([6.0, 6.0, 1.5], np.float32)
    pts[0, n - 50:] = np.nan
    tp = torch.from_numpy(pts)
    idx = ops.pillar_index(tp.to(DEV), vs3, rg3)
    assert idx.grid == (128, 128, 32)
    for f in range(F):
        info = orc.voxelize_frame(tp[f], vs3, rg3)
        vc, cmap, cnt = ext.unique_pillars(info["voxel_coords"].numpy())
        a, b, q0, q1 = idx.pt_off(f), idx.pt_off(f + 1), idx.pil_off(f), idx.pil_off(f + 1)
        assert int(info["voxel_coords"][:, 0].max()) > 0                           # several z layers are occupied
        assert torch.equal(idx.pt_coor[a:b].cpu(), info["voxel_coords"]) and torch.equal(idx.pt_idx[a:b].cpu(), info["point_idxes"])
        assert np.array_equal(idx.pil_coor[q0:q1].cpu().numpy(), vc) and np.array_equal(idx.pil_cnt[q0:q1].cpu().numpy(), cnt)
        assert np.array_equal(idx.pt_pillar[a:b].cpu().numpy() - q0, cmap)
        assert torch.equal(idx.pt_offs[a:b].cpu(), info["point_offsets"])
    # per-sample feature net with 16 output channels, as Flow4D configures it
    torch.manual_seed(2)
    net = d.DynamicPillarFeatureNet(3, vs3, rg3, feat_channels=(16,), mode="avg").to(DEV).train()
    state = {"embedder.feature_net." + k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    info = d.DynamicVoxelizer(vs3, rg3)(tp[1:2].to(DEV))[0]
    vf, vcoors, pf = net(info["points"], info["voxel_coords"])
    o = orc.voxelize_frame(tp[1], vs3, rg3)
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    rvf, rvc, rpf, _, _ = orc.pillar_feature_net(o["points"], o["voxel_coords"], state, voxel_size=vs3, pc_range=rg3,
                                                 training=True, buffers=buffers)
    assert torch.equal(vcoors.cpu(), rvc)
    np.testing.assert_allclose(vf.detach().cpu().numpy(), rvf.numpy(), rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(pf.detach().cpu().numpy(), rpf.numpy(), rtol=1e-4, atol=2e-5)


def test_pillar_index_empty_frames():
    pts = torch.full((2, 16, 3), float("nan"))
    pts[1, :4] = torch.tensor([[0.0, 0, 0], [0.01, 0.01, 0], [200.0, 0, 0], [5.0, 5.0, 0.5]])
    idx = ops.pillar_index(pts.to(DEV), VS, RG)
    assert idx.host_counts()[:4] == [0, 3, 0, 2]
    assert idx.pt_idx[:3].cpu().tolist() == [0, 1, 3]


# ----------------------------------------------------------------------------- fused pillar feature net
def _pfn_case(F, n, seed, rg, grid):
    batch = synth.make_batch(F, n, seed=seed)
    pts = batch["pc0"]
    if rg is SMALL_RG:
        pts = pts.clone()
        pts[..., :2] /= 8.0
        pts = pts.half().float()
    state = orc.random_state(seed, "gru")
    return pts, state


@pytest.mark.parametrize("training", [True, False])
def test_pillar_feature_net_forward_backward(training):
    F, grid = 3, (64, 64)
    pts, state = _pfn_case(F, 2500, 17, SMALL_RG, grid)
    # oracle: per-frame loop with autograd
    p = "embedder.feature_net.pfn_layers.0."
    w = state[p + "0.weight"].clone().requires_grad_(True)
    g = state[p + "1.weight"].clone().requires_grad_(True)
    b = state[p + "1.bias"].clone().requires_grad_(True)
    st = dict(state)
    st[p + "0.weight"], st[p + "1.weight"], st[p + "1.bias"] = w, g, b
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    imgs, metas = [], []
    for f in range(F):
        info = orc.voxelize_frame(pts[f], VS, SMALL_RG)
        vf, vc, _, cmap, cnt = orc.pillar_feature_net(info["points"], info["voxel_coords"], st, voxel_size=VS,
                                                      pc_range=SMALL_RG, training=training,
                                                      buffers=buffers if training else None)
        imgs.append(orc.pillars_to_image(vf, vc, *grid))
        metas.append((vf, vc))
    ref_img = torch.cat(imgs, 0)
    gimg = torch.randn(ref_img.shape, generator=torch.Generator().manual_seed(1))
    (ref_img * gimg).sum().backward()

    net = d.DynamicPillarFeatureNet(3, VS, SMALL_RG, feat_channels=(32,), mode="avg").to(DEV)
    net.load_state_dict({k[len("embedder.feature_net."):]: v for k, v in state.items() if k.startswith("embedder.feature_net.")})
    net.train(training)
    idx = ops.pillar_index(pts.to(DEV), VS, SMALL_RG)
    image, pil_feats, pil_mean = net.forward_fused(idx)
    assert image.shape == (F, 64, 64, 32)
    got = image.permute(0, 3, 1, 2).cpu()
    np.testing.assert_allclose(got.detach().numpy(), ref_img.detach().numpy(), rtol=1e-4, atol=2e-5)
    off = 0
    for f in range(F):
        m = metas[f][0].shape[0]
        np.testing.assert_allclose(pil_feats[off:off + m].cpu().numpy(), metas[f][0].detach().numpy(), rtol=1e-4, atol=2e-5)
        off += m
    (image * gimg.permute(0, 2, 3, 1).to(DEV)).sum().backward()
    lin, bn = net.pfn_layers[0][0], net.pfn_layers[0][1]
    for got_g, ref_g, name in ((lin.weight.grad, w.grad, "weight"), (bn.weight.grad, g.grad, "gamma"), (bn.bias.grad, b.grad, "beta")):
        tol = 2e-4 * max(1.0, float(ref_g.abs().max()))
        assert float((got_g.cpu() - ref_g).abs().max()) <= tol, name
    if training:  # 2B sequential running-stat updates in frame order
        np.testing.assert_allclose(bn.running_mean.cpu().numpy(), buffers[p + "1.running_mean"].numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(bn.running_var.cpu().numpy(), buffers[p + "1.running_var"].numpy(), rtol=1e-5, atol=1e-6)
        assert int(bn.num_batches_tracked) == F


def test_pillar_feature_net_per_sample_api_matches_fused():
    pts, state = _pfn_case(1, 3000, 5, SMALL_RG, (64, 64))
    net = d.DynamicPillarFeatureNet(3, VS, SMALL_RG, feat_channels=(32,), mode="avg").to(DEV).eval()
    net.load_state_dict({k[len("embedder.feature_net."):]: v for k, v in state.items() if k.startswith("embedder.feature_net.")})
    vox = d.DynamicVoxelizer(VS, SMALL_RG)
    info = vox(pts.to(DEV))[0]
    vf, vc, pf = net(info["points"], info["voxel_coords"])
    idx = ops.pillar_index(pts.to(DEV), VS, SMALL_RG)
    _, pil_feats, _ = net.forward_fused(idx)
    m = idx.n_pillars(0)
    assert vf.shape[0] == m and torch.equal(vc, idx.pil_coor[:m])
    np.testing.assert_allclose(vf.detach().cpu().numpy(), pil_feats[:m].cpu().numpy(), rtol=1e-4, atol=2e-5)
    canvas = d.PointPillarsScatter(32, (64, 64))(vf, vc)
    assert canvas.shape == (1, 32, 64, 64) and int((canvas.abs().sum(1) > 0).sum()) <= m


# ----------------------------------------------------------------------------- decoder gather
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decoder_gather_forward_backward(dtype):
    B, H, W = 2, 64, 64
    pts, _ = _pfn_case(2 * B, 2000, 23, SMALL_RG, (H, W))
    idx = ops.pillar_index(pts.to(DEV), VS, SMALL_RG)
    n0 = idx.pt_off(B)
    gen = torch.Generator(device=DEV).manual_seed(3)
    img = torch.randn((2 * B, H, W, 32), device=DEV, generator=gen).to(dtype).requires_grad_(True)
    unet = torch.randn((B, H, W, 64), device=DEV, generator=gen).to(dtype).requires_grad_(True)
    h0 = ops.decoder_gather(img, unet, idx, B, n0, torch.float32)
    coor = idx.pt_coor[:n0].long()
    bidx = torch.cat([torch.full((idx.n_valid(b),), b, device=DEV) for b in range(B)])
    ref = torch.cat([img[bidx, coor[:, 1], coor[:, 2]], img[bidx + B, coor[:, 1], coor[:, 2]],
                     unet[bidx, coor[:, 1], coor[:, 2]]], 1).float()
    assert torch.equal(h0, ref.detach())  # pure data movement: exact
    gh = torch.randn(h0.shape, device=DEV, generator=gen)
    gi, gu = torch.autograd.grad(h0, (img, unet), gh)
    ri, ru = torch.autograd.grad(ref, (img, unet), gh)
    tol = 1e-5 if dtype == torch.float32 else 0.15  # bf16 rounding of a sum of up to hundreds of rows
    assert float((gi.float() - ri.float()).abs().max()) <= tol * max(1.0, float(ri.float().abs().max()))
    assert float((gu.float() - ru.float()).abs().max()) <= tol * max(1.0, float(ru.float().abs().max()))


def test_decoder_gather_backward_heavy_pillars_shared_by_a_block():
    """Pillar populations are heavy-tailed: pillars with more than 96 points are summed by the eight warps of a block
    together (a list of up to 64 per block, the rest falls back to one warp each).  160 k single-point pillars + a run of
    160 adjacent pillars with 100 points each (more heavy pillars than one block's list of 64 holds) + one pillar with 3000."""
    B, H, W = 1, 512, 512
    gen = torch.Generator().manual_seed(5)
    cells = torch.randperm(H * W, generator=gen)[:160000]
    cy, cx = cells // W, cells % W
    keep = ~((cy == 300) & (cx >= 100) & (cx < 260)) & ~((cy == 17) & (cx == 33))
    cy, cx = cy[keep], cx[keep]
    run = torch.arange(100, 260).repeat_interleave(100)
    ys = torch.cat([cy, torch.full((run.numel(),), 300), torch.full((3000,), 17)])
    xs = torch.cat([cx, run, torch.full((3000,), 33)])
    n = ys.numel()
    perm = torch.randperm(n, generator=gen)
    ys, xs = ys[perm], xs[perm]
    jit = torch.rand(n, 2, generator=gen) * 0.1 + 0.05
    pc0 = torch.stack([RG[0] + (xs + jit[:, 0] * 5) * VS[0], RG[1] + (ys + jit[:, 1] * 5) * VS[1], torch.zeros(n)], 1)
    pc0 = pc0.half().float()            # coordinates stay inside their cell: offsets 0.25 .. 0.75 of a 0.2 m cell
    pts = torch.stack([pc0, pc0])       # frames: pc0, pc1
    idx = ops.pillar_index(pts.to(DEV), VS, RG)
    n0 = idx.pt_off(B)
    assert n0 == n
    coor = idx.pt_coor[:n0].long()
    cnt = torch.bincount(coor[:, 1] * W + coor[:, 2])
    assert int(cnt.max()) == 3000 and int((cnt == 100).sum()) == 160
    gdev = torch.Generator(device=DEV).manual_seed(3)
    img = torch.randn((2 * B, H, W, 32), device=DEV, generator=gdev).requires_grad_(True)
    unet = torch.randn((B, H, W, 64), device=DEV, generator=gdev).requires_grad_(True)
    h0 = ops.decoder_gather(img, unet, idx, B, n0, torch.float32)
    bidx = torch.zeros(n0, dtype=torch.long, device=DEV)
    ref = torch.cat([img[bidx, coor[:, 1], coor[:, 2]], img[bidx + B, coor[:, 1], coor[:, 2]],
                     unet[bidx, coor[:, 1], coor[:, 2]]], 1)
    assert torch.equal(h0, ref.detach())
    gh = torch.randn(h0.shape, device=DEV, generator=gdev)
    gi, gu = torch.autograd.grad(h0, (img, unet), gh)
    ri, ru = torch.autograd.grad(ref, (img, unet), gh)
    assert float((gi - ri).abs().max()) <= 2e-5 * float(ri.abs().max())
    assert float((gu - ru).abs().max()) <= 2e-5 * float(ru.abs().max())
    # deferred image rows (the training step's path): the image part arrives as compact rows, the UNet part carries its
    # per-channel sums (the bias gradient of the UNet's last convolution) taken from the pillar sums
    sink = {}
    h1 = ops.decoder_gather(img, unet, idx, B, n0, torch.float32, None, sink)
    (gu1,) = torch.autograd.grad(h1, (unet,), gh)
    assert float((gu1 - ru).abs().max()) <= 2e-5 * float(ru.abs().max())
    rows = sink["gather"][0]
    gimg = torch.zeros_like(img)
    ops.gather_img_rows_add(rows, idx, B, H, W, gimg)
    assert float((gimg - ri).abs().max()) <= 2e-5 * float(ri.abs().max())
    cs = getattr(gu1, "_dfb_colsum", None)
    assert cs is not None and cs[1] == gu1._version
    ref_cs = ru.double().sum((0, 1, 2))
    assert float((cs[0].double() - ref_cs).abs().max()) <= 1e-4 * max(1.0, float(ref_cs.abs().max()))


# ----------------------------------------------------------------------------- ego warp + losses
def test_ego_warp_matches_oracle():
    B, N = 3, 1000
    gen = torch.Generator().manual_seed(2)
    pc0 = (torch.randn(B, N, 3, generator=gen) * 20).half().float()
    pc0[1, 900:] = float("nan")
    poses0, poses1 = [], []
    for b in range(B):
        a = 0.01 * (b + 1)
        p0, p1 = torch.eye(4), torch.eye(4)
        p1[:3, :3] = torch.tensor([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1.0]])
        p1[:3, 3] = torch.tensor([1.0 + b, 0.2, 0.01])
        p0[:3, 3] = torch.tensor([0.3, -0.1, 0.0])
        poses0.append(p0)
        poses1.append(p1)
    out = torch.empty((B, N + 7, 3), device=DEV)
    pf, pose01 = ops.ego_warp(pc0.to(DEV), torch.stack(poses0).to(DEV), torch.stack(poses1).to(DEV), None, out)
    for b in range(B):
        t01 = orc.cal_pose0to1(poses0[b], poses1[b])
        warped, ref_pf = orc.ego_compensate(pc0[b], t01)
        np.testing.assert_allclose(pose01[b].cpu().numpy(), t01.numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(out[b, :N].cpu().numpy(), warped.numpy(), rtol=0, atol=1e-5, equal_nan=True)
        np.testing.assert_allclose(pf[b].cpu().numpy(), ref_pf.numpy(), rtol=0, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("name", ["deflowLoss", "ff3dLoss", "zeroflowLoss"])
def test_loss_single_sample_api(name):
    gen = torch.Generator().manual_seed(4)
    n = 5000
    gt = torch.randn(n, 3, generator=gen) * torch.tensor([0.0, 0.03, 0.3])[torch.randint(0, 3, (n, 1), generator=gen)]
    est = (gt + 0.05 * torch.randn(n, 3, generator=gen)).requires_grad_(True)
    est.data[7] = gt[7]  # zero error: gradient must be 0, not NaN
    cls = torch.randint(0, 3, (n,), generator=gen).to(torch.uint8)
    if name == "deflowLoss":
        gt[11] = float("nan")
        ref = orc.deflow_loss(est, gt)
    elif name == "zeroflowLoss":
        gt[11] = float("nan")
        gt[12] = float("inf")
        ref = orc.zeroflow_loss(est, gt)
    else:
        ref = orc.ff3d_loss(est, gt, cls)
    ref.backward()
    e2 = est.detach().to(DEV).requires_grad_(True)
    fn = {"deflowLoss": d.deflowLoss, "ff3dLoss": d.ff3dLoss, "zeroflowLoss": d.zeroflowLoss}[name]
    out = fn({"est_flow": e2, "gt_flow": gt.to(DEV), "gt_classes": cls.to(DEV)})["loss"]
    out.backward()
    assert abs(float(out) - float(ref)) <= 1e-5 * max(1.0, abs(float(ref)))
    np.testing.assert_allclose(e2.grad.cpu().numpy(), est.grad.numpy(), rtol=1e-4, atol=1e-8)


def test_deflow_loss_empty_buckets():
    est = torch.zeros(10, 3, device=DEV)
    gt = torch.zeros(10, 3, device=DEV)
    gt[:, 0] = 0.5  # every point in the fastest bucket; the other two are empty and skipped
    out = d.deflowLoss({"est_flow": est, "gt_flow": gt})["loss"]
    assert abs(float(out) - 0.5) < 1e-6


# ----------------------------------------------------------------------------- the whole model vs golden
def _model_for(cfg, precision="fp32"):
    m = d.DeFlow(cfg["voxel_size"], cfg["range"], cfg["grid"], cfg["decoder"], 4, precision=precision)
    m.load_state_dict(orc.random_state(cfg["seed_state"], cfg["decoder"]), strict=True)
    return m.to(DEV).train(cfg["training"])


@pytest.mark.parametrize("name", ["deflow_small_gru", "deflow_small_linear", "deflow_small_gru_eval"])
def test_model_matches_reference_golden(name):
    fx, batch, cfg = load_fixture(name)
    m = _model_for(cfg)
    gb = batch_to(batch, DEV)
    res = m(gb)
    B = len(batch["pose0"])
    for b in range(B):
        assert np.array_equal(res["pc0_valid_point_idxes"][b].cpu().numpy(), fx[f"pc0_idx_{b}"])
        assert np.array_equal(res["pc1_valid_point_idxes"][b].cpu().numpy(), fx[f"pc1_idx_{b}"])
        np.testing.assert_allclose(res["pose_flow"][b].cpu().numpy(), fx[f"pose_flow_{b}"], atol=1e-5, equal_nan=True)
        err = np.abs(res["flow"][b].detach().cpu().numpy() - fx[f"flow_{b}"]).max()
        assert err <= 1e-3, f"flow error {err}"   # north_star bound
        assert err <= 2e-4, f"flow error {err}"   # what strict fp32 should reach
    assert res["num_occupied_voxels"] == list(fx["num_occupied_voxels"])
    loss = d.training_step_loss(gb, res, cfg["loss"])
    assert abs(float(loss) - float(fx["loss_total"])) <= 2e-4 * max(1.0, abs(float(fx["loss_total"])))
    if cfg["training"]:
        loss.backward()
        named = dict(m.named_parameters())
        for k in fx:
            if k.startswith("grad::"):
                g, ref = named[k[6:]].grad.cpu().numpy(), fx[k]
                assert np.abs(g - ref).max() <= 2e-3 * max(1.0, np.abs(ref).max()), k
        sd = m.state_dict()
        for k in fx:
            if k.startswith("buf::"):
                np.testing.assert_allclose(sd[k[5:]].cpu().numpy(), fx[k], rtol=1e-4, atol=1e-5, err_msg=k)
        # per-sample reference-API loss (OSF/src/trainer.py:120-142 loop) == fused batch loss
        tot = 0.0
        for b in range(B):
            i = res["pc0_valid_point_idxes"][b]
            dct = {"est_flow": res["flow"][b].detach(), "gt_flow": gb["flow"][b][i] - res["pose_flow"][b][i],
                   "gt_classes": gb["flow_category_indices"][b][i]}
            tot += float((d.deflowLoss if cfg["loss"] == "deflowLoss" else d.ff3dLoss)(dct)["loss"])
        assert abs(tot - float(loss)) <= 1e-5 * max(1.0, abs(tot))


@pytest.mark.parametrize("name", ["deflow_cfg1_20k", "deflow_av2_real_20k"])
def test_config1_and_real_sweep(name):
    """BASELINE.json configs[0] (20k-pt pair, 512x512, B=1) and the reference's only real-data fixture."""
    fx, batch, cfg = load_fixture(name)
    m = _model_for(cfg)
    with torch.no_grad():
        res = m(batch_to(batch, DEV))
    assert np.array_equal(res["pc0_valid_point_idxes"][0].cpu().numpy(), fx["pc0_idx_0"])
    assert np.array_equal(res["pc1_valid_point_idxes"][0].cpu().numpy(), fx["pc1_idx_0"])
    err = np.abs(res["flow"][0].cpu().numpy() - fx["flow_0"]).max()
    assert err <= 1e-3, f"flow error {err}"


def test_reference_signature_modules():
    """DynamicEmbedder.forward / ConvGRUDecoder.forward with the reference's own call signatures."""
    fx, batch, cfg = load_fixture("deflow_small_gru_eval")
    m = _model_for(cfg)
    gb = batch_to(batch, DEV)
    with torch.no_grad():
        res = m(gb)
        pc0s = torch.stack([gb["pc0"][b] + 0 for b in range(2)])  # (no ego compensation here; just the API)
        img0, infos0 = m.embedder(pc0s)
        img1, infos1 = m.embedder(gb["pc1"])
        assert img0.shape == (2, 32, 64, 64) and set(infos0[0]) == {"points", "voxel_coords", "point_idxes", "point_offsets"}
        assert infos0[0]["point_idxes"].dtype == torch.int64 and infos0[0]["voxel_coords"].dtype == torch.int32
        feat = m.backbone(img0, img1)
        flows = m.head(torch.cat([img0, img1], 1), feat, infos0)
        assert len(flows) == 2 and flows[0].shape == (infos0[0]["points"].shape[0], 3)
        # same computation through the flat path
        ref_state = orc.random_state(cfg["seed_state"], "gru")
        o = orc.gru_decoder_single(torch.cat([img0, img1], 1)[0].cpu(), feat[0].cpu(), infos0[0]["point_offsets"].cpu(),
                                   infos0[0]["voxel_coords"].cpu(), ref_state)
        assert float((flows[0].cpu() - o).abs().max()) <= 2e-4


def test_benchmark_shape_parity_forward_backward_all_gradients_vs_oracle():
    """The benchmarked shape (BASELINE configs[1]: 80 000 points / frame, 512 x 512 pillars), B = 2, parity mode:
    forward, loss and the gradient of EVERY parameter -- all 29 convolution weights, the GRU gate weights, the PFN --
    against the CPU oracle on the same inputs.  Flow bound: north_star 1e-3 abs."""
    B, n = 2, 80000
    batch = synth.make_batch(B, n, seed=5)
    state = orc.random_state(31, "gru")
    for k, v in state.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    torch.set_num_threads(max(torch.get_num_threads(), __import__("os").cpu_count() or 1))
    ref = orc.deflow_forward(batch, state, VS, RG, (512, 512), "gru", 4, training=True, buffers=buffers)
    ref_loss = orc.training_step_loss(batch, ref, "deflowLoss")
    ref_loss.backward()

    m = d.DeFlow(VS, RG, [512, 512], "gru", 4, precision="fp32")
    m.load_state_dict({k: v.detach() for k, v in orc.random_state(31, "gru").items()}, strict=True)
    m = m.to(DEV).train()
    gb = batch_to(batch, DEV)
    res = m(gb)
    loss = d.training_step_loss(gb, res, "deflowLoss")
    loss.backward()
    worst = 0.0
    for b in range(B):
        assert np.array_equal(res["pc0_valid_point_idxes"][b].cpu().numpy(), ref["pc0_valid_point_idxes"][b].numpy())
        assert np.array_equal(res["pc1_valid_point_idxes"][b].cpu().numpy(), ref["pc1_valid_point_idxes"][b].numpy())
        worst = max(worst, float((res["flow"][b].detach().cpu() - ref["flow"][b].detach()).abs().max()))
    assert worst <= 1e-3, f"flow error {worst}"
    assert abs(float(loss) - float(ref_loss)) <= 2e-4 * max(1.0, abs(float(ref_loss)))
    bad = []
    for k, p in m.named_parameters():
        g, r = p.grad.detach().cpu().double(), state[k].grad.double()
        # norm-relative; the floor covers convolution biases in front of a BatchNorm, whose exact gradient is 0
        if float((g - r).norm()) > 2e-3 * float(r.norm()) + 2e-4 * max(1.0, float(r.abs().max())):
            bad.append((k, float((g - r).norm()), float(r.norm())))
    assert not bad, bad
    sd = m.state_dict()
    for k in buffers:
        if "num_batches" not in k:
            np.testing.assert_allclose(sd[k].cpu().numpy(), buffers[k].numpy(), rtol=2e-4, atol=2e-5, err_msg=k)


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_properties_config2_shape():
    """BASELINE configs[1] shape (80k points / frame, 512x512), a few frames: size-independent invariants."""
    F, n = 4, 80000
    batch = synth.make_batch(F, n, seed=99)
    pts = batch["pc0"].to(DEV)
    idx = ops.pillar_index(pts, VS, RG)
    idx2 = ops.pillar_index(pts, VS, RG)
    hc = idx.host_counts()
    assert hc == idx2.host_counts()                                  # idempotent
    n_tot, m_tot = idx.pt_off(F), idx.pil_off(F)
    assert int(idx.pil_cnt[:m_tot].sum()) == n_tot                   # counts partition the valid points
    assert torch.equal(idx.pt_pillar[:n_tot], idx2.pt_pillar[:n_tot])
    key = idx.pil_pix[:m_tot].long()
    assert bool((key[1:] > key[:-1]).all())                          # pillars strictly sorted (frame, y, x)
    pi = idx.pt_idx[:n_tot]
    for f in range(F):
        a, b = idx.pt_off(f), idx.pt_off(f + 1)
        assert bool((pi[a + 1:b] > pi[a:b - 1]).all())               # stable compaction keeps point order
    # scatter-mean of a constant is the constant; sum of pillar sums == total sum (linearity)
    n1 = idx.pt_off(1)
    coors1 = idx.pt_coor[:n1].contiguous()
    vf, vc, cmap, cnt = ops.dynamic_point_to_voxel_forward(torch.ones((n1, 32), device=DEV), coors1, "mean")
    assert float((vf - 1).abs().max()) == 0.0
    assert vf.shape[0] == idx.n_pillars(0) and torch.equal(cnt, idx.pil_cnt[:idx.n_pillars(0)])
    feats = torch.randn((n1, 32), device=DEV)
    sf = ops.dynamic_point_to_voxel_forward(feats, coors1, "sum")[0]
    np.testing.assert_allclose(sf.double().sum(0).cpu().numpy(), feats.double().sum(0).cpu().numpy(), rtol=1e-6, atol=1e-3)


def test_reused_canvas_matches_fresh_canvas():
    """DynamicEmbedder.reuse_canvas: the persistent pseudo-image with a sparse clear of the previous call's pillar rows
    equals a freshly zero-filled canvas, call after call, for changing inputs (PointPillarsScatter zero canvas,
    OSF/src/models/basic/encoder.py:135-141)."""
    import deflow_b200 as d
    from deflow_b200 import synth
    vs, rg = [0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]
    torch.manual_seed(3)
    emb = d.DynamicEmbedder(vs, [512, 512], rg, 32).to("cuda").eval()
    ref = d.DynamicEmbedder(vs, [512, 512], rg, 32).to("cuda").eval()
    ref.load_state_dict(emb.state_dict())
    emb.reuse_canvas = True
    for seed, n in [(1, 3000), (2, 5000), (3, 10), (4, 4000)]:
        b = synth.make_batch(2, n, seed=seed)
        m = min(b["pc0"].shape[1], b["pc1"].shape[1])
        pts = torch.cat([b["pc0"][:, :m], b["pc1"][:, :m]], 0).contiguous().cuda()
        with torch.no_grad():
            got, idx = emb.embed(pts, torch.bfloat16)
            want, _ = ref.embed(pts, torch.bfloat16)
        assert torch.equal(got, want)
        assert int((got.float().abs().sum(-1) > 0).sum()) <= idx.pil_off(idx.F)


def test_split_frames_backward_is_the_slicing_gradient():
    """deflow._SplitFrames hands every frame half to two consumers (encoder, last skip convolution) on separate autograd
    edges; its backward = the gradient of plain slicing with both consumers added, in one pass (ops.add_cat2)."""
    from deflow_b200.deflow import _SplitFrames
    torch.manual_seed(0)
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(6, 4, 5, 32, device=DEV).to(dt).requires_grad_(True)
        a0, a1, b0, b1 = _SplitFrames.apply(x, 3, None)
        assert torch.equal(a0, x[:3]) and torch.equal(a1, x[3:]) and torch.equal(b0, x[:3]) and torch.equal(b1, x[3:])
        w = torch.randn(3, 4, 5, 32, device=DEV).to(dt)
        ((a0 * w).sum() + (2 * a1 * w).sum() + (3 * b0 * w).sum() + (5 * b1 * w).sum()).backward()
        want = torch.cat([(w.float() + 3 * w.float()), (2 * w.float() + 5 * w.float())], 0)
        tol = 0 if dt == torch.float32 else 2e-2
        assert float((x.grad.float() - want).abs().max()) <= tol * float(want.abs().max())
    # one consumer only, and one half unused
    y = torch.randn(4, 2, 2, 32, device=DEV, requires_grad=True)
    a0, a1, b0, b1 = _SplitFrames.apply(y, 2, None)
    a0.sum().backward()
    assert torch.equal(y.grad, torch.cat([torch.ones(2, 2, 2, 32), torch.zeros(2, 2, 2, 32)]).to(DEV))


def test_fused_image_gradient_path_equals_autograd_accumulation(monkeypatch):
    """The pseudo-image's three gradients (encoder, last skip convolution, decoder gather) met in _SplitFrames.backward
    (deferred gather rows added in place) against the plain autograd accumulation (DFB_IMG_GRAD_FUSE=0): same parameter
    gradients of the pillar feature net, which is all that lies upstream of the image."""
    fx, batch, cfg = load_fixture("deflow_small_gru")
    gb = batch_to(batch, DEV)
    grads = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("DFB_IMG_GRAD_FUSE", fuse)
        m = _model_for(cfg)
        res = m(gb)
        d.training_step_loss(gb, res, cfg["loss"]).backward()
        grads[fuse] = {k: p.grad.clone() for k, p in m.named_parameters()}
    for k in grads["1"]:
        a, b = grads["1"][k], grads["0"][k]
        # (a different summation order of the three image gradients + unordered fp32 atomics: 7e-5 measured on values ~0.5)
        assert float((a - b).abs().max()) <= 5e-4 * max(1.0, float(b.abs().max())), k
