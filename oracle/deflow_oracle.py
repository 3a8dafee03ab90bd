"""Plain fp32 PyTorch-CPU restatement of the DeFlow hot path (functional style).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  OSF = /root/reference/OpenSceneFlow.
Every function cites the reference lines it follows.  Parameters are read from a
``state`` dict that uses the reference's own ``state_dict`` key names, so a state
dict taken from the reference model (or from ``deflow_b200.DeFlow``) can be fed in
unchanged.  Pinned by ``tests/golden/deflow_*.npz`` which were produced by the
reference's own modules (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F

from . import mmcv_ext_oracle as ext

DEFAULT_VOXEL_SIZE = (0.2, 0.2, 6.0)
DEFAULT_RANGE = (-51.2, -51.2, -3.0, 51.2, 51.2, 3.0)


# ----------------------------------------------------------------------------- pose
def cal_pose0to1(pose0: torch.Tensor, pose1: torch.Tensor) -> torch.Tensor:
    """OSF/src/models/basic/__init__.py:4-15 -- inverse(pose1) @ pose0 in float64, returned fp32."""
    p1 = pose1.to(torch.float64)
    inv = torch.eye(4, dtype=torch.float64)
    inv[:3, :3] = p1[:3, :3].T
    # the reference builds the translation with an fp32 product and a sum over axis 1
    inv[:3, 3] = (pose1[:3, :3].T * -pose1[:3, 3]).sum(dim=1)
    return (inv @ pose0.to(torch.float64)).to(torch.float32)


def ego_compensate(pc0: torch.Tensor, pose_0to1: torch.Tensor):
    """OSF/src/models/deflow.py:72-74."""
    warped = pc0 @ pose_0to1[:3, :3].T + pose_0to1[:3, 3]
    return warped, warped - pc0


# ----------------------------------------------------------------------------- voxelizer
def voxelize_frame(points: torch.Tensor, voxel_size=DEFAULT_VOXEL_SIZE, pc_range=DEFAULT_RANGE) -> dict:
    """One sample of DynamicVoxelizer.forward (OSF/src/models/basic/encoder.py:567-600)."""
    idx = torch.arange(points.shape[0])
    not_nan = ~torch.isnan(points).any(dim=1)
    pts = points[not_nan]
    idx = idx[not_nan]
    coors = torch.from_numpy(ext.dynamic_voxelize_forward(pts.detach().numpy(), voxel_size, pc_range))
    keep = (coors != -1).all(dim=1)
    pts, coors, idx = pts[keep], coors[keep], idx[keep]
    # _get_point_offsets, encoder.py:506-523 (fp32 tensors, this exact operation order)
    rng = torch.tensor(pc_range, dtype=pts.dtype)
    vs = torch.tensor(voxel_size, dtype=pts.dtype)
    centers = coors[:, [2, 1, 0]] * vs + rng[:3] + vs / 2
    return {"points": pts, "voxel_coords": coors, "point_idxes": idx, "point_offsets": pts[:, :3] - centers}


def scatter_mean(feats: torch.Tensor, cmap: torch.Tensor, count: torch.Tensor) -> torch.Tensor:
    """Differentiable mean reduce; forward follows scatter_points_cuda.cu:45-61, its autograd
    derivative is exactly add_reduce_traceback_grad_kernel (scatter_points_cuda_kernel.cuh:114-141)."""
    m = count.shape[0]
    acc = torch.zeros((m, feats.shape[1]), dtype=feats.dtype).index_add(0, cmap.long(), feats)
    return acc / count.to(feats.dtype).unsqueeze(-1)


# ----------------------------------------------------------------------------- pillar feature net
def pillar_feature_net(points, coors, state, prefix="embedder.feature_net.", voxel_size=DEFAULT_VOXEL_SIZE,
                       pc_range=DEFAULT_RANGE, training=True, buffers=None):
    """DynamicPillarFeatureNet.forward (encoder.py:430-475) for the DeFlow configuration
    (one PFN layer, cluster + voxel centre decorations, mean reduce).
    -> voxel_feats[M,32], voxel_coors[M,3], point_feats[N,32], cmap, count"""
    vx, vy, vz = voxel_size
    x_off = vx / 2 + pc_range[0]          # encoder.py:257-259 (python doubles)
    y_off = vy / 2 + pc_range[1]
    z_off = vz / 2 + pc_range[2]
    vcoors, cmap, count = ext.unique_pillars(coors.numpy())
    vcoors, cmap, count = torch.from_numpy(vcoors), torch.from_numpy(cmap), torch.from_numpy(count)
    voxel_mean = scatter_mean(points, cmap, count)                 # cluster_scatter, :442
    points_mean = voxel_mean[cmap.long()]                          # map_voxel_center_to_point, :443
    f_cluster = points[:, :3] - points_mean[:, :3]
    f_center = torch.zeros((points.shape[0], 3), dtype=points.dtype)
    f_center[:, 0] = points[:, 0] - (coors[:, 2].to(points.dtype) * vx + x_off)   # :452-457
    f_center[:, 1] = points[:, 1] - (coors[:, 1].to(points.dtype) * vy + y_off)
    f_center[:, 2] = points[:, 2] - (coors[:, 0].to(points.dtype) * vz + z_off)
    feats = torch.cat([points, f_cluster, f_center], dim=-1)
    w = state[prefix + "pfn_layers.0.0.weight"]
    bn = prefix + "pfn_layers.0.1."
    lin = feats @ w.T                                              # Linear(9,32,bias=False) :368
    rm = buffers[bn + "running_mean"] if buffers is not None else None
    rv = buffers[bn + "running_var"] if buffers is not None else None
    if rm is None and not training:
        rm, rv = state[bn + "running_mean"], state[bn + "running_var"]
    bnout = F.batch_norm(lin, rm, rv, state[bn + "weight"], state[bn + "bias"], training, 0.01, 1e-3)  # :370
    point_feats = F.relu(bnout)
    voxel_feats = scatter_mean(point_feats, cmap, count)           # pfn_scatter :468
    return voxel_feats, vcoors, point_feats, cmap, count


def pillars_to_image(voxel_feats, voxel_coors, ny, nx):
    """PointPillarsScatter.forward_single (encoder.py:126-147)."""
    canvas = torch.zeros((voxel_feats.shape[1], ny * nx), dtype=voxel_feats.dtype)
    ind = (voxel_coors[:, 1] * nx + voxel_coors[:, 2]).long()
    canvas = canvas.index_copy(1, ind, voxel_feats.t())
    return canvas.view(1, voxel_feats.shape[1], ny, nx)


def embed(points_b, state, grid, voxel_size, pc_range, training, buffers, sync_world=1):
    """DynamicEmbedder.forward (encoder.py:618-631).

    sync_world = R > 1 restates the reference's default ``sync_bn: true`` (OSF/conf/config.yaml:23, OSF/train.py:128 ->
    torch.nn.SyncBatchNorm) for a batch that holds the samples of R ranks, rank-major: the BatchNorm1d call of local
    sample b pools the points of sample b of EVERY rank, i.e. it equals one batch_norm over their concatenation."""
    nb = points_b.shape[0]
    if sync_world <= 1 or not training:
        infos, imgs = [], []
        for b in range(nb):
            info = voxelize_frame(points_b[b], voxel_size, pc_range)
            vf, vc, _, _, _ = pillar_feature_net(info["points"], info["voxel_coords"], state,
                                                 voxel_size=voxel_size, pc_range=pc_range,
                                                 training=training, buffers=buffers)
            imgs.append(pillars_to_image(vf, vc, grid[0], grid[1]))
            infos.append(info)
        return torch.cat(imgs, dim=0), infos
    assert nb % sync_world == 0
    B = nb // sync_world
    vx, vy, vz = voxel_size
    offs = (vx / 2 + pc_range[0], vy / 2 + pc_range[1], vz / 2 + pc_range[2])
    infos = [voxelize_frame(points_b[b], voxel_size, pc_range) for b in range(nb)]
    imgs = [None] * nb
    w = state["embedder.feature_net.pfn_layers.0.0.weight"]
    bn = "embedder.feature_net.pfn_layers.0.1."
    for b in range(B):                          # local sample index: the same call on every rank
        members = [b + r * B for r in range(sync_world)]
        lins, maps = [], []
        for m in members:
            pts, coors = infos[m]["points"], infos[m]["voxel_coords"]
            vcoors, cmap, count = ext.unique_pillars(coors.numpy())
            vcoors, cmap, count = torch.from_numpy(vcoors), torch.from_numpy(cmap), torch.from_numpy(count)
            mean = scatter_mean(pts, cmap, count)[cmap.long()]
            centre = torch.stack([coors[:, 2].to(pts.dtype) * vx + offs[0], coors[:, 1].to(pts.dtype) * vy + offs[1],
                                  coors[:, 0].to(pts.dtype) * vz + offs[2]], 1)
            lins.append(torch.cat([pts, pts[:, :3] - mean[:, :3], pts[:, :3] - centre], -1) @ w.T)
            maps.append((vcoors, cmap, count))
        rm = buffers[bn + "running_mean"] if buffers is not None else None
        rv = buffers[bn + "running_var"] if buffers is not None else None
        pooled = F.batch_norm(torch.cat(lins, 0), rm, rv, state[bn + "weight"], state[bn + "bias"], True, 0.01, 1e-3)
        o = 0
        for m, lin, (vcoors, cmap, count) in zip(members, lins, maps):
            pf = F.relu(pooled[o:o + lin.shape[0]])
            o += lin.shape[0]
            imgs[m] = pillars_to_image(scatter_mean(pf, cmap, count), vcoors, grid[0], grid[1])
    return torch.cat(imgs, dim=0), infos


# ----------------------------------------------------------------------------- UNet
def conv_bn_gelu(x, state, p, stride, training, buffers):
    """ConvWithNorms (OSF/src/models/basic/__init__.py:61-79)."""
    y = F.conv2d(x, state[p + "conv.weight"], state[p + "conv.bias"], stride=stride, padding=1)
    bn = p + "batchnorm."
    if buffers is not None:
        rm, rv = buffers[bn + "running_mean"], buffers[bn + "running_var"]
    elif not training:
        rm, rv = state[bn + "running_mean"], state[bn + "running_var"]
    else:
        rm = rv = None
    y = F.batch_norm(y, rm, rv, state[bn + "weight"], state[bn + "bias"], training, 0.1, 1e-5)
    return F.gelu(y)


ENC_LAYOUT = (("backbone.encoder_step_1.", 4), ("backbone.encoder_step_2.", 6), ("backbone.encoder_step_3.", 6))


def unet_encoder(x, state, training, buffers):
    """encoder_step_{1,2,3} (unet.py:49-64, 79-85): first conv of each step has stride 2."""
    outs = []
    for prefix, n in ENC_LAYOUT:
        for i in range(n):
            x = conv_bn_gelu(x, state, f"{prefix}{i}.", 2 if i == 0 else 1, training, buffers)
        outs.append(x)
    return outs


def upsample_skip(a, b, state, p):
    """UpsampleSkip.forward (unet.py:33-37): convs have bias, no norm / activation."""
    u1 = F.conv2d(a, state[p + "u1_u2.0.weight"], state[p + "u1_u2.0.bias"])
    u2 = F.interpolate(u1, scale_factor=2, mode="bilinear", align_corners=False)
    u3 = F.conv2d(b, state[p + "u3.weight"], state[p + "u3.bias"])
    u4 = F.conv2d(torch.cat([u2, u3], dim=1), state[p + "u4_u5.0.weight"], state[p + "u4_u5.0.bias"], padding=1)
    return F.conv2d(u4, state[p + "u4_u5.1.weight"], state[p + "u4_u5.1.bias"], padding=1)


def unet(img0, img1, state, training=True, buffers=None):
    """FastFlow3DUNet.forward (unet.py:70-100)."""
    f0, l0, r0 = unet_encoder(img0, state, training, buffers)
    f1, l1, r1 = unet_encoder(img1, state, training, buffers)
    s = upsample_skip(torch.cat([r0, r1], 1), torch.cat([l0, l1], 1), state, "backbone.decoder_step1.")
    t = upsample_skip(s, torch.cat([f0, f1], 1), state, "backbone.decoder_step2.")
    u = upsample_skip(t, torch.cat([img0, img1], 1), state, "backbone.decoder_step3.")
    return F.conv2d(u, state["backbone.decoder_step4.weight"], state["backbone.decoder_step4.bias"], padding=1)


# ----------------------------------------------------------------------------- decoders
def gather_pillar_vectors(before, after, voxel_coords):
    """decoder.py:215-225 -- [N,64] from ``before`` then [N,64] from ``after`` at (y, x)."""
    vc = voxel_coords.long()
    a = after[:, vc[:, 1], vc[:, 2]].T
    b = before[:, vc[:, 1], vc[:, 2]].T
    return torch.cat([b, a], dim=1)


def conv_gru(h, x, state, p="head.gru."):
    """ConvGRU.forward (decoder.py:184-193); k=1 Conv1d == matmul with weight[:, :, 0]."""
    wz, wr, wq = (state[p + f"conv{g}.weight"][:, :, 0] for g in "zrq")
    bz, br, bq = (state[p + f"conv{g}.bias"] for g in "zrq")
    hx = torch.cat([h, x], dim=1)
    z = torch.sigmoid(hx @ wz.T + bz)
    r = torch.sigmoid(hx @ wr.T + br)
    q = torch.tanh(torch.cat([r * h, x], dim=1) @ wq.T + bq)
    return (1 - z) * h + z * q


def gru_decoder_single(before, after, offsets, voxel_coords, state, num_iters=4):
    """ConvGRUDecoder.forward_single (decoder.py:210-237)."""
    h = gather_pillar_vectors(before, after, voxel_coords)
    x = offsets @ state["head.offset_encoder.weight"].T + state["head.offset_encoder.bias"]
    for _ in range(num_iters):
        h = conv_gru(h, x, state)
    y = torch.cat([h, x], dim=1) @ state["head.decoder.0.weight"].T + state["head.decoder.0.bias"]
    return F.gelu(y) @ state["head.decoder.2.weight"].T + state["head.decoder.2.bias"]


def linear_decoder_single(before, after, offsets, voxel_coords, state):
    """LinearDecoder.forward_single (decoder.py:81-104)."""
    h = gather_pillar_vectors(before, after, voxel_coords)
    x = offsets @ state["head.offset_encoder.weight"].T + state["head.offset_encoder.bias"]
    y = torch.cat([h, x], dim=1) @ state["head.decoder.0.weight"].T + state["head.decoder.0.bias"]
    return F.gelu(y) @ state["head.decoder.2.weight"].T + state["head.decoder.2.bias"]


# ----------------------------------------------------------------------------- model
def deflow_forward(batch: Dict, state: Dict[str, torch.Tensor], voxel_size=DEFAULT_VOXEL_SIZE,
                   pc_range=DEFAULT_RANGE, grid=(512, 512), decoder="gru", num_iters=4,
                   training=True, buffers=None, return_internals=False, sync_world=1) -> Dict:
    """DeFlow.forward (OSF/src/models/deflow.py:49-114) / FastFlow3D.forward (fastflow3d.py:74-103).
    sync_world = R: the batch holds the samples of R data-parallel ranks (rank-major) trained with SyncBatchNorm: BatchNorm2d
    over the whole batch IS the pooled statistic; the per-sample BatchNorm1d calls are pooled across ranks (see embed)."""
    bsz = len(batch["pose0"])
    pose_flows, pc0s = [], []
    for b in range(bsz):
        if "ego_motion" in batch:
            t01 = batch["ego_motion"][b]
        else:
            t01 = cal_pose0to1(batch["pose0"][b], batch["pose1"][b])
        warped, pf = ego_compensate(batch["pc0"][b], t01)
        pose_flows.append(pf)
        pc0s.append(warped)
    pc0s = torch.stack(pc0s, 0)
    img0, info0 = embed(pc0s, state, grid, voxel_size, pc_range, training, buffers, sync_world)
    img1, info1 = embed(batch["pc1"], state, grid, voxel_size, pc_range, training, buffers, sync_world)
    feat = unet(img0, img1, state, training, buffers)
    before = torch.cat([img0, img1], dim=1)
    flows = []
    for b in range(bsz):
        if decoder == "gru":
            flows.append(gru_decoder_single(before[b], feat[b], info0[b]["point_offsets"],
                                            info0[b]["voxel_coords"], state, num_iters))
        else:
            flows.append(linear_decoder_single(before[b], feat[b], info0[b]["point_offsets"],
                                               info0[b]["voxel_coords"], state))
    res = {
        "flow": flows,
        "pose_flow": pose_flows,
        "pc0_valid_point_idxes": [e["point_idxes"] for e in info0],
        "pc0_points_lst": [e["points"] for e in info0],
        "pc1_valid_point_idxes": [e["point_idxes"] for e in info1],
        "pc1_points_lst": [e["points"] for e in info1],
        "num_occupied_voxels": [feat.shape[-1] * feat.shape[-2]],
    }
    if return_internals:
        res["_img0"], res["_img1"], res["_unet"], res["_info0"], res["_info1"] = img0, img1, feat, info0, info1
    return res


# ----------------------------------------------------------------------------- losses
def deflow_loss(est_flow: torch.Tensor, gt_flow: torch.Tensor) -> torch.Tensor:
    """deflowLoss (OSF/src/lossfuncs.py:102-125)."""
    ok = ~gt_flow.isnan() & ~est_flow.isnan() & ~gt_flow.isinf() & ~est_flow.isinf()
    pred = est_flow[ok].reshape(-1, 3)
    gt = gt_flow[ok].reshape(-1, 3)
    speed = gt.norm(dim=1, p=2) / 0.1
    err = torch.linalg.vector_norm(pred - gt, dim=-1)
    total = 0.0
    for sel in (speed > 1.0, speed < 0.4, (speed >= 0.4) & (speed <= 1.0)):
        part = err[sel].mean()
        if not bool(part.isnan()):
            total = total + part
    return total


def ff3d_loss(est_flow, gt_flow, classes) -> torch.Tensor:
    """ff3dLoss (OSF/src/lossfuncs.py:148-157)."""
    err = torch.linalg.vector_norm(est_flow - gt_flow, dim=-1)
    return (err * ((classes > 0).float() * 0.9 + 0.1)).mean()


def zeroflow_loss(est_flow, gt_flow) -> torch.Tensor:
    """zeroflowLoss (OSF/src/lossfuncs.py:128-145)."""
    ok = ~gt_flow.isnan() & ~est_flow.isnan() & ~gt_flow.isinf() & ~est_flow.isinf()
    pred = est_flow[ok].reshape(-1, 3)
    gt = gt_flow[ok].reshape(-1, 3)
    err = torch.linalg.vector_norm(pred - gt, dim=-1)
    speed = torch.linalg.vector_norm(gt, dim=-1) * 10.0
    scale = torch.max(torch.ones_like(speed) * 0.1, torch.min(1.8 * speed - 0.8, torch.ones_like(speed)))
    return (err * scale).mean()


def training_step_loss(batch, res, loss="deflowLoss") -> torch.Tensor:
    """The arithmetic of ModelWrapper.training_step (OSF/src/trainer.py:116-152):
    gt = flow[idx] - pose_flow[idx]; per-sample losses are SUMMED over the batch."""
    total = 0.0
    for b in range(len(batch["pose0"])):
        idx = res["pc0_valid_point_idxes"][b]
        gt = batch["flow"][b][idx] - res["pose_flow"][b][idx]
        if loss == "deflowLoss":
            total = total + deflow_loss(res["flow"][b], gt)
        elif loss == "zeroflowLoss":
            total = total + zeroflow_loss(res["flow"][b], gt)
        else:
            total = total + ff3d_loss(res["flow"][b], gt, batch["flow_category_indices"][b][idx])
    return total


# ----------------------------------------------------------------------------- parameters
def param_shapes(decoder="gru") -> Dict[str, tuple]:
    """state_dict layout of the reference model (checkpoint row)."""
    s: Dict[str, tuple] = {}
    p = "embedder.feature_net.pfn_layers.0."
    s[p + "0.weight"] = (32, 9)
    for k, shp in (("weight", (32,)), ("bias", (32,)), ("running_mean", (32,)), ("running_var", (32,)),
                   ("num_batches_tracked", ())):
        s[p + "1." + k] = shp
    chans = {"backbone.encoder_step_1.": (32, 64), "backbone.encoder_step_2.": (64, 128),
             "backbone.encoder_step_3.": (128, 256)}
    for prefix, n in ENC_LAYOUT:
        cin, cout = chans[prefix]
        for i in range(n):
            q = f"{prefix}{i}."
            s[q + "conv.weight"] = (cout, cin if i == 0 else cout, 3, 3)
            s[q + "conv.bias"] = (cout,)
            for k, shp in (("weight", (cout,)), ("bias", (cout,)), ("running_mean", (cout,)),
                           ("running_var", (cout,)), ("num_batches_tracked", ())):
                s[q + "batchnorm." + k] = shp
    for name, (skip, lat, out) in (("backbone.decoder_step1.", (512, 256, 256)),
                                   ("backbone.decoder_step2.", (256, 128, 128)),
                                   ("backbone.decoder_step3.", (128, 64, 64))):
        s[name + "u1_u2.0.weight"] = (lat, skip, 1, 1)
        s[name + "u1_u2.0.bias"] = (lat,)
        s[name + "u3.weight"] = (lat, lat, 1, 1)
        s[name + "u3.bias"] = (lat,)
        s[name + "u4_u5.0.weight"] = (out, 2 * lat, 3, 3)
        s[name + "u4_u5.0.bias"] = (out,)
        s[name + "u4_u5.1.weight"] = (out, out, 3, 3)
        s[name + "u4_u5.1.bias"] = (out,)
    s["backbone.decoder_step4.weight"] = (64, 64, 3, 3)
    s["backbone.decoder_step4.bias"] = (64,)
    if decoder == "gru":
        s["head.offset_encoder.weight"] = (64, 3)
        s["head.offset_encoder.bias"] = (64,)
        for g in "zrq":
            s[f"head.gru.conv{g}.weight"] = (128, 192, 1)
            s[f"head.gru.conv{g}.bias"] = (128,)
        s["head.decoder.0.weight"] = (32, 192)
    else:
        s["head.offset_encoder.weight"] = (128, 3)
        s["head.offset_encoder.bias"] = (128,)
        s["head.decoder.0.weight"] = (32, 256)
    s["head.decoder.0.bias"] = (32,)
    s["head.decoder.2.weight"] = (3, 32)
    s["head.decoder.2.bias"] = (3,)
    return s


def random_state(seed=0, decoder="gru", scale=1.0) -> Dict[str, torch.Tensor]:
    """Random parameters with the reference's shapes: Xavier-uniform-like weights, non-trivial
    biases and BN affine terms so that every term of the arithmetic is exercised."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in param_shapes(decoder).items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            out[k] = torch.rand(shp, generator=g) * 0.5 + 0.75
        elif k.endswith("running_mean"):
            out[k] = torch.randn(shp, generator=g) * 0.1
        elif len(shp) >= 2:
            fan_out = shp[0] * int(np.prod(shp[2:])) if len(shp) > 2 else shp[0]
            fan_in = int(np.prod(shp[1:]))
            bound = scale * math.sqrt(6.0 / (fan_in + fan_out))
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif "batchnorm.weight" in k or k.endswith("pfn_layers.0.1.weight"):
            out[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            out[k] = 0.05 * torch.randn(shp, generator=g)
    return out
