"""numpy restatement of the three ``mmcv._ext`` functions on the DeFlow path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  REF = /root/reference/
OpenSceneFlow/assets/cuda/mmcv.

Functions
---------
dynamic_voxelize_forward        REF/voxelization_cuda_kernel.cuh:13-50,
                                REF/voxelization_cuda.cu:246-286, REF/voxelize.py:66-74
dynamic_point_to_voxel_forward  REF/scatter_points_cuda.cu:9-66,
                                REF/scatter_points_cuda_kernel.cuh:91-112
dynamic_point_to_voxel_backward REF/scatter_points_cuda.cu:68-132,
                                REF/scatter_points_cuda_kernel.cuh:114-185

All integer outputs (coords, maps, counts, ordering) are exact restatements;
floating sums are accumulated in float64 and rounded once to the feature dtype
(the reference uses unordered fp32 atomics, so only tolerance parity is defined
for them -- REF/scatter_points.py:77-79 quotes 5e-7).
"""
from __future__ import annotations

import numpy as np

REDUCE_TYPES = ("sum", "mean", "max")


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _float_to_int_cuda(v: np.ndarray) -> np.ndarray:
    """``(int) floorf(x)`` as the GPU evaluates it: cvt.rzi.s32.f32 saturates and
    maps NaN to 0 (REF/voxelization_cuda_kernel.cuh:26 ``int c_x = floorf(...)``)."""
    v = np.floor(v.astype(np.float32))
    out = np.zeros(v.shape, dtype=np.int64)
    finite = np.isfinite(v)
    out[finite] = np.clip(v[finite], -2147483648.0, 2147483647.0).astype(np.int64)
    out[np.isposinf(v)] = 2147483647
    out[np.isneginf(v)] = -2147483648
    return out.astype(np.int32)


def grid_size(voxel_size, coors_range) -> tuple:
    """``round((max - min) / voxel)`` in fp32, half away from zero
    (REF/voxelization_cuda.cu:269-271).  Returns (grid_x, grid_y, grid_z)."""
    vs = _f32(voxel_size)
    rng = _f32(coors_range)
    g = (rng[3:6] - rng[0:3]) / vs
    g = np.where(g >= 0, np.floor(g.astype(np.float64) + 0.5), np.ceil(g.astype(np.float64) - 0.5))
    return int(g[0]), int(g[1]), int(g[2])


def dynamic_voxelize_forward(points: np.ndarray, voxel_size, coors_range,
                             coors: np.ndarray | None = None) -> np.ndarray:
    """points f32[N, F>=3] -> coors i32[N,3] in (z, y, x) order.

    Test order is x, then y, then z with early exit; a failed x test writes only
    column 0, a failed y test columns 0-1, a failed z test all three (REF/
    voxelization_cuda_kernel.cuh:27-43).  ``coors`` is the caller's zero-initialised
    output (REF/voxelize.py:67); untouched slots keep whatever it held."""
    pts = _f32(points)
    n = pts.shape[0]
    if coors is None:
        coors = np.zeros((n, 3), dtype=np.int32)
    assert coors.shape == (n, 3) and coors.dtype == np.int32
    vs = _f32(voxel_size)
    rng = _f32(coors_range)
    gx, gy, gz = grid_size(voxel_size, coors_range)
    with np.errstate(invalid="ignore", over="ignore", divide="ignore"):
        cx = _float_to_int_cuda((pts[:, 0] - rng[0]) / vs[0])
        cy = _float_to_int_cuda((pts[:, 1] - rng[1]) / vs[1])
        cz = _float_to_int_cuda((pts[:, 2] - rng[2]) / vs[2])
    bad_x = (cx < 0) | (cx >= gx)
    bad_y = ~bad_x & ((cy < 0) | (cy >= gy))
    bad_z = ~bad_x & ~bad_y & ((cz < 0) | (cz >= gz))
    ok = ~(bad_x | bad_y | bad_z)
    coors[bad_x, 0] = -1
    coors[bad_y, 0] = -1
    coors[bad_y, 1] = -1
    coors[bad_z, :] = -1
    coors[ok, 0] = cz[ok]
    coors[ok, 1] = cy[ok]
    coors[ok, 2] = cx[ok]
    return coors


def unique_pillars(coors: np.ndarray):
    """The integer half of dynamic_point_to_voxel_forward
    (REF/scatter_points_cuda.cu:24-37): rows with any negative component become
    (-1,-1,-1); lexicographic sorted unique with inverse and counts; a leading
    negative row is stripped and the map shifted so those points map to -1."""
    coors = np.asarray(coors, dtype=np.int32)
    clean = coors.copy()
    clean[(coors < 0).any(axis=1)] = -1
    if clean.shape[0] == 0:
        return (np.zeros((0, coors.shape[1]), np.int32), np.zeros((0,), np.int32),
                np.zeros((0,), np.int32))
    out_coors, inverse, counts = np.unique(clean, axis=0, return_inverse=True, return_counts=True)
    inverse = inverse.reshape(-1)
    if out_coors[0, 0] < 0:
        out_coors = out_coors[1:]
        counts = counts[1:]
        inverse = inverse - 1
    return out_coors.astype(np.int32), inverse.astype(np.int32), counts.astype(np.int32)


def dynamic_point_to_voxel_forward(feats: np.ndarray, coors: np.ndarray, reduce_type: str = "max"):
    """-> (voxel_feats[M,C], voxel_coors i32[M,3], point2voxel_map i32[N], count i32[M])."""
    if reduce_type not in REDUCE_TYPES:
        raise RuntimeError("do not support reduce type " + str(reduce_type))  # scatter_points.cpp:32
    feats = np.asarray(feats)
    coors = np.asarray(coors, dtype=np.int32)
    n, c = feats.shape
    if n == 0:  # scatter_points_cuda.cu:15-18
        return feats.copy(), coors.copy(), np.zeros((0,), np.int32), np.zeros((0,), np.int32)
    out_coors, cmap, count = unique_pillars(coors)
    m = out_coors.shape[0]
    valid = cmap >= 0
    if reduce_type == "max":
        red = np.full((m, c), -np.inf, dtype=feats.dtype)
        np.maximum.at(red, cmap[valid], feats[valid])
    else:
        acc = np.zeros((m, c), dtype=np.float64)
        np.add.at(acc, cmap[valid], feats[valid].astype(np.float64))
        if reduce_type == "mean":
            # the reference divides the fp32 sum by float(count) (scatter_points_cuda.cu:59-60)
            red = (acc.astype(feats.dtype) / count[:, None].astype(feats.dtype)).astype(feats.dtype)
        else:
            red = acc.astype(feats.dtype)
    return red, out_coors, cmap, count


def dynamic_point_to_voxel_backward(grad_reduced: np.ndarray, feats: np.ndarray, reduced_feats: np.ndarray,
                                    cmap: np.ndarray, count: np.ndarray, reduce_type: str) -> np.ndarray:
    """-> grad_feats[N,C] (zero where map == -1)."""
    if reduce_type not in REDUCE_TYPES:
        raise RuntimeError("do not support reduce type " + str(reduce_type))
    n, c = feats.shape
    grad = np.zeros((n, c), dtype=grad_reduced.dtype)  # scatter_points_cuda.cu:77
    m = reduced_feats.shape[0]
    if n == 0 or m == 0:
        return grad
    valid = cmap >= 0
    if reduce_type == "sum":
        grad[valid] = grad_reduced[cmap[valid]]
    elif reduce_type == "mean":
        grad[valid] = grad_reduced[cmap[valid]] / count[cmap[valid]][:, None].astype(grad_reduced.dtype)
    else:
        # smallest point index attaining the max takes the gradient (atomicMin trace-back,
        # scatter_points_cuda_kernel.cuh:143-185)
        src = np.full((m, c), n, dtype=np.int64)
        idx = np.nonzero(valid)[0]
        hit = feats[idx] == reduced_feats[cmap[idx]]
        for col in range(c):
            rows = idx[hit[:, col]]
            np.minimum.at(src[:, col], cmap[rows], rows)
        mm, cc = np.nonzero(src < n)
        grad[src[mm, cc], cc] = grad_reduced[mm, cc]
    return grad


def hard_voxelize_forward(points, voxel_size, coors_range, max_points, max_voxels):
    """HardVoxelizeForwardCUDAKernelLauncher, deterministic (voxelization_cuda.cu:8-148; kernels
    voxelization_cuda_kernel.cuh:52-170): voxels are numbered in order of first appearance (determin_voxel_num), a voxel
    keeps its first max_points points (point_to_voxelidx_kernel), voxels beyond max_voxels are dropped.
    -> voxels f32[M,max_points,F], coors i32[M,3], num_points_per_voxel i32[M]."""
    points = np.ascontiguousarray(points, np.float32)
    n, f = points.shape
    tmp = dynamic_voxelize_forward(points, voxel_size, coors_range, np.zeros((n, 3), np.int32))
    voxels = np.zeros((max_voxels, max_points, f), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    seen = {}            # coordinate -> [voxel id or -1, points seen so far]
    voxel_num = 0
    for i in range(n):
        if tmp[i, 0] == -1:
            continue
        key = (int(tmp[i, 0]), int(tmp[i, 1]), int(tmp[i, 2]))
        ent = seen.get(key)
        if ent is None:
            vid = voxel_num if voxel_num < max_voxels else -1
            if vid >= 0:
                voxel_num += 1
                coors[vid] = tmp[i]
            ent = seen[key] = [vid, 0]
        rank = ent[1]
        ent[1] += 1
        if rank < max_points and ent[0] >= 0:
            voxels[ent[0], rank] = points[i]
            num[ent[0]] += 1
    return voxels[:voxel_num], coors[:voxel_num], num[:voxel_num]
