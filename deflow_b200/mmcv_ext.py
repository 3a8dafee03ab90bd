"""Drop-in for the reference's vendored mmcv slice (OSF/assets/cuda/mmcv): the ``_ext`` namespace with
the four names ``mmcv._ext`` exports (OSF/assets/cuda/mmcv/pybind.cpp:32-50) and the two module
classes ``Voxelization`` / ``DynamicScatter`` the models import (OSF/assets/cuda/mmcv/__init__.py:2-3).

``install_as_mmcv_ext()`` registers the namespace as ``sys.modules['mmcv._ext']`` so that the
reference's own ``voxelize.py`` / ``scatter_points.py`` run unmodified on top of these kernels
(see INTEGRATION.md).
"""
from __future__ import annotations

import sys
import types
from typing import Tuple

import torch
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops

_ext = types.ModuleType("mmcv._ext")
_ext.dynamic_voxelize_forward = ops.dynamic_voxelize_forward
_ext.dynamic_point_to_voxel_forward = ops.dynamic_point_to_voxel_forward
_ext.dynamic_point_to_voxel_backward = ops.dynamic_point_to_voxel_backward
_ext.hard_voxelize_forward = ops.hard_voxelize_forward


def install_as_mmcv_ext():
    """Make ``importlib.import_module('mmcv._ext')`` (OSF/assets/cuda/mmcv/scatter_points.py:11-19)
    resolve to the deflow_b200 kernels."""
    pkg = sys.modules.get("mmcv") or types.ModuleType("mmcv")
    pkg._ext = _ext
    sys.modules["mmcv"] = pkg
    sys.modules["mmcv._ext"] = _ext
    return _ext


class _DynamicScatter(Function):
    """OSF/assets/cuda/mmcv/scatter_points.py:22-67."""

    @staticmethod
    def forward(ctx, feats, coors, reduce_type="max"):
        voxel_feats, voxel_coors, point2voxel_map, voxel_points_count = _ext.dynamic_point_to_voxel_forward(
            feats, coors, reduce_type)
        ctx.reduce_type = reduce_type
        ctx.save_for_backward(feats, voxel_feats, point2voxel_map, voxel_points_count)
        ctx.mark_non_differentiable(voxel_coors)
        return voxel_feats, voxel_coors

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_voxel_feats, grad_voxel_coors=None):
        feats, voxel_feats, point2voxel_map, voxel_points_count = ctx.saved_tensors
        grad_feats = torch.empty_like(feats)  # fully written by the kernel (zeros where map == -1)
        _ext.dynamic_point_to_voxel_backward(grad_feats, grad_voxel_feats.contiguous(), feats, voxel_feats,
                                             point2voxel_map, voxel_points_count, ctx.reduce_type)
        return grad_feats, None, None


dynamic_scatter = _DynamicScatter.apply


class DynamicScatter(nn.Module):
    """OSF/assets/cuda/mmcv/scatter_points.py:70-154 (same ctor, same forward contract)."""

    def __init__(self, voxel_size, point_cloud_range, average_points: bool):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.average_points = average_points

    def forward_single(self, points, coors) -> Tuple[torch.Tensor, torch.Tensor]:
        reduce = "mean" if self.average_points else "max"
        return dynamic_scatter(points.contiguous(), coors.contiguous(), reduce)

    def forward(self, points, coors) -> Tuple[torch.Tensor, torch.Tensor]:
        if coors.size(-1) == 3:
            return self.forward_single(points, coors)
        batch_size = int(coors[-1, 0]) + 1  # scatter_points.py:131-146
        voxels, voxel_coors = [], []
        for i in range(batch_size):
            inds = torch.where(coors[:, 0] == i)
            voxel, voxel_coor = self.forward_single(points[inds], coors[inds][:, 1:])
            voxel_coors.append(torch.nn.functional.pad(voxel_coor, (1, 0), mode="constant", value=i))
            voxels.append(voxel)
        return torch.cat(voxels, dim=0), torch.cat(voxel_coors, dim=0)

    def __repr__(self):
        return (f"{self.__class__.__name__}(voxel_size={self.voxel_size}, point_cloud_range="
                f"{self.point_cloud_range}, average_points={self.average_points})")


class _Voxelization(Function):
    """OSF/assets/cuda/mmcv/voxelize.py:22-112: dynamic branch (max_points == -1 or max_voxels == -1) and hard branch."""

    @staticmethod
    def forward(ctx, points, voxel_size, coors_range, max_points=35, max_voxels=20000, deterministic=True):
        if max_points == -1 or max_voxels == -1:
            coors = points.new_zeros(size=(points.size(0), 3), dtype=torch.int)
            _ext.dynamic_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                                          torch.tensor(coors_range, dtype=torch.float), coors, NDim=3)
            return coors
        # voxelize.py:86-111
        voxels = points.new_zeros(size=(max_voxels, max_points, points.size(1)))
        coors = points.new_zeros(size=(max_voxels, 3), dtype=torch.int)
        num_points_per_voxel = points.new_zeros(size=(max_voxels,), dtype=torch.int)
        voxel_num = torch.zeros(size=(), dtype=torch.long)
        _ext.hard_voxelize_forward(points, torch.tensor(voxel_size, dtype=torch.float),
                                   torch.tensor(coors_range, dtype=torch.float), voxels, coors, num_points_per_voxel,
                                   voxel_num, max_points=max_points, max_voxels=max_voxels, NDim=3,
                                   deterministic=deterministic)
        return voxels[:voxel_num], coors[:voxel_num], num_points_per_voxel[:voxel_num]


voxelization = _Voxelization.apply


class Voxelization(nn.Module):
    """OSF/assets/cuda/mmcv/voxelize.py:115-189."""

    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
        super().__init__()
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.max_num_points = max_num_points
        self.max_voxels = max_voxels if isinstance(max_voxels, tuple) else (max_voxels, max_voxels)
        self.deterministic = deterministic
        pcr = torch.tensor(point_cloud_range, dtype=torch.float32)
        vs = torch.tensor(voxel_size, dtype=torch.float32)
        grid = torch.round((pcr[3:] - pcr[:3]) / vs).long()
        self.grid_shape = grid
        self.pcd_shape = [*grid.tolist()[:2], 1][::-1]

    def forward(self, input):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        return voxelization(input, self.voxel_size, self.point_cloud_range, self.max_num_points, max_voxels,
                            self.deterministic)
