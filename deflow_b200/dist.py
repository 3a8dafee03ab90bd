"""Data-parallel plumbing: one process per GPU, gradients averaged with ONE all-reduce of a flat fp32 buffer.

The reference trains with Lightning DDP (OSF/train.py:125): bucketed NCCL all-reduce of 6 891 939 fp32 gradients
(27.6 MB) overlapped with backward.  On an NVSwitch domain that message takes ~0.1 ms against a >40 ms step, so the
path shards by frame pair with one collective per step (ReduceOp.AVG on the flat buffer); an overlapped two-slice variant is
opt-in and measured slower on NVSwitch (DESIGN.md "multi-GPU").
Works with any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Iterable, List

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend: str = "nccl", device=None):
    world, rank, _ = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return world, rank


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


def broadcast_module(module: torch.nn.Module, src: int = 0):
    """Replicate parameters and buffers of rank `src` (DDP does this at construction)."""
    if world_size() == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t, src)


def _avg_inplace(t: torch.Tensor, async_op: bool = False):
    """Mean over ranks: ReduceOp.AVG on NCCL (no separate divide pass); gloo has no AVG, so SUM + divide there."""
    if dist.get_backend() == "nccl":
        return dist.all_reduce(t, op=dist.ReduceOp.AVG, async_op=async_op)
    work = dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=False)
    t.div_(dist.get_world_size())
    return None if not async_op else work


class GradAverager:
    """Gradient mean across ranks on a flat fp32 buffer, in at most two collectives.

    ``zero()`` drops the ``.grad`` tensors (autograd then *assigns* each gradient instead of launching an add kernel
    per parameter); ``average()`` packs them into the flat buffer (one multi-tensor copy), averages it over the ranks
    and re-points every ``.grad`` at its slice, so clipping and the optimizer work on views.

    ``plan_early_slice(model)`` + ``arm_early_slice()`` (TrainStep, world size > 1): the parameters whose gradients the
    backward finishes FIRST -- the flow head and the UNet decoder, the tail of ``model.parameters()`` -- form one
    contiguous slice of the flat buffer; a post-accumulate hook on each of them counts down, and when the last one has
    its gradient the slice is packed and its all-reduce is issued asynchronously (NCCL's own stream), overlapping the
    encoder + pillar-feature-net backward that is still running.  ``average()`` then only has the head of the buffer
    left."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, self.offsets, off = [], [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            self.offsets.append(off)
            off += p.numel()
        self.early_from = None      # index into self.params where the early slice starts
        self._early_left = -1
        self._early_work = None
        self._early_done = False
        self._hooks = []
        self.zero()

    def zero(self):
        for p in self.params:
            p.grad = None

    # -------------------------------------------------------------------------------------- early slice
    def plan_early_slice(self, model: torch.nn.Module, late_prefixes=("embedder.", "backbone.encoder_step")):
        """Everything after the last parameter whose name starts with one of ``late_prefixes`` is 'early'."""
        names = {id(p): n for n, p in model.named_parameters()}
        last_late = -1
        for i, p in enumerate(self.params):
            if names.get(id(p), "").startswith(tuple(late_prefixes)):
                last_late = i
        self.early_from = last_late + 1
        for h in self._hooks:
            h.remove()
        self._hooks = []
        if self.early_from >= len(self.params):
            self.early_from = None
            return
        for p in self.params[self.early_from:]:
            self._hooks.append(p.register_post_accumulate_grad_hook(self._on_early_grad))

    def arm_early_slice(self):
        if self.early_from is not None:
            self._early_left = len(self.params) - self.early_from
            self._early_work, self._early_done = None, False

    def _on_early_grad(self, _p):
        if self._early_left <= 0:
            return
        self._early_left -= 1
        if self._early_left == 0 and world_size() > 1:
            i0 = self.early_from
            self._pack(range(i0, len(self.params)))
            self._early_work = _avg_inplace(self.flat[self.offsets[i0]:], async_op=True)
            self._early_done = True

    # -------------------------------------------------------------------------------------- pack + average
    def _pack(self, indices):
        src, dst = [], []
        for i in indices:
            p, v = self.params[i], self.views[i]
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def average(self):
        early = self._early_done and self.early_from is not None
        n_late = self.early_from if early else len(self.params)
        self._pack(range(n_late))
        for p, v in zip(self.params, self.views):
            p.grad = v
        if world_size() > 1:
            if early:
                if n_late > 0:
                    _avg_inplace(self.flat[:self.offsets[self.early_from]])
                if self._early_work is not None:
                    self._early_work.wait()
            else:
                _avg_inplace(self.flat)
        self._early_left, self._early_work, self._early_done = -1, None, False

    @property
    def nbytes(self):
        return self.flat.numel() * 4


def max_over_ranks(value: float, device) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier():
    if world_size() > 1:
        dist.barrier()


class StatSync:
    """Sums small statistics tensors over the ranks of the default process group (SyncBatchNorm exchange)."""

    def __init__(self):
        self.world = world_size()
        self.calls = 0
        self.bytes = 0

    def all_reduce_sum(self, t: torch.Tensor):
        self.calls += 1
        self.bytes += t.numel() * t.element_size()
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t


def enable_sync_bn(model: torch.nn.Module, sync=None):
    """BatchNorm batch statistics over ALL ranks -- the reference's default ``sync_bn: true`` (OSF/conf/config.yaml:23,
    OSF/train.py:128: Lightning converts every BatchNorm to torch.nn.SyncBatchNorm).  Marks every BatchNorm module of the
    model; the kernels behind them (conv.conv_bn_gelu for the 16 encoder BatchNorm2d, ops.pillar_feature_net for the
    pillar feature net's BatchNorm1d) then exchange their statistics through ``sync``: forward = per-channel sum /
    sum of squares (BatchNorm2d) or the 54 feature moments + point counts of all 2B frames (BatchNorm1d); backward = the
    two per-channel sums of the BatchNorm backward.  Parameter gradients stay local sums and are averaged with all the
    other gradients.  ``sync=None`` -> a StatSync over the default process group (no-op exchange at world size 1).
    Returns the sync object (``.calls`` / ``.bytes`` count the exchanges)."""
    sync = sync if sync is not None else StatSync()
    n = 0
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.dfb_sync = sync
            n += 1
    model.dfb_sync = sync
    return sync


def disable_sync_bn(model: torch.nn.Module):
    for m in model.modules():
        if hasattr(m, "dfb_sync"):
            del m.dfb_sync
