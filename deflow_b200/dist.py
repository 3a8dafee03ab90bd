"""Data-parallel plumbing: one process per GPU, gradients averaged with ONE all-reduce of a flat fp32 buffer.

The reference trains with Lightning DDP (OSF/train.py:125): bucketed NCCL all-reduce of 6 891 939 fp32 gradients
(27.6 MB) overlapped with backward.  On an NVSwitch domain that message takes ~0.1 ms against a >40 ms step, so the
path shards by frame pair with a single collective per step and no overlap machinery (DESIGN.md "multi-GPU").
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
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


class GradAverager:
    """Gradient mean across ranks with ONE collective on a flat fp32 buffer.

    ``zero()`` drops the ``.grad`` tensors (autograd then *assigns* each gradient instead of launching an add kernel
    per parameter); ``average()`` packs them into the flat buffer (one multi-tensor copy), all-reduces it when there
    is more than one rank, and re-points every ``.grad`` at its slice, so clipping and the optimizer work on views."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()
        self.zero()

    def zero(self):
        for p in self.params:
            p.grad = None

    def average(self):
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v
        w = world_size()
        if w > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(w)

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
