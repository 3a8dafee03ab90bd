"""Per-point flow decoders with the reference's parameter tree (OSF/src/models/basic/decoder.py).

state_dict keys: head.offset_encoder.*, head.gru.conv{z,r,q}.* (Conv1d k=1, weight [128,192,1]),
head.decoder.{0,2}.*.  The pillar->point gather runs for all samples at once on NHWC images
(deflow_b200/csrc/decoder_gather.cu); the reference loops over samples and gathers channel-strided
NCHW columns (decoder.py:215-225, 239-253).
"""
from __future__ import annotations

import os
from typing import List

import torch
from torch import nn

from . import ops


class ConvGRU(nn.Module):
    """decoder.py:177-193: z, r, q gates as k=1 Conv1d over the [h(128), x(64)] channel stack."""

    def __init__(self, input_dim: int = 64, hidden_dim: int = 128):
        super().__init__()
        self.convz = nn.Conv1d(input_dim + hidden_dim, hidden_dim, 1)
        self.convr = nn.Conv1d(input_dim + hidden_dim, hidden_dim, 1)
        self.convq = nn.Conv1d(input_dim + hidden_dim, hidden_dim, 1)

    def forward(self, h, x):
        """Reference layout: h[N,128,1], x[N,64,1] -> h'[N,128,1] (decoder.py:184-193): one GRU iteration on the
        tensor-core gate GEMMs + gate kernels, split-precision ("bf16x3") fp32 arithmetic, with its own backward.
        (ConvGRUDecoder runs all iterations inside one fused launch sequence instead of calling this.)"""
        from . import gru
        return gru.gru_step(h.squeeze(-1), x.squeeze(-1), self).unsqueeze(-1)


class _GatherDecoder(nn.Module):
    def _forward_tensor_core(self, img_nhwc, unet_nhwc, idx, B, n_rows, offsets, iters, sink=None):
        """bf16 perf mode: gather -> (GRU iterations) -> MLP head on the tensor cores (deflow_b200/gru.py)."""
        from . import gru
        n_pad = max((n_rows + 7) // 8 * 8, 8)
        if self.compute_dtype == torch.float32:
            # parity mode: fp32 gate tensors, split-precision ("bf16x3") tensor-core GEMMs
            h0 = ops.decoder_gather(img_nhwc, unet_nhwc, idx, B, n_rows, torch.float32, n_pad, sink)
            return gru.decode(h0, offsets, n_rows, self, iters, parity=True)
        if iters > 0 and os.environ.get("DFB_GRU", "fused") == "fused":
            # persistent fused kernels (csrc/gru_fused.cu); DFB_GRU=unfused selects the GEMM + elementwise launches
            h0 = ops.decoder_gather(img_nhwc, unet_nhwc, idx, B, n_rows, torch.bfloat16, n_pad, sink)
            return gru.decode_fused(h0, offsets, n_rows, self)
        h0 = ops.decoder_gather(img_nhwc, unet_nhwc, idx, B, n_rows, torch.float32, n_pad, sink)
        return gru.decode(h0, offsets, n_rows, self, iters)

    def forward(self, before_pseudoimages, after_pseudoimages, voxelizer_infos) -> List[torch.Tensor]:
        """Reference signature (decoder.py:239-253 / 106-119): before [B,64,H,W] = cat(img0, img1),
        after [B,64,H,W], list of per-sample dicts.  Rebuilds the flat index from the dicts."""
        B = before_pseudoimages.shape[0]
        img = torch.cat([before_pseudoimages[:, :32], before_pseudoimages[:, 32:]], dim=0)
        img = img.permute(0, 2, 3, 1).contiguous()
        after = after_pseudoimages.permute(0, 2, 3, 1).contiguous()
        idx = _index_from_infos(voxelizer_infos, img.shape[1], img.shape[2])
        n = idx.pt_off(B)
        flow = self.forward_flat(img, after, idx, B, n, torch.cat([e["point_offsets"] for e in voxelizer_infos], 0))
        return [flow[idx.pt_off(b):idx.pt_off(b + 1)] for b in range(B)]


class ConvGRUDecoder(_GatherDecoder):
    """decoder.py:195-253."""

    def __init__(self, pseudoimage_channels: int = 64, num_iters: int = 4):
        super().__init__()
        self.offset_encoder = nn.Linear(3, pseudoimage_channels)
        self.gru = ConvGRU(input_dim=pseudoimage_channels, hidden_dim=pseudoimage_channels * 2)
        self.decoder = nn.Sequential(nn.Linear(pseudoimage_channels * 3, pseudoimage_channels // 2), nn.GELU(),
                                     nn.Linear(pseudoimage_channels // 2, 3))
        self.num_iters = num_iters
        self.compute_dtype = torch.float32

    def forward_flat(self, img_nhwc, unet_nhwc, idx, B, n_rows, offsets=None, sink=None):
        """All pc0 points of the batch at once -> flow [n_rows, 3]."""
        offsets = idx.pt_offs[:n_rows] if offsets is None else offsets
        return self._forward_tensor_core(img_nhwc, unet_nhwc, idx, B, n_rows, offsets, self.num_iters, sink)


class LinearDecoder(_GatherDecoder):
    """decoder.py:71-119."""

    def __init__(self, pseudoimage_channels: int = 64):
        super().__init__()
        self.offset_encoder = nn.Linear(3, 128)
        self.decoder = nn.Sequential(nn.Linear(pseudoimage_channels * 4, 32), nn.GELU(), nn.Linear(32, 3))
        self.compute_dtype = torch.float32

    def forward_flat(self, img_nhwc, unet_nhwc, idx, B, n_rows, offsets=None, sink=None):
        offsets = idx.pt_offs[:n_rows] if offsets is None else offsets
        return self._forward_tensor_core(img_nhwc, unet_nhwc, idx, B, n_rows, offsets, 0, sink)


def _index_from_infos(infos, H, W) -> ops.PillarIndex:
    """Flat index over the pc0 frames only, rebuilt from per-sample voxel info dicts (used by the
    reference-signature ``forward``; DeFlow.forward passes its PillarIndex straight through)."""
    dev = infos[0]["voxel_coords"].device
    B = len(infos)
    keys, n_pts = [], []
    for b, e in enumerate(infos):
        c = e["voxel_coords"].long()
        keys.append(b * H * W + c[:, 1] * W + c[:, 2])
        n_pts.append(c.shape[0])
    key = torch.cat(keys)
    pix, inv, cnt = torch.unique(key, sorted=True, return_inverse=True, return_counts=True)
    order = torch.argsort(inv, stable=True)
    n_pil = [int(((pix >= b * H * W) & (pix < (b + 1) * H * W)).sum()) for b in range(B)]
    idx = ops.PillarIndex()
    F_ = B
    pt_off = [0]
    pil_off = [0]
    for b in range(B):
        pt_off.append(pt_off[-1] + n_pts[b])
        pil_off.append(pil_off[-1] + n_pil[b])
    host = n_pts + n_pil + pt_off + pil_off
    idx.F, idx.cap, idx.pil_cap = F_, max(int(key.shape[0]), 1), max(int(pix.shape[0]), 1)
    idx._host = host
    idx.counts = torch.tensor(host, dtype=torch.int32, device=dev)
    idx.pt_pillar = inv.to(torch.int32).contiguous()
    idx.pil_pix = pix.to(torch.int32).contiguous()
    start = torch.zeros(pix.shape[0] + 1, dtype=torch.int32, device=dev)
    start[1:] = torch.cumsum(cnt, 0).to(torch.int32)
    idx.pil_start = start
    idx.sorted_pt = order.to(torch.int32).contiguous()
    idx.pt_offs = torch.cat([e["point_offsets"] for e in infos], 0)
    return idx
