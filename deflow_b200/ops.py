"""Host side of the hot-path operators: tensor allocation, stream plumbing and autograd wiring
around the C-ABI kernels (include/deflow_b200.h).  PyTorch is used for device memory, streams and
the autograd tape only; all arithmetic on these paths happens in deflow_b200/csrc/*.cu.

OSF = /root/reference/OpenSceneFlow.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import IndexArgs, PfnArgs, PfnBwdArgs, check


def _stream(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        # the reference extension raises the same way for CPU tensors
        # (OSF/assets/cuda/mmcv/pytorch_device_registry.hpp:116-122)
        raise RuntimeError(f"{what}: implementation for device {t.device.type} not found (deflow_b200 is CUDA sm_100a only)")


def _f3(vals: Sequence[float]):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def grid_size(voxel_size, pc_range):
    """(grid_x, grid_y, grid_z) = round((max-min)/voxel) in fp32 (OSF/assets/cuda/mmcv/voxelization_cuda.cu:269-271)."""
    out = (C.c_int * 3)()
    check(_lib.lib().dfb_grid_size(_f3(voxel_size), _f3(pc_range), out), "grid_size")
    return int(out[0]), int(out[1]), int(out[2])


# ----------------------------------------------------------------------------------------------
# Batched pillar index
# ----------------------------------------------------------------------------------------------
class PillarIndex:
    """Flat, frame-after-frame description of the valid points and occupied pillars of F frames.

    Everything a per-sample ``DynamicVoxelizer`` + ``unique_dim`` produces in the reference
    (OSF/src/models/basic/encoder.py:567-600, OSF/assets/cuda/mmcv/scatter_points_cuda.cu:24-37),
    for all frames at once.  ``counts`` (device int32) = n_valid[F] | n_pillars[F] | pt_off[F+1] |
    pil_off[F+1]; ``host_counts()`` fetches it once (the only host sync of a forward)."""

    def __init__(self):
        self._host = None

    def host_counts(self):
        if self._host is None:
            self._host = self.counts.cpu().tolist()
        return self._host

    def n_valid(self, f):
        return self.host_counts()[f]

    def n_pillars(self, f):
        return self.host_counts()[self.F + f]

    def pt_off(self, f):
        return self.host_counts()[2 * self.F + f]

    def pil_off(self, f):
        return self.host_counts()[3 * self.F + 1 + f]

    def frame_info(self, f) -> dict:
        """The per-sample dict of DynamicVoxelizer.forward (encoder.py:591-598)."""
        a, b = self.pt_off(f), self.pt_off(f + 1)
        return {"points": self.pt_xyz[a:b], "voxel_coords": self.pt_coor[a:b], "point_idxes": self.pt_idx[a:b],
                "point_offsets": self.pt_offs[a:b]}


def pillar_index(points: torch.Tensor, voxel_size, pc_range) -> PillarIndex:
    """points f32[F, Nmax, S>=3] NaN-padded (collate_fn_pad layout, OSF/src/dataset.py:22-74)."""
    _need_cuda(points, "pillar_index")
    assert points.dtype == torch.float32 and points.dim() == 3 and points.shape[2] >= 3
    points = points.contiguous()
    F, Nmax, S = points.shape
    dev = points.device
    lib = _lib.lib()
    vs, rg = _f3(voxel_size), _f3(pc_range)
    words, blocks = C.c_longlong(), C.c_longlong()
    check(lib.dfb_index_workspace(F, Nmax, vs, rg, C.byref(words), C.byref(blocks)), "index_workspace")
    gx, gy, gz = grid_size(voxel_size, pc_range)
    cap = max(F * Nmax, 1)
    pil_cap = max(min(cap, F * gx * gy * gz), 1)
    i32 = dict(dtype=torch.int32, device=dev)
    idx = PillarIndex()
    idx.F, idx.Nmax, idx.cap, idx.pil_cap = F, Nmax, cap, pil_cap
    idx.grid = (gx, gy, gz)
    idx.voxel_size, idx.pc_range = [float(v) for v in voxel_size], [float(v) for v in pc_range]
    idx.counts = torch.empty(4 * F + 2, **i32)
    idx.pt_xyz = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    idx.pt_coor = torch.empty((cap, 3), **i32)
    idx.pt_idx = torch.empty(cap, dtype=torch.int64, device=dev)
    idx.pt_offs = torch.empty((cap, 3), dtype=torch.float32, device=dev)
    idx.pt_pillar = torch.empty(cap, **i32)
    idx.pil_coor = torch.empty((pil_cap, 3), **i32)
    idx.pil_pix = torch.empty(pil_cap, **i32)
    idx.pil_start = torch.empty(pil_cap + 1, **i32)
    idx.sorted_pt = torch.empty(cap, **i32)
    idx.csr_rec = torch.empty((cap, 4), dtype=torch.float32, device=dev)
    nw, nb = max(F * words.value, 1), max(F * max(blocks.value, 1), 1)
    # everything that must start at zero lives in ONE allocation (one memset): occ bytes | pil_cnt | blk_cnt | tickets.
    # occ = occupancy byte map (32 bytes per bitmap word); the bitmap itself is written by the scan kernel.
    zero = torch.empty(8 * nw + pil_cap + nb + 4, **i32)
    occ = zero[:8 * nw]
    idx.pil_cnt, blk, tickets = zero[8 * nw:8 * nw + pil_cap], zero[8 * nw + pil_cap:8 * nw + pil_cap + nb], zero[-4:]
    bitmap = torch.empty(nw, **i32)
    word_rank = torch.empty(nw, **i32)
    scan_ws = torch.empty(max(int(lib.dfb_index_scan_workspace(F, words.value, pil_cap)), 1), **i32)
    slot = torch.empty(cap, **i32)
    a = IndexArgs()
    a.F, a.Nmax, a.pt_stride, a.pil_cap = F, Nmax, S, pil_cap
    a.voxel_size, a.range = vs, rg
    a.pts, a.keys, a.bitmap, a.word_rank = points.data_ptr(), None, bitmap.data_ptr(), word_rank.data_ptr()
    a.blk_cnt, a.pt_slot, a.counts = blk.data_ptr(), slot.data_ptr(), idx.counts.data_ptr()
    a.pt_xyz, a.pt_coor, a.pt_idx, a.pt_offs = (idx.pt_xyz.data_ptr(), idx.pt_coor.data_ptr(), idx.pt_idx.data_ptr(),
                                                idx.pt_offs.data_ptr())
    a.pt_pillar, a.pil_cnt, a.pil_coor, a.pil_pix = (idx.pt_pillar.data_ptr(), idx.pil_cnt.data_ptr(),
                                                     idx.pil_coor.data_ptr(), idx.pil_pix.data_ptr())
    a.pil_start, a.sorted_pt, a.csr_rec = idx.pil_start.data_ptr(), idx.sorted_pt.data_ptr(), idx.csr_rec.data_ptr()
    a.scan_ws, a.tickets = scan_ws.data_ptr(), tickets.data_ptr()
    a.zero_base, a.zero_bytes = zero.data_ptr(), zero.numel() * 4
    a.occ = occ.data_ptr()
    check(lib.dfb_pillar_index(C.byref(a), _stream(points)), "pillar_index")
    idx._keep = (points,)
    return idx


def ego_warp(pc0: torch.Tensor, pose0: Optional[torch.Tensor], pose1: Optional[torch.Tensor],
             ego: Optional[torch.Tensor], out: torch.Tensor):
    """cal_pose0to1 + warp (OSF/src/models/basic/__init__.py:4-15, OSF/src/models/deflow.py:60-77).
    pc0 f32[B,Nmax,3]; writes the warped clouds into out[b, :Nmax] (out: f32[>=B, Nmax'>=Nmax, 3])
    and returns (pose_flow f32[B,Nmax,3], pose_0to1 f32[B,4,4])."""
    _need_cuda(pc0, "ego_warp")
    pc0 = pc0.contiguous()
    B, Nmax, _ = pc0.shape
    assert out.is_contiguous() and out.shape[1] >= Nmax and out.shape[2] == 3 and out.dtype == torch.float32
    pose_flow = torch.empty_like(pc0)
    pose01 = torch.empty((B, 4, 4), dtype=torch.float32, device=pc0.device)
    p0 = pose0.contiguous().float() if pose0 is not None else None
    p1 = pose1.contiguous().float() if pose1 is not None else None
    eg = ego.contiguous().float() if ego is not None else None
    check(_lib.lib().dfb_ego_warp(_ptr(p0), _ptr(p1), _ptr(eg), pc0.data_ptr(), B, Nmax, out.data_ptr(),
                                  out.shape[1] * 3, pose_flow.data_ptr(), pose01.data_ptr(), _stream(pc0)), "ego_warp")
    return pose_flow, pose01


# ----------------------------------------------------------------------------------------------
# Fused pillar feature net
# ----------------------------------------------------------------------------------------------
def _pfn_args(idx: PillarIndex, H, W, training, center_off, eps, momentum, weight, gamma, beta, rm, rv, pil_mean,
              stats, bn_params, pil_feats, image, pt_mask, pil_hdr, partials=None, image_ready=None, phase=0,
              sync_stats=None, sync_counts=None) -> PfnArgs:
    a = PfnArgs()
    a.F, a.H, a.W, a.training = idx.F, H, W, int(training)
    a.voxel_size = _f3(idx.voxel_size)
    a.center_off = _f3(center_off)
    a.range_min = _f3(idx.pc_range[:3])
    a.eps, a.momentum = eps, momentum
    a.counts, a.pt_xyz, a.pt_coor, a.pt_pillar = (idx.counts.data_ptr(), idx.pt_xyz.data_ptr(), idx.pt_coor.data_ptr(),
                                                  idx.pt_pillar.data_ptr())
    a.pil_cnt, a.pil_coor, a.pil_pix = idx.pil_cnt.data_ptr(), idx.pil_coor.data_ptr(), idx.pil_pix.data_ptr()
    a.pil_start, a.sorted_pt = idx.pil_start.data_ptr(), idx.sorted_pt.data_ptr()
    a.weight, a.gamma, a.beta = weight.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    a.running_mean, a.running_var = _ptr(rm), _ptr(rv)
    a.pil_mean, a.stats, a.bn_params = pil_mean.data_ptr(), stats.data_ptr(), bn_params.data_ptr()
    a.pil_feats, a.image = _ptr(pil_feats), image.data_ptr()
    a.image_bf16 = int(image.dtype == torch.bfloat16)
    a.pil_cap = idx.pil_cap
    a.csr_rec, a.pt_mask, a.partials = idx.csr_rec.data_ptr(), pt_mask.data_ptr(), _ptr(partials)
    a.pil_hdr = pil_hdr.data_ptr()
    a.image_ready_event = None if image_ready is None else image_ready.cuda_event
    a.phase, a.sync_stats, a.sync_counts = phase, _ptr(sync_stats), _ptr(sync_counts)
    return a


class _PillarFeatureNet(torch.autograd.Function):
    """decorate(9) -> Linear(9,32) -> BatchNorm1d -> ReLU -> pillar mean -> NHWC pseudo-image
    (DynamicPillarFeatureNet.forward + PointPillarsScatter, OSF/src/models/basic/encoder.py:430-475, 126-147)."""

    @staticmethod
    def forward(ctx, weight, gamma, beta, idx: PillarIndex, running_mean, running_var, training, eps, momentum,
                center_off, image_dtype, want_feats, image, image_ready, sync=None):
        _need_cuda(weight, "pillar_feature_net")
        gx, gy, gz = idx.grid
        H, W = gy, gx
        dev = weight.device
        w, g, b = weight.detach().contiguous().float(), gamma.detach().contiguous().float(), beta.detach().contiguous().float()
        pil_mean = torch.empty((idx.pil_cap, 3), dtype=torch.float32, device=dev)
        stats = torch.empty((idx.F, 2, 32), dtype=torch.float64, device=dev)
        bn_params = torch.empty((idx.F, 4, 32), dtype=torch.float32, device=dev)
        pil_feats = torch.empty((idx.pil_cap, 32), dtype=torch.float32, device=dev) if want_feats else None
        pt_mask = torch.empty(idx.cap, dtype=torch.int32, device=dev)
        partials = torch.empty((idx.cap // 32 + 1, 2, 32), dtype=torch.float32, device=dev)
        pil_hdr = torch.empty((idx.pil_cap, 12), dtype=torch.float32, device=dev)
        ctx.set_materialize_grads(False)   # no zero-filled gradients for the non-differentiable pillar outputs
        if image is None:
            image = torch.empty((idx.F, H, W, 32), dtype=image_dtype, device=dev)
        else:
            assert tuple(image.shape) == (idx.F, H, W, 32) and image.dtype == image_dtype and image.is_contiguous()
            image = image.detach()   # a new tensor object on the same storage: the output is not the input itself
        sync = sync if (sync is not None and training) else None
        sync_counts = None
        if sync is None:
            a = _pfn_args(idx, H, W, training, center_off, eps, momentum, w, g, b, running_mean, running_var, pil_mean,
                          stats, bn_params, pil_feats, image, pt_mask, pil_hdr, partials, image_ready)
            check(_lib.lib().dfb_pfn_forward(C.byref(a), _stream(weight)), "pfn_forward")
        else:
            # SyncBatchNorm (OSF/train.py:128): the BatchNorm1d call of sample-frame f pools the points of frame f of
            # EVERY rank.  Phase 1 leaves this rank's feature moments in `stats`; they and the per-frame point counts
            # are summed over the ranks (two small collectives for all 2B frames); phase 2 normalises with the pooled ones.
            a = _pfn_args(idx, H, W, training, center_off, eps, momentum, w, g, b, running_mean, running_var, pil_mean,
                          stats, bn_params, pil_feats, image, pt_mask, pil_hdr, partials, image_ready, phase=1)
            check(_lib.lib().dfb_pfn_forward(C.byref(a), _stream(weight)), "pfn_forward (moments)")
            sync_stats = stats.clone()
            sync_counts = idx.counts[:idx.F].clone()
            sync.all_reduce_sum(sync_stats)
            sync.all_reduce_sum(sync_counts)
            a = _pfn_args(idx, H, W, training, center_off, eps, momentum, w, g, b, running_mean, running_var, pil_mean,
                          stats, bn_params, pil_feats, image, pt_mask, pil_hdr, partials, image_ready, phase=2,
                          sync_stats=sync_stats, sync_counts=sync_counts)
            check(_lib.lib().dfb_pfn_forward(C.byref(a), _stream(weight)), "pfn_forward (normalise)")
        ctx.idx, ctx.cfg = idx, (H, W, training, eps, momentum, center_off)
        ctx.sync, ctx.sync_counts = sync, sync_counts
        ctx.save_for_backward(w, g, b, pil_mean, stats, bn_params, pt_mask, pil_hdr)
        if pil_feats is None:
            pil_feats = pil_mean.new_empty(0)
        ctx.mark_non_differentiable(pil_feats, pil_mean)
        return image, pil_feats, pil_mean

    @staticmethod
    def backward(ctx, grad_image, _g1, _g2):
        w, g, b, pil_mean, stats, bn_params, pt_mask, pil_hdr = ctx.saved_tensors
        idx = ctx.idx
        H, W, training, eps, momentum, center_off = ctx.cfg
        if grad_image is None:
            return (None,) * 15
        grad_image = grad_image.contiguous()
        dev = w.device
        ba = PfnBwdArgs()
        # `image` is not touched by the backward kernels; pass the gradient buffer to carry the dtype flag
        ba.fwd = _pfn_args(idx, H, W, training, center_off, eps, momentum, w, g, b, None, None, pil_mean, stats,
                           bn_params, None, grad_image, pt_mask, pil_hdr)
        gw = torch.zeros_like(w)
        gg = torch.zeros_like(g)
        gb = torch.zeros_like(b)
        bwd_stats = torch.empty((idx.F, 32, 10), dtype=torch.float64, device=dev)
        ba.grad_image, ba.grad_weight, ba.grad_gamma, ba.grad_beta = (grad_image.data_ptr(), gw.data_ptr(),
                                                                     gg.data_ptr(), gb.data_ptr())
        ba.bwd_stats, ba.grad_accum = bwd_stats.data_ptr(), None
        if ctx.sync is None:
            check(_lib.lib().dfb_pfn_backward(C.byref(ba), _stream(w)), "pfn_backward")
        else:
            ba.phase = 1
            check(_lib.lib().dfb_pfn_backward(C.byref(ba), _stream(w)), "pfn_backward (point pass)")
            pooled = bwd_stats.clone()
            ctx.sync.all_reduce_sum(pooled)
            ba.phase, ba.sync_bwd_stats = 2, pooled.data_ptr()
            ba.fwd.sync_counts = ctx.sync_counts.data_ptr()
            check(_lib.lib().dfb_pfn_backward(C.byref(ba), _stream(w)), "pfn_backward (finalise)")
        return (gw, gg, gb) + (None,) * 12


def clear_rows(image: torch.Tensor, pix: torch.Tensor, counts: torch.Tensor, count_index: int):
    """Zero the rows image.view(-1, C)[pix[:counts[count_index]]] (the pillars a previous forward wrote into a reused
    pseudo-image) instead of zero-filling the whole canvas."""
    _need_cuda(image, "clear_rows")
    assert image.is_contiguous() and pix.dtype == torch.int32 and counts.dtype == torch.int32
    row_bytes = image.shape[-1] * image.element_size()
    check(_lib.lib().dfb_clear_rows(image.data_ptr(), row_bytes, pix.data_ptr(), counts.data_ptr(), count_index,
                                    pix.numel(), _stream(image)), "clear_rows")


def pillar_feature_net(weight, gamma, beta, idx, running_mean, running_var, training, eps, momentum, center_off,
                       image_dtype=torch.float32, want_feats=True, image=None, image_ready=None, sync=None):
    """image / image_ready: an image buffer the caller is zero-filling on another stream, and the torch.cuda.Event that
    marks the end of that fill (DynamicEmbedder.embed overlaps it with the index kernels); default: allocated and
    zero-filled here."""
    return _PillarFeatureNet.apply(weight, gamma, beta, idx, running_mean, running_var, training, eps, momentum,
                                   center_off, image_dtype, want_feats, image, image_ready, sync)


# ----------------------------------------------------------------------------------------------
# Decoder gather
# ----------------------------------------------------------------------------------------------
class _DecoderGather(torch.autograd.Function):
    """h0[p] = [img0[y,x], img1[y,x], unet[y,x]] for every pc0 point (OSF/src/models/basic/decoder.py:215-225)."""

    @staticmethod
    def forward(ctx, img, unet, idx: PillarIndex, B, n_rows, out_dtype, n_alloc, sink=None):
        _need_cuda(img, "decoder_gather")
        ctx.sink = sink
        assert img.is_contiguous() and unet.is_contiguous() and img.dtype == unet.dtype
        F, H, W, c = img.shape
        assert c == 32 and F == 2 * B and tuple(unet.shape) == (B, H, W, 64)
        if n_alloc is not None and n_alloc > n_rows:  # zero padding rows for the tensor-core decoder
            h0 = torch.empty((n_alloc, 128), dtype=out_dtype, device=img.device)
            h0[n_rows:].zero_()
        else:
            h0 = torch.empty((n_rows, 128), dtype=out_dtype, device=img.device)
        check(_lib.lib().dfb_decoder_gather(img.data_ptr(), unet.data_ptr(), int(img.dtype == torch.bfloat16), B, H, W,
                                            idx.counts.data_ptr(), idx.F, idx.pt_pillar.data_ptr(),
                                            idx.pil_pix.data_ptr(), h0.data_ptr(), int(out_dtype == torch.bfloat16),
                                            n_rows, _stream(img)), "decoder_gather")
        ctx.idx, ctx.meta = idx, (B, H, W, img.dtype)
        return h0

    @staticmethod
    def backward(ctx, grad_h0):
        idx = ctx.idx
        B, H, W, dt = ctx.meta
        grad_h0 = grad_h0.contiguous()
        assert grad_h0.dtype in (torch.float32, torch.bfloat16)
        g_unet = torch.empty((B, H, W, 64), dtype=dt, device=grad_h0.device)
        if ctx.sink is not None:
            # Deferred image part (deflow._SplitFrames.backward): only the UNet gradient is produced now; the rows that belong
            # to the pseudo-image are added into its gradient once the other consumers of the image have written theirs --
            # no dense zero-filled [2B,H,W,32] tensor, no dense addition.
            rows = torch.empty((idx.pil_cap, 64), dtype=torch.float32, device=grad_h0.device)
            # per-channel sums of g_unet (the bias gradient of the UNet's last convolution) from the pillar sums, attached the
            # way the data-gradient kernels attach theirs (conv.bias_grad): saves a pass over the dense 64-channel tensor
            colsum = torch.zeros(64, dtype=torch.float32, device=grad_h0.device)
            check(_lib.lib().dfb_decoder_gather_backward_rows(
                grad_h0.data_ptr(), int(grad_h0.dtype == torch.bfloat16), B, H, W, idx.counts.data_ptr(), idx.F,
                idx.pil_pix.data_ptr(), idx.pil_start.data_ptr(), idx.sorted_pt.data_ptr(), rows.data_ptr(), g_unet.data_ptr(),
                int(dt == torch.bfloat16), idx.pil_cap, colsum.data_ptr(), _stream(grad_h0)), "decoder_gather_backward_rows")
            g_unet._dfb_colsum = (colsum, g_unet._version)
            ctx.sink["gather"] = (rows, idx, B, H, W)
            return None, g_unet, None, None, None, None, None, None
        g_img = torch.empty((2 * B, H, W, 32), dtype=dt, device=grad_h0.device)
        check(_lib.lib().dfb_decoder_gather_backward(grad_h0.data_ptr(), int(grad_h0.dtype == torch.bfloat16), B, H, W,
                                                     idx.counts.data_ptr(), idx.F, idx.pil_pix.data_ptr(),
                                                     idx.pil_start.data_ptr(), idx.sorted_pt.data_ptr(),
                                                     g_img.data_ptr(), g_unet.data_ptr(), int(dt == torch.bfloat16),
                                                     idx.pil_cap, _stream(grad_h0)), "decoder_gather_backward")
        return g_img, g_unet, None, None, None, None, None, None


def gather_img_rows_add(rows, idx, B, H, W, g_img):
    """dfb_gather_img_rows_add: g_img[2B,H,W,32] += the deferred image rows of the decoder gather's backward."""
    check(_lib.lib().dfb_gather_img_rows_add(rows.data_ptr(), B, H, W, idx.counts.data_ptr(), idx.F, idx.pil_pix.data_ptr(),
                                             g_img.data_ptr(), int(g_img.dtype == torch.bfloat16), idx.pil_cap,
                                             _stream(g_img)), "gather_img_rows_add")


def add_cat2(a0, b0, a1, b1):
    """cat([a0 + b0, a1 + b1], 0) in one pass (b0 = b1 = None: plain concatenation)."""
    _need_cuda(a0, "add_cat2")
    for t in (a0, a1, b0, b1):
        assert t is None or (t.is_contiguous() and t.shape == a0.shape and t.dtype == a0.dtype)
    out = torch.empty((2 * a0.shape[0],) + tuple(a0.shape[1:]), dtype=a0.dtype, device=a0.device)
    check(_lib.lib().dfb_add_cat2(a0.data_ptr(), _ptr(b0), a1.data_ptr(), _ptr(b1), a0.numel() * a0.element_size(),
                                  int(a0.dtype == torch.bfloat16), out.data_ptr(), _stream(a0)), "add_cat2")
    return out


def decoder_gather(img, unet, idx, B, n_rows, out_dtype=torch.float32, n_alloc=None, sink=None):
    return _DecoderGather.apply(img, unet, idx, B, n_rows, out_dtype, n_alloc, sink)


# ----------------------------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------------------------
LOSS_KINDS = {"deflowLoss": 0, "ff3dLoss": 1, "zeroflowLoss": 2}


class _FlowLoss(torch.autograd.Function):
    """Sum over the samples of deflowLoss / ff3dLoss with gt = flow[idx] - pose_flow[idx]
    (OSF/src/lossfuncs.py:102-125, 148-157; OSF/src/trainer.py:120-142)."""

    @staticmethod
    def forward(ctx, est, flow_gt, pose_flow, classes, idx: PillarIndex, B, kind):
        _need_cuda(est, "flow_loss")
        est_c = est.detach().contiguous().float()
        flow_gt, pose_flow = flow_gt.contiguous(), pose_flow.contiguous()
        assert flow_gt.dtype == torch.float32 and flow_gt.shape == pose_flow.shape
        Nmax = flow_gt.shape[1]
        n = est_c.shape[0]
        ws = torch.empty(B * 8, dtype=torch.float64, device=est.device)
        loss = torch.empty(1, dtype=torch.float32, device=est.device)
        grad = torch.empty_like(est_c)
        cls = classes.contiguous() if classes is not None else None
        if cls is not None:
            assert cls.dtype == torch.uint8
        check(_lib.lib().dfb_flow_loss(kind, est_c.data_ptr(), flow_gt.data_ptr(), pose_flow.data_ptr(), _ptr(cls),
                                       idx.pt_idx.data_ptr(), idx.counts.data_ptr(), idx.F, B, Nmax, ws.data_ptr(),
                                       loss.data_ptr(), grad.data_ptr(), n, _stream(est)), "flow_loss")
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


def flow_loss(est_flat, flow_gt, pose_flow, classes, idx, B, loss_name="deflowLoss"):
    if loss_name not in LOSS_KINDS:
        raise RuntimeError(f"unknown loss function {loss_name}")
    return _FlowLoss.apply(est_flat, flow_gt, pose_flow, classes, idx, B, LOSS_KINDS[loss_name])


# ----------------------------------------------------------------------------------------------
# mmcv._ext drop-ins (generic N x C / arbitrary int32 coords)
# ----------------------------------------------------------------------------------------------
REDUCE = {"sum": 0, "mean": 1, "max": 2}


def dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3):
    """mmcv._ext.dynamic_voxelize_forward (OSF/assets/cuda/mmcv/voxelization.cpp:62-74): voxel_size /
    coors_range are CPU float tensors, coors is the caller's zero-initialised int32[N,3] output."""
    _need_cuda(points, "dynamic_voxelize_forward")
    if NDim != 3:
        raise RuntimeError("dynamic_voxelize_forward: NDim must be 3")
    assert points.dtype == torch.float32 and points.is_contiguous() and coors.dtype == torch.int32 and coors.is_contiguous()
    vs = _f3(torch.as_tensor(voxel_size, dtype=torch.float32).tolist())
    rg = _f3(torch.as_tensor(coors_range, dtype=torch.float32).tolist())
    check(_lib.lib().dfb_dynamic_voxelize_forward(points.data_ptr(), points.shape[0], points.shape[1], vs, rg,
                                                  coors.data_ptr(), _stream(points)), "dynamic_voxelize_forward")


def _scatter_index(coors, extent=None):
    """Unique voxels of int32 coordinate rows (any row with a negative component is invalid): dfb_scatter_index.
    -> point2voxel_map[n], voxel_coors[n,3] (first M rows valid, sorted), counts[n], pil_start[n+1], sorted_pt[n], counts6."""
    n = coors.shape[0]
    dev = coors.device
    i32 = dict(dtype=torch.int32, device=dev)
    if extent is None:
        # exclusive coordinate bound: one small reduction + host read, like the reference's own
        # .item() sync after unique_dim (scatter_points_cuda.cu:29)
        ext = (coors.max(dim=0).values + 1).clamp_(min=1).tolist()
    else:
        ext = [int(e) for e in extent]
    extent_c = (C.c_int * 3)(*ext)
    cells = ext[0] * ext[1] * ext[2]
    words = (cells + 31) // 32
    bitmap, word_rank = torch.empty(words, **i32), torch.empty(words, **i32)
    blk, slot = torch.empty((n + 1023) // 1024, **i32), torch.empty(n, **i32)
    cmap, vcoors, vcount = torch.empty(n, **i32), torch.empty((n, 3), **i32), torch.empty(n, **i32)
    pil_start, sorted_pt, counts6 = torch.empty(n + 1, **i32), torch.empty(n, **i32), torch.empty(6, **i32)
    check(_lib.lib().dfb_scatter_index(coors.data_ptr(), n, extent_c, bitmap.data_ptr(), word_rank.data_ptr(), blk.data_ptr(),
                                       slot.data_ptr(), cmap.data_ptr(), vcoors.data_ptr(), vcount.data_ptr(),
                                       pil_start.data_ptr(), sorted_pt.data_ptr(), counts6.data_ptr(), _stream(coors)),
          "scatter_index")
    return cmap, vcoors, vcount, pil_start, sorted_pt, counts6


def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
    """mmcv._ext.dynamic_point_to_voxel_forward (OSF/assets/cuda/mmcv/scatter_points_cuda.cu:9-66)
    -> [voxel_feats, voxel_coors, point2voxel_map, voxel_points_count]."""
    if reduce_type not in REDUCE:
        raise RuntimeError("do not support reduce type " + str(reduce_type))  # scatter_points.cpp:32
    _need_cuda(feats, "dynamic_point_to_voxel_forward")
    assert feats.dtype == torch.float32, "deflow_b200 scatter supports float32 features"
    feats, coors = feats.contiguous(), coors.contiguous()
    n, c = feats.shape
    dev = feats.device
    i32 = dict(dtype=torch.int32, device=dev)
    if n == 0:  # scatter_points_cuda.cu:15-18
        return [feats.clone(), coors.clone(), torch.empty(0, **i32), torch.empty(0, **i32)]
    cmap, vcoors, vcount, pil_start, sorted_pt, counts6 = _scatter_index(coors)
    lib = _lib.lib()
    st = _stream(feats)
    out = torch.empty((n, c), dtype=torch.float32, device=dev)
    check(lib.dfb_scatter_reduce(feats.data_ptr(), n, c, pil_start.data_ptr(), sorted_pt.data_ptr(),
                                 counts6[1:].data_ptr(), n, REDUCE[reduce_type], out.data_ptr(), st), "scatter_reduce")
    m = int(counts6[1].item())
    return [out[:m], vcoors[:m], cmap, vcount[:m]]


def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                    reduce_type):
    """mmcv._ext.dynamic_point_to_voxel_backward (OSF/assets/cuda/mmcv/scatter_points_cuda.cu:68-132)."""
    if reduce_type not in REDUCE:
        raise RuntimeError("do not support reduce type " + str(reduce_type))
    _need_cuda(grad_feats, "dynamic_point_to_voxel_backward")
    assert grad_feats.is_contiguous() and grad_feats.dtype == torch.float32
    grad_reduced_feats = grad_reduced_feats.contiguous()
    n, c = grad_feats.shape
    m = reduced_feats.shape[0]
    ws = None
    if reduce_type == "max":
        ws = torch.empty(max(m * c, 1), dtype=torch.int32, device=grad_feats.device)
    check(_lib.lib().dfb_dynamic_point_to_voxel_backward(
        grad_feats.data_ptr(), grad_reduced_feats.data_ptr(), feats.contiguous().data_ptr(),
        reduced_feats.contiguous().data_ptr(), coors_idx.contiguous().data_ptr(), reduce_count.contiguous().data_ptr(),
        n, m, c, REDUCE[reduce_type], _ptr(ws), _stream(grad_feats)), "dynamic_point_to_voxel_backward")


def hard_voxelize_forward(points, voxel_size, coors_range, voxels, coors, num_points_per_voxel, voxel_num, max_points,
                          max_voxels, NDim=3, deterministic=True):
    """mmcv._ext.hard_voxelize_forward (OSF/assets/cuda/mmcv/voxelization.cpp:36-60, voxelization_cuda.cu:8-148): fills the
    caller's pre-zeroed ``voxels [max_voxels,max_points,F]``, ``coors [max_voxels,3]``, ``num_points_per_voxel [max_voxels]``
    and sets ``voxel_num`` (a 0-d int64 tensor, voxelize.py:94).  Voxels are numbered in order of first appearance and keep
    their first ``max_points`` points -- the reference's deterministic result (``deterministic=False`` may return any valid
    assignment there; the deterministic one is returned here too)."""
    _need_cuda(points, "hard_voxelize_forward")
    if NDim != 3:
        raise RuntimeError("hard_voxelize_forward: NDim must be 3")
    points = points.contiguous()
    assert points.dtype == torch.float32 and voxels.is_contiguous() and coors.is_contiguous() and coors.dtype == torch.int32
    assert num_points_per_voxel.dtype == torch.int32 and num_points_per_voxel.is_contiguous()
    n, c = points.shape
    if n == 0:
        voxel_num.fill_(0)
        return
    dev = points.device
    i32 = dict(dtype=torch.int32, device=dev)
    vs_t = torch.as_tensor(voxel_size, dtype=torch.float32)
    rg_t = torch.as_tensor(coors_range, dtype=torch.float32)
    tmp = torch.zeros((n, 3), **i32)
    dynamic_voxelize_forward(points, vs_t, rg_t, tmp, 3)
    gx, gy, gz = grid_size(vs_t.tolist(), rg_t.tolist())
    cmap, _, _, pil_start, sorted_pt, _ = _scatter_index(tmp, (gz, gy, gx))
    lib = _lib.lib()
    ws = torch.empty(int(lib.dfb_hard_voxelize_workspace(n)), **i32)
    vnum = torch.zeros(1, **i32)
    check(lib.dfb_hard_voxelize_assign(points.data_ptr(), n, c, tmp.data_ptr(), cmap.data_ptr(), pil_start.data_ptr(),
                                       sorted_pt.data_ptr(), int(max_points), int(max_voxels), voxels.data_ptr(),
                                       coors.data_ptr(), num_points_per_voxel.data_ptr(), vnum.data_ptr(), ws.data_ptr(),
                                       _stream(points)), "hard_voxelize_assign")
    voxel_num.fill_(int(vnum.item()))      # the reference returns the count to the host as well (voxelization_cuda.cu:144-147)
