"""Host side of the tensor-core convolution kernels (deflow_b200/csrc/conv_igemm.cu): NHWC bf16 activations,
weights kept in the reference's torch layout [Cout, Cin, k, k] fp32 (state_dict contract) and re-packed to
bf16 GEMM operands on the device at the start of EVERY forward (WeightBank: one launch for all layers, no cache).  OSF = /root/reference/OpenSceneFlow (unet.py:49-68)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import ConvArgs, check
from .ops import _ptr, _stream


# Optional per-launch timing (bench.py): a list that receives (kernel name, algorithmic FLOPs, start event, end event)
TIMING = None


class _timed:
    def __init__(self, name, flops, ref):
        self.on = TIMING is not None
        if self.on:
            self.rec = (name, flops, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.rec[2].record(torch.cuda.current_stream(ref.device))
            self.dev = ref.device

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.on:
            self.rec[3].record(torch.cuda.current_stream(self.dev))
            TIMING.append(self.rec)


class _ZeroPool:
    """Small zero-initialised accumulators (BatchNorm statistics, bias / affine gradients) carved from one zero-filled
    chunk: a training step asks for ~200 of them, and one fill kernel per request is 200 tiny launches serialised
    between the convolutions.  A chunk is handed out once and never recycled (its views keep it alive)."""
    CHUNK = 1 << 20   # bytes

    def __init__(self):
        self.buf = {}

    def take(self, shape, dtype, device):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = ((n * torch.empty((), dtype=dtype).element_size() + 255) // 256) * 256
        if nbytes > self.CHUNK // 4:
            return torch.zeros(shape, dtype=dtype, device=device)
        key = (device.type, device.index)
        ent = self.buf.get(key)
        if ent is None or ent[1] + nbytes > self.CHUNK:
            ent = [torch.zeros(self.CHUNK, dtype=torch.uint8, device=device), 0]
            self.buf[key] = ent
        off = ent[1]
        ent[1] += nbytes
        return ent[0][off:off + n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(shape)


_zeros = _ZeroPool()


def zeros(shape, dtype, device):
    return _zeros.take(tuple(shape) if not isinstance(shape, int) else (shape,), dtype, torch.device(device))


def _halo(ksize, stride, cins):
    return ksize == 3 and stride == 1 and all(c % 64 == 0 for c in cins)


def _halo_name(n_out, out_rows, out_dtype):
    """Kernel a halo launch resolves to in dfb_conv2d (csrc/conv_igemm.cu): 64 output channels with bf16 output and an
    even row count go to the row-pair kernel unless DFB_HALO_PAIR=0."""
    import os
    if n_out == 64 and out_dtype == torch.bfloat16 and out_rows % 2 == 0 and os.environ.get("DFB_HALO_PAIR", "1") != "0":
        return "k_conv_igemm_halo_pair"
    return f"k_conv_igemm_halo<{n_out}>"


def _args(mode, n, H, W, ksize, stride, xs, cins, cout, w, bias=None, y=None, stats=None, cin_total=0, cin_off=0,
          xs_lo=None):
    a = ConvArgs()
    a.mode, a.n, a.H, a.W, a.ksize, a.stride = mode, n, H, W, ksize, stride
    a.n_src = len(xs)
    for i, (x, c) in enumerate(zip(xs, cins)):
        a.x[i] = x.data_ptr()
        a.cin[i] = c
    a.cin_total, a.cin_off, a.cout = cin_total, cin_off, cout
    a.w = _ptr(w)
    a.bias = _ptr(bias)
    a.y = _ptr(y)
    a.y_fp32 = int(y is not None and y.dtype == torch.float32)
    a.stats = _ptr(stats)
    a.split3 = int(xs_lo is not None)
    if xs_lo is not None:
        for i, x in enumerate(xs_lo):
            a.x_lo[i] = x.data_ptr()
    return a


def split(x: torch.Tensor):
    """fp32 tensor -> (hi, lo) bf16 pair with hi + lo = x to ~16 significant bits: the operands of the split-precision
    ("bf16x3") parity mode.  Cached on the tensor (an activation feeds up to three convolutions)."""
    hit = getattr(x, "_dfb_split", None)
    if hit is not None and hit[2] == x._version:
        return hit[0], hit[1]
    xc = x.detach().contiguous()
    hi = torch.empty(xc.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty_like(hi)
    check(_lib.lib().dfb_split_bf16x2(xc.data_ptr(), xc.numel(), hi.data_ptr(), lo.data_ptr(), _stream(x)), "split_bf16x2")
    x._dfb_split = (hi, lo, x._version)
    return hi, lo


def pack_weights(w: torch.Tensor, need_dgrad: bool = True, split3: bool = False):
    """[Cout,Cin,k,k] fp32 -> (w_fwd bf16 [Cout, k*k*Cin], w_dgrad bf16 [Cin, k*k*Cout]); with split3 every row holds
    its hi part followed by its lo part (rows twice as long)."""
    w = w.detach().contiguous().float()
    cout, cin, k, _ = w.shape
    m = 2 if split3 else 1
    wf = torch.empty((cout, m * k * k * cin), dtype=torch.bfloat16, device=w.device)
    wd = torch.empty((cin, m * k * k * cout), dtype=torch.bfloat16, device=w.device) if need_dgrad else None
    check(_lib.lib().dfb_conv_pack_weights(w.data_ptr(), cout, cin, k, int(split3), wf.data_ptr(), _ptr(wd), _stream(w)),
          "pack_weights")
    return wf, wd


def out_size(H, W, ksize, stride):
    p = ksize // 2
    return (H + 2 * p - ksize) // stride + 1, (W + 2 * p - ksize) // stride + 1


def conv2d_forward(xs: Sequence[torch.Tensor], w_fwd, bias, cout, ksize, stride, stats=None, out_dtype=torch.bfloat16):
    """xs: 1-2 NHWC bf16 tensors [n,H,W,c_i] (concatenated along channels) -> y [n,Ho,Wo,cout]."""
    n, H, W, _ = xs[0].shape
    s3 = xs[0].dtype == torch.float32   # parity mode: fp32 activations, (hi, lo) operand pairs, fp32 output
    for x in xs:
        assert x.is_contiguous() and x.dtype == xs[0].dtype and x.shape[:3] == xs[0].shape[:3]
    xs_lo = None
    if s3:
        pairs = [split(x) for x in xs]
        xs, xs_lo = [p[0] for p in pairs], [p[1] for p in pairs]
        out_dtype = torch.float32
    Ho, Wo = out_size(H, W, ksize, stride)
    y = torch.empty((n, Ho, Wo, cout), dtype=out_dtype, device=xs[0].device)
    cins = [x.shape[3] for x in xs]
    a = _args(0, n, H, W, ksize, stride, xs, cins, cout, w_fwd, bias, y, stats, xs_lo=xs_lo)
    kc = 64 if all(c % 64 == 0 for c in cins) else 32
    name = _halo_name(cout, Ho, out_dtype) if _halo(ksize, stride, cins) else f"k_conv_igemm<{cout},{kc}>"
    with _timed(name, 2.0 * n * Ho * Wo * cout * sum(cins) * ksize * ksize, y):
        check(_lib.lib().dfb_conv2d(C.byref(a), _stream(y)), "conv2d forward")
    return y


def conv2d_dgrad(gy: torch.Tensor, w_dgrad, H, W, cin, cin_total, cin_off, ksize, stride, out_dtype=torch.bfloat16,
                 colsum: bool = False):
    """gy [n,Ho,Wo,cout] bf16 -> grad of the input channel slice [cin_off, cin_off+cin): [n,H,W,cin].
    colsum: also accumulate the per-channel sum of the result in the epilogue and attach it as ``gx._dfb_colsum`` --
    it IS the bias gradient of the convolution that produced this input (saves a separate pass over gx)."""
    assert gy.is_contiguous()
    n, _, _, cout = gy.shape
    gy_lo = None
    if gy.dtype == torch.float32:
        gy, lo = split(gy)
        gy_lo, out_dtype = [lo], torch.float32
    gx = torch.empty((n, H, W, cin), dtype=out_dtype, device=gy.device)
    stats = zeros((2, cin), torch.float64, gy.device) if colsum else None
    a = _args(1, n, H, W, ksize, stride, [gy], [cin], cout, w_dgrad, None, gx, stats, cin_total, cin_off, xs_lo=gy_lo)
    a.stats_sum_only = int(colsum)
    kc = 64 if cout % 64 == 0 else 32
    name = _halo_name(cin, H, out_dtype) if _halo(ksize, stride, [cout]) else f"k_conv_igemm<{cin},{kc}>"
    with _timed(name, 2.0 * n * gy.shape[1] * gy.shape[2] * cout * cin * ksize * ksize, gy):
        check(_lib.lib().dfb_conv2d(C.byref(a), _stream(gy)), "conv2d dgrad")
    if colsum:
        gx._dfb_colsum = (stats[0], gx._version)
    return gx


def conv2d_dgrad_two(gy: torch.Tensor, w_dgrad, H, W, c0, c1, cin_total, ksize, colsum: bool = True):
    """1x1 stride-1 data gradient for BOTH sources of a concatenated input in one pass over gy:
    -> (gx0 [n,H,W,c0], gx1 [n,H,W,c1]) bf16; with ``colsum`` each carries its per-channel sum (``_dfb_colsum``)."""
    assert gy.is_contiguous() and gy.dtype == torch.bfloat16 and ksize == 1
    n, _, _, cout = gy.shape
    gx0 = torch.empty((n, H, W, c0), dtype=torch.bfloat16, device=gy.device)
    gx1 = torch.empty((n, H, W, c1), dtype=torch.bfloat16, device=gy.device)
    stats = zeros((2, c0 + c1), torch.float64, gy.device) if colsum else None
    a = _args(1, n, H, W, ksize, 1, [gy], [c0], cout, w_dgrad, None, gx0, stats, cin_total, 0)
    a.stats_sum_only = int(colsum)
    a.y2, a.cin2 = gx1.data_ptr(), c1
    kc = 64 if cout % 64 == 0 else 32
    with _timed(f"k_conv_igemm<{c0 + c1},{kc}>", 2.0 * n * H * W * cout * (c0 + c1), gy):
        check(_lib.lib().dfb_conv2d(C.byref(a), _stream(gy)), "conv2d dgrad (two outputs)")
    if colsum:
        gx0._dfb_colsum = (stats[0][:c0], gx0._version)
        gx1._dfb_colsum = (stats[0][c0:], gx1._version)
    return gx0, gx1


_BORDER_COLSUM = __import__("os").environ.get("DFB_BORDER_COLSUM", "1") != "0"   # A/B switch


def conv3x3_dgrad_colsum(gy: torch.Tensor, gy_total: torch.Tensor, w: torch.Tensor, cin_off: int, cin: int) -> torch.Tensor:
    """sum over pixels of the 3x3 / stride 1 data gradient for the input channels [cin_off, cin_off + cin), from gy's border
    rows / columns, its total per-channel sum and the fp32 weights -- the data gradient itself is not read."""
    n, H, W, cout = gy.shape
    wf = w.detach().float().contiguous()
    tot = gy_total.detach().float().contiguous()
    ws = torch.empty(8 * cout, dtype=torch.float32, device=gy.device)
    out = torch.empty(cin, dtype=torch.float32, device=gy.device)
    check(_lib.lib().dfb_conv3x3_dgrad_colsum(gy.data_ptr(), int(gy.dtype == torch.float32), n, H, W, cout, tot.data_ptr(),
                                              wf.data_ptr(), wf.shape[1], cin_off, cin, ws.data_ptr(), out.data_ptr(),
                                              _stream(gy)), "conv3x3_dgrad_colsum")
    return out


def bias_grad(gy: torch.Tensor) -> torch.Tensor:
    """sum over pixels of gy: taken from the producing data-gradient kernel's epilogue when available."""
    cs = getattr(gy, "_dfb_colsum", None)
    # valid only while gy is the very tensor the data-gradient kernel wrote: autograd accumulating a second consumer's
    # gradient into it in place bumps its version, and then the sums are recomputed
    if cs is not None and cs[1] == gy._version and cs[0].shape[0] == gy.shape[-1]:
        return cs[0].float()
    return channel_sum(gy)


def conv2d_wgrad(xs: Sequence[torch.Tensor], gy: torch.Tensor, ksize, stride, grad_w: Optional[torch.Tensor] = None,
                 wacc_slot: Optional[torch.Tensor] = None, grad_bias: Optional[torch.Tensor] = None):
    """-> grad_w fp32 in torch layout [cout, sum(c_i), k, k].  wacc_slot: a [k*k, cout, cin] fp32 accumulator of a
    WeightBank with deferred gradients -- the result is ADDED there and nothing is returned (the bank unpacks all of its
    weight gradients with one launch at the end of the backward pass).  grad_bias: fp32 [cout], += the per-channel sums
    of gy from the same launch (1x1 layers with an odd number of 64-channel input groups: the gate / head matrices of the
    point decoder, 128 + 64 inputs)."""
    if gy.dtype == torch.float32:
        assert grad_bias is None
        # split precision: x_hi*gy_hi + x_hi*gy_lo + x_lo*gy_hi accumulated into the same fp32 gradient
        pairs = [split(x) for x in xs]
        gh, gl = split(gy)
        hi, lo = [p[0] for p in pairs], [p[1] for p in pairs]
        grad_w = conv2d_wgrad(hi, gh, ksize, stride, grad_w, wacc_slot)
        grad_w = conv2d_wgrad(hi, gl, ksize, stride, grad_w, wacc_slot)
        return conv2d_wgrad(lo, gh, ksize, stride, grad_w, wacc_slot)
    n, H, W, _ = xs[0].shape
    cout = gy.shape[3]
    cins = [x.shape[3] for x in xs]
    ct = sum(cins)
    assert gy.is_contiguous() and gy.dtype == torch.bfloat16
    deferred = wacc_slot is not None
    if deferred:
        assert tuple(wacc_slot.shape) == (ksize * ksize, cout, ct) and wacc_slot.is_contiguous()
        wacc, acc, grad_w = wacc_slot, 2, None        # bit 1: keep the accumulator's contents; no unpack
    else:
        wacc = torch.empty((ksize * ksize, cout, ct), dtype=torch.float32, device=gy.device)
        acc = int(grad_w is not None)
        if grad_w is None:
            grad_w = torch.empty((cout, ct, ksize, ksize), dtype=torch.float32, device=gy.device)
    a = _args(0, n, H, W, ksize, stride, xs, cins, cout, None, None, gy, None)
    if grad_bias is not None:
        assert grad_bias.dtype == torch.float32 and grad_bias.numel() == cout and grad_bias.is_contiguous()
        a.grad_bias = grad_bias.data_ptr()
    ncol = max(cout, 64)
    name = f"k_conv_wgrad_halo<{ncol}>" if (ksize == 3 and stride == 1) else f"k_conv_wgrad<{ncol}>"
    if ksize == 3 and stride == 1 and cout == 64 and sum((c + 63) // 64 for c in cins) <= 8 \
            and __import__("os").environ.get("DFB_WGRAD_X", "1") != "0":
        name = "k_conv_wgrad_x"      # cross-shift kernel for 64 output channels (dfb_conv2d_wgrad)
    with _timed(name, 2.0 * n * gy.shape[1] * gy.shape[2] * cout * ct * ksize * ksize, gy):
        check(_lib.lib().dfb_conv2d_wgrad(C.byref(a), wacc.data_ptr(), _ptr(grad_w), int(acc), _stream(gy)), "conv2d wgrad")
    return grad_w


# ----------------------------------------------------------------------------------------------
# HBM-bound UNet passes (deflow_b200/csrc/unet_elem.cu)
# ----------------------------------------------------------------------------------------------
def bn2d_finalize(stats, count, training, eps, momentum, gamma, beta, running_mean, running_var):
    Cn = gamma.shape[0]
    bn = torch.empty((4, Cn), dtype=torch.float32, device=gamma.device)
    check(_lib.lib().dfb_bn2d_finalize(_ptr(stats), float(count), Cn, int(training), eps, momentum, gamma.data_ptr(),
                                       beta.data_ptr(), _ptr(running_mean), _ptr(running_var), bn.data_ptr(),
                                       _stream(gamma)), "bn2d_finalize")
    return bn


def bn_gelu_apply(x, bn):
    y = torch.empty_like(x)
    Cn = x.shape[-1]
    check(_lib.lib().dfb_bn_gelu_apply(x.data_ptr(), bn.data_ptr(), Cn, x.numel() // Cn, y.data_ptr(),
                                       int(x.dtype == torch.float32), _stream(x)), "bn_gelu_apply")
    return y


def bn_gelu_backward(x, gy, bn, training, g_gamma, g_beta, g_bias, sync=None):
    Cn = x.shape[-1]
    gx = torch.empty_like(x)
    red = torch.empty(2 * Cn, dtype=torch.float64, device=x.device)
    args = (x.data_ptr(), gy.data_ptr(), bn.data_ptr(), Cn, x.numel() // Cn, int(training), red.data_ptr(), gx.data_ptr(),
            g_gamma.data_ptr(), g_beta.data_ptr(), _ptr(g_bias), int(x.dtype == torch.float32))
    if sync is None or not training:
        check(_lib.lib().dfb_bn_gelu_backward(*args, _stream(x)), "bn_gelu_backward")
        return gx
    # SyncBatchNorm: this rank's two per-channel sums (and the affine gradients from them, which stay local), their sum
    # over the ranks, then the data gradient with the pooled sums and the pooled pixel count
    check(_lib.lib().dfb_bn_gelu_backward_phase(*args, 1, 0.0, _stream(x)), "bn_gelu_backward (reduce)")
    sync.all_reduce_sum(red)
    check(_lib.lib().dfb_bn_gelu_backward_phase(*args, 2, float(x.numel() // Cn) * sync.world, _stream(x)),
          "bn_gelu_backward (apply)")
    return gx


def channel_sum(g):
    Cn = g.shape[-1]
    out = zeros((Cn,), torch.float32, g.device)
    check(_lib.lib().dfb_channel_sum(g.data_ptr(), Cn, g.numel() // Cn, out.data_ptr(), None,
                                     int(g.dtype == torch.float32), _stream(g)), "channel_sum")
    return out


def upsample2x(x, backward=False):
    n, h, w, Cn = x.shape
    if backward:
        h, w = h // 2, w // 2
        out = torch.empty((n, h, w, Cn), dtype=x.dtype, device=x.device)
    else:
        out = torch.empty((n, 2 * h, 2 * w, Cn), dtype=x.dtype, device=x.device)
    check(_lib.lib().dfb_upsample2x(x.data_ptr(), n, h, w, Cn, out.data_ptr(), int(backward),
                                    int(x.dtype == torch.float32), _stream(x)), "upsample2x")
    return out


# ----------------------------------------------------------------------------------------------
# autograd wiring
# ----------------------------------------------------------------------------------------------
class WeightBank:
    """bf16 GEMM operands of every convolution weight of a module tree, re-packed from the fp32 masters by ONE kernel
    launch at the start of every forward (``refresh``).  Nothing is cached across forwards: an optimizer step is
    invisible to autograd's version counters (``torch.optim.Adam(fused=True)`` does not bump ``_version``, neither does
    ``param.data.copy_``), so any cache keyed on the parameter goes stale -- which froze the UNet at its step-1
    weights in round 1.  The packed buffers persist (same storage every step); the device descriptor table is rebuilt
    only when a parameter's storage moves (``.to(device)``, ``load_state_dict`` with assign)."""

    def __init__(self, weights: Sequence[torch.Tensor]):
        self.weights = list(weights)
        assert 0 < len(self.weights) <= 64
        self._key = None
        self._table = {}
        self._bufs = {}
        self._index = {id(w): i for i, w in enumerate(self.weights)}
        self.defer_grads, self.grad_targets = False, None
        self._grad_key, self._bw_active = None, False
        self._unpack_key, self._unpack_table = None, None

    def _build(self, split3: bool):
        dev = self.weights[0].device
        m = 2 if split3 else 1
        n_el = [w.numel() for w in self.weights]
        bufs = []
        descs = (_lib.PackDesc * len(self.weights))()
        first = 0
        for i, w in enumerate(self.weights):
            cout, cin, k, _ = w.shape
            assert w.dtype == torch.float32 and w.is_contiguous() and w.device == dev
            wf = torch.empty((cout, m * k * k * cin), dtype=torch.bfloat16, device=dev)
            wd = torch.empty((cin, m * k * k * cout), dtype=torch.bfloat16, device=dev)
            bufs.append((wf, wd))
            d = descs[i]
            d.w, d.w_fwd, d.w_dgrad, d.first = w.data_ptr(), wf.data_ptr(), wd.data_ptr(), first
            d.cout, d.cin, d.ksize = cout, cin, k
            first += n_el[i]
        host = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).clone()
        self._table[split3] = (host.to(dev), first)
        self._bufs[split3] = bufs

    # ------------------------------------------------------------------ deferred weight gradients
    # With ``defer_grads`` (set by trainer.TrainStep) the weight-gradient kernels of the bank's convolutions leave their
    # results in ONE persistent fp32 accumulator ([tap][cout][cin] per weight, zeroed once per backward pass): the two calls
    # of a shared encoder weight and the three products of the split-precision mode sum there in place, the autograd
    # Functions return no weight gradient, and a callback queued on the autograd engine unpacks ALL gradients to the torch
    # layout with one launch when the backward pass ends -- straight into ``grad_targets`` (the views of the trainer's
    # flat gradient buffer) when given.  Replaces 54 unpack launches, 54 memsets and the 64 additions autograd performs
    # for the shared encoder weights per step.  Off by default: callers that use torch.autograd.grad() or gradient hooks
    # on these weights need the per-launch gradients.
    def enable_deferred_grads(self, grad_targets=None):
        self.defer_grads = True
        self.grad_targets = grad_targets      # {id(weight): fp32 tensor [cout,cin,k,k]} or None
        self._grad_key = None

    def _build_grad(self):
        dev = self.weights[0].device
        total = sum(w.numel() for w in self.weights)
        self._acc = torch.zeros(total, dtype=torch.float32, device=dev)
        self._acc_views, off = [], 0
        for w in self.weights:
            cout, cin, k, _ = w.shape
            self._acc_views.append(self._acc[off:off + w.numel()].view(k * k, cout, cin))
            off += w.numel()
        self._grad_total = total

    def wgrad_slot(self, w):
        """The accumulator of weight ``w`` for this backward pass (None: not a deferred weight).  The first request of a
        backward pass zeroes the accumulator and queues the unpack on the autograd engine."""
        if not getattr(self, "defer_grads", False):
            return None
        i = self._index.get(id(w))
        if i is None:
            return None
        key = tuple(x.data_ptr() for x in self.weights)
        if self._grad_key != key:
            self._build_grad()
            self._grad_key, self._unpack_key = key, None
        if not self._bw_active:
            self._acc.zero_()
            self._bw_active = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finish_backward)
        return self._acc_views[i]

    def _finish_backward(self):
        self._bw_active = False
        dev = self.weights[0].device
        plan, flat = [], None
        for w in self.weights:
            tgt = self.grad_targets.get(id(w)) if self.grad_targets else None
            accumulate = 0
            if tgt is None:
                if w.grad is not None:
                    tgt, accumulate = w.grad, 1
                else:
                    if flat is None:
                        flat = torch.empty(self._grad_total, dtype=torch.float32, device=dev)
                        off = 0
                    tgt = None      # filled below from `flat`
            elif w.grad is not None and w.grad.data_ptr() == tgt.data_ptr():
                accumulate = 1                      # a second backward into the same buffer (gradient accumulation)
            plan.append([w, tgt, accumulate])
        first = 0
        for ent in plan:
            if ent[1] is None:
                ent[1] = flat[first:first + ent[0].numel()].view_as(ent[0])
            first += ent[0].numel()
        # the device descriptor table is rebuilt (one small H2D copy) only when a target pointer or flag changes: in a
        # training loop it is the same every step, and nothing here synchronises the host with the device
        key = tuple((t.data_ptr(), a) for _, t, a in plan)
        if key != self._unpack_key:
            descs = (_lib.UnpackDesc * len(plan))()
            off = 0
            for i, (w, tgt, accumulate) in enumerate(plan):
                cout, cin, k, _ = w.shape
                d = descs[i]
                d.wacc, d.grad, d.first = self._acc_views[i].data_ptr(), tgt.data_ptr(), off
                d.cout, d.cin, d.ksize, d.accumulate = cout, cin, k, accumulate
                off += w.numel()
            self._unpack_table = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).clone().to(dev)
            self._unpack_key = key
        check(_lib.lib().dfb_wgrad_unpack_multi(self._unpack_table.data_ptr(), len(plan), first, _stream(self.weights[0])),
              "wgrad_unpack_multi")
        for w, tgt, _ in plan:
            if w.grad is None or w.grad.data_ptr() != tgt.data_ptr():
                w.grad = tgt

    def refresh(self, split3: bool):
        """Pack every weight now (one launch) and hang (token, w_fwd, w_dgrad) on its parameter for ``packed``;
        returns the token of this refresh."""
        key = tuple((w.data_ptr(), w.device) for w in self.weights)
        if key != self._key:
            self._key, self._table, self._bufs = key, {}, {}
        if split3 not in self._table:
            self._build(split3)
        table, total = self._table[split3]
        w0 = self.weights[0]
        check(_lib.lib().dfb_conv_pack_weights_multi(table.data_ptr(), len(self.weights), total, int(split3), _stream(w0)),
              "pack_weights_multi")
        WeightBank._tokens += 1
        token = WeightBank._tokens
        name = "_dfb_packed3" if split3 else "_dfb_packed"
        for w, pair in zip(self.weights, self._bufs[split3]):
            setattr(w, name, (token, pair[0], pair[1]))
            w._dfb_bank = self
        return token

    _tokens = 0


# tokens of the refreshes whose forward is running: packed() serves only operands packed by one of them
_live_tokens: List[int] = []


class bank_scope:
    """``with bank_scope(bank, split3):`` -- refresh the bank and let ``packed`` serve its operands inside the block."""

    def __init__(self, bank: WeightBank, split3: bool):
        self.bank, self.split3 = bank, split3

    def __enter__(self):
        _live_tokens.append(self.bank.refresh(self.split3))

    def __exit__(self, *a):
        _live_tokens.pop()


def packed(w: torch.Tensor, split3: bool = False):
    """bf16 GEMM operands of a convolution weight.  Inside a ``bank_scope`` they are the ones the bank packed at the
    start of THIS forward; anywhere else the weight is packed on the spot (never served from a cache: see WeightBank)."""
    hit = getattr(w, "_dfb_packed3" if split3 else "_dfb_packed", None)
    if hit is not None and hit[0] in _live_tokens:
        return hit[1], hit[2]
    return pack_weights(w, True, split3)


def packed_raw(w4: torch.Tensor, split3: bool = False):
    """Packed operands of a derived (non-parameter) weight tensor built inside one forward: cached on that tensor
    object, which does not outlive the autograd graph of the call."""
    name = "_dfb_packed3" if split3 else "_dfb_packed"
    hit = getattr(w4, name, None)
    if hit is None:
        hit = pack_weights(w4, True, split3)
        setattr(w4, name, hit)
    return hit


def _wgrad_slot(w):
    """Deferred-gradient accumulator of a bank weight (None: return the gradient through autograd as usual)."""
    bank = getattr(w, "_dfb_bank", None)
    return bank.wgrad_slot(w) if bank is not None else None


class _Conv(torch.autograd.Function):
    """Conv2d(k, stride 1, pad k//2, bias) over 1-2 channel-concatenated NHWC bf16 sources."""

    @staticmethod
    def forward(ctx, w, b, x0, x1):
        xs = [x0] if x1 is None else [x0, x1]
        k = w.shape[2]
        wf, wd = packed(w, x0.dtype == torch.float32)
        y = conv2d_forward(xs, wf, b.detach().float().contiguous(), w.shape[0], k, 1)
        ctx.save_for_backward(wd, w, *xs)
        ctx.k, ctx.wshape = k, tuple(w.shape)
        ctx.need = (ctx.needs_input_grad[2], x1 is not None and ctx.needs_input_grad[3])
        return y

    @staticmethod
    def backward(ctx, gy):
        wd, w, *xs = ctx.saved_tensors
        gb = bias_grad(gy)       # before .contiguous(): the attribute lives on the tensor the producer returned
        gy = gy.contiguous()
        k = ctx.k
        ct = ctx.wshape[1]
        slot = _wgrad_slot(w)
        gw = conv2d_wgrad(xs, gy, k, 1, None, slot)
        gxs, off = [None, None], 0
        # Per-channel sums of the data gradient = the bias gradient of the convolution that produced this input.  3x3: from
        # the epilogue of the (MMA-bound) halo kernel, where they are nearly free.  1x1: the epilogue is what bounds these
        # HBM-bound launches and the column sums doubled it (B200: 0.077 -> 0.140 ms for dec2.u1), while for a 1x1 stride-1
        # convolution they follow exactly from sum_p gx[p, ci] = sum_co W[co, ci] * sum_p gy[p, co] = (W^T gb)[ci].
        mv = (w.detach().reshape(w.shape[0], ct).float().t() @ gb.float()) if k == 1 else None
        # 3x3 with 64 input channels per source (the row-pair kernel, whose epilogue is its limit: +50 % with the sums):
        # from the border rows / columns of gy and W instead (dfb_conv3x3_dgrad_colsum); wider layers keep the epilogue sums
        borders = k == 3 and _BORDER_COLSUM and all(x.shape[3] == 64 for x in xs)
        if (len(xs) == 2 and k == 1 and all(ctx.need) and gy.dtype == torch.bfloat16
                and xs[0].shape[3] % 32 == 0 and xs[1].shape[3] % 32 == 0 and ct in (32, 64, 128, 256)):
            c0 = xs[0].shape[3]
            g0, g1 = conv2d_dgrad_two(gy, wd, xs[0].shape[1], xs[0].shape[2], c0, xs[1].shape[3], ct, k, colsum=False)
            g0._dfb_colsum, g1._dfb_colsum = (mv[:c0], g0._version), (mv[c0:], g1._version)
            return gw, gb, g0, g1
        for i, x in enumerate(xs):
            c = x.shape[3]
            if ctx.need[i]:
                gxs[i] = conv2d_dgrad(gy, wd, x.shape[1], x.shape[2], c, ct, off, k, 1, colsum=(k != 1 and not borders))
                if k == 1:
                    gxs[i]._dfb_colsum = (mv[off:off + c], gxs[i]._version)
                elif borders:
                    gxs[i]._dfb_colsum = (conv3x3_dgrad_colsum(gy, gb, w, off, c), gxs[i]._version)
            off += c
        return gw, gb, gxs[0], gxs[1]


def conv_bias(w, b, x0, x1=None):
    return _Conv.apply(w, b, x0, x1)


class _ConvBnGelu(torch.autograd.Function):
    """ConvWithNorms (OSF/src/models/basic/__init__.py:61-79): Conv2d(3, stride, 1) -> BatchNorm2d -> GELU."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, running_mean, running_var, stride, training, eps, momentum, sync=None):
        wf, wd = packed(w, x.dtype == torch.float32)
        cout = w.shape[0]
        sync = sync if training else None
        stats = (torch.zeros((2, cout), dtype=torch.float64, device=x.device) if sync is not None
                 else zeros((2, cout), torch.float64, x.device)) if training else None
        raw = conv2d_forward([x], wf, b.detach().float().contiguous(), cout, 3, stride, stats)
        count = raw.numel() // cout
        if sync is not None:
            # SyncBatchNorm (OSF/train.py:128): per-channel sum / sum of squares from the convolution epilogue, summed over
            # the ranks (one small collective per layer call); every rank runs the same B x H x W, so the pooled count is
            # world x count
            sync.all_reduce_sum(stats)
            count = count * sync.world
        ctx.sync = sync
        bn = bn2d_finalize(stats, count, training, eps, momentum, gamma.detach().float(), beta.detach().float(),
                           running_mean, running_var)
        act = bn_gelu_apply(raw, bn)
        ctx.save_for_backward(x, raw, bn, wd, w)
        ctx.cfg = (stride, training, tuple(w.shape))
        return act

    @staticmethod
    def backward(ctx, gact):
        x, raw, bn, wd, w = ctx.saved_tensors
        stride, training, wshape = ctx.cfg
        cout, cin = wshape[0], wshape[1]
        gg, gbeta, gbias = zeros((3, cout), torch.float32, x.device).unbind(0)
        graw = bn_gelu_backward(raw, gact.contiguous(), bn, training, gg, gbeta, gbias, ctx.sync)
        gw = conv2d_wgrad([x], graw, 3, stride, None, _wgrad_slot(w))
        gx = conv2d_dgrad(graw, wd, x.shape[1], x.shape[2], cin, cin, 0, 3, stride) if ctx.needs_input_grad[0] else None
        return gx, gw, gbias, gg, gbeta, None, None, None, None, None, None, None


def conv_bn_gelu(x, conv_mod, bn_mod, training):
    out = _ConvBnGelu.apply(x, conv_mod.weight, conv_mod.bias, bn_mod.weight, bn_mod.bias, bn_mod.running_mean,
                            bn_mod.running_var, conv_mod.stride[0], training, bn_mod.eps, bn_mod.momentum,
                            getattr(bn_mod, "dfb_sync", None))
    if training:
        _pending_nbt.append(bn_mod.num_batches_tracked)
    return out


# BatchNorm2d.num_batches_tracked += 1 for every call, applied in one foreach launch per forward (flush_batch_counters)
_pending_nbt = []


def flush_batch_counters():
    if _pending_nbt:
        uniq, counts = {}, {}
        for t in _pending_nbt:
            uniq[id(t)] = t
            counts[id(t)] = counts.get(id(t), 0) + 1
        by_count = {}
        for k, t in uniq.items():
            by_count.setdefault(counts[k], []).append(t)
        for c, ts in by_count.items():
            torch._foreach_add_(ts, c)
        _pending_nbt.clear()


class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return upsample2x(x)

    @staticmethod
    def backward(ctx, g):
        cs = getattr(g, "_dfb_colsum", None)
        out = upsample2x(g.contiguous(), backward=True)
        # every output pixel of the bilinear x2 spreads weights that sum to 1 (edge replication included), so the adjoint
        # preserves per-channel sums exactly: the bias-gradient sums ride through to the 1x1 convolution in front
        if cs is not None and cs[1] == g._version:
            out._dfb_colsum = (cs[0], out._version)
        return out


def upsample_bilinear2x(x):
    return _Upsample2x.apply(x)
