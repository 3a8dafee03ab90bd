"""Per-point flow decoders in bf16 perf mode: the gate / MLP matrices are 1x1 tcgen05 convolutions over the point
list (csrc/conv_igemm.cu), the gate arithmetic between them is csrc/gru_elem.cu, and the backward is written out
by hand (one autograd node for the whole decoder, intermediate activations saved in bf16).

Reference: ConvGRUDecoder.forward_single / LinearDecoder.forward_single (OSF/src/models/basic/decoder.py:210-237,
81-104) -- per-sample Python loops over cuDNN/cuBLAS calls there."""
from __future__ import annotations

import os

import torch

from . import _lib, conv as tc
from ._lib import check
from .ops import _stream

BF = torch.bfloat16


def _img(t):  # [n_pad, C] -> [1, n_pad/8, 8, C] view for the convolution kernels
    return t.view(1, t.shape[0] // 8, 8, t.shape[1])


def _s3(t):
    return t.dtype == torch.float32


def _conv(srcs, w4, b, cout):
    wf, _ = tc.packed_raw(w4, _s3(srcs[0]))
    return tc.conv2d_forward([_img(s) for s in srcs], wf, b, cout, 1, 1).view(-1, cout)


def _dgrad(gy, w4, lo, hi):
    _, wd = tc.packed_raw(w4, _s3(gy))
    n_pad = gy.shape[0]
    return tc.conv2d_dgrad(_img(gy), wd, n_pad // 8, 8, hi - lo, w4.shape[1], lo, 1, 1).view(n_pad, hi - lo)


def _wgrad(srcs, gy, acc=None, gbias=None):
    return tc.conv2d_wgrad([_img(s) for s in srcs], _img(gy), 1, 1, acc, grad_bias=gbias)


# The bias gradients of the gate / head matrices ride in the weight-gradient launches (idle accumulator rows of the
# 128 + 64-channel input, csrc/conv_igemm.cu k_conv_wgrad `ones_half`) instead of separate passes over dq / dzr / dy1
# (k_channel_sum_s, 0.57 ms per step at config 2).  DFB_WGRAD_BIAS=0 restores the separate passes (A/B).
_WGRAD_BIAS = os.environ.get("DFB_WGRAD_BIAS", "1") != "0"


class _Decoder(torch.autograd.Function):
    """h0 [n_pad,128] fp32 (gathered pillar vectors), offsets [n,3] -> flow [n,3]: GEMMs as 1x1 tensor-core
    convolutions over the point list + the gate kernels of csrc/gru_elem.cu.
    parity = False: bf16 gate tensors (the launch-per-stage predecessor of the fused kernels, kept for A/B runs and the
    linear decoder); parity = True: fp32 gate tensors and split-precision ("bf16x3") GEMMs -- the fp32 parity mode."""

    @staticmethod
    def forward(ctx, h0, offsets, n, iters, parity, w_off, b_off, wz, bz, wr, br, wq, bq, w1, b1, w2, b2):
        lib = _lib.lib()
        dev = h0.device
        n_pad = h0.shape[0]
        st = _stream(h0)
        cx = w_off.shape[0]
        DT = torch.float32 if parity else BF
        f32 = int(parity)
        f = lambda t: t.detach().float().contiguous()  # noqa: E731
        x = torch.empty((n_pad, cx), dtype=DT, device=dev)
        offsets = offsets.contiguous()
        check(lib.dfb_offset_encode(offsets.data_ptr(), f(w_off).data_ptr(), f(b_off).data_ptr(), n, n_pad, cx,
                                    x.data_ptr(), f32, st), "offset_encode")
        if parity:
            hb = h0          # the GEMM operand is the fp32 state itself (split into hi / lo inside the convolution)
        else:
            hb = torch.empty((n_pad, 128), dtype=BF, device=dev)
            check(lib.dfb_to_bf16_pad(h0.data_ptr(), n_pad, n_pad, 128, hb.data_ptr(), st), "to_bf16")
        saved = {"hs": [], "hbs": [], "zr": [], "q": [], "rh": []}
        h = h0
        wzr4 = bzr = wq4 = None
        if iters > 0:
            wzr4 = torch.cat([f(wz)[:, :, 0], f(wr)[:, :, 0]], 0).reshape(256, 192, 1, 1).contiguous()
            bzr = torch.cat([f(bz), f(br)])
            wq4 = f(wq).reshape(128, 192, 1, 1)
            bqf = f(bq)
        for _ in range(iters):
            zr = _conv([hb, x], wzr4, bzr, 256)
            rh = torch.empty((n_pad, 128), dtype=DT, device=dev)
            check(lib.dfb_gru_rh(zr.data_ptr(), h.data_ptr(), n, n_pad, rh.data_ptr(), f32, st), "gru_rh")
            q = _conv([rh, x], wq4, bqf, 128)
            h_new = torch.empty_like(h)
            hb_new = None if parity else torch.empty_like(hb)
            check(lib.dfb_gru_update(zr.data_ptr(), q.data_ptr(), h.data_ptr(), n, n_pad, h_new.data_ptr(),
                                     None if parity else hb_new.data_ptr(), f32, st), "gru_update")
            saved["hs"].append(h); saved["hbs"].append(hb); saved["zr"].append(zr); saved["q"].append(q); saved["rh"].append(rh)
            h, hb = h_new, (h_new if parity else hb_new)
        w14 = f(w1).reshape(32, w1.shape[1], 1, 1)
        y1 = _conv([hb, x], w14, f(b1), 32)
        flow = torch.empty((n, 3), dtype=torch.float32, device=dev)
        w2f = f(w2)
        check(lib.dfb_head_out(y1.data_ptr(), w2f.data_ptr(), f(b2).data_ptr(), n, flow.data_ptr(), f32, st), "head_out")
        ctx.saved = (saved, x, hb, y1, offsets, wzr4, wq4, w14, w2f)
        ctx.meta = (n, n_pad, iters, cx, parity)
        return flow

    @staticmethod
    def backward(ctx, dflow):
        lib = _lib.lib()
        saved, x, hb_last, y1, offsets, wzr4, wq4, w14, w2f = ctx.saved
        n, n_pad, iters, cx, parity = ctx.meta
        DT = torch.float32 if parity else BF
        f32 = int(parity)
        dev = dflow.device
        st = _stream(dflow)
        dflow = dflow.contiguous().float()
        z32 = lambda *s: tc.zeros(s, torch.float32, dev)  # noqa: E731
        dy1 = torch.empty((n_pad, 32), dtype=DT, device=dev)
        gw2, gb2 = z32(3, 32), z32(3)
        check(lib.dfb_head_out_backward(y1.data_ptr(), w2f.data_ptr(), dflow.data_ptr(), n, n_pad, dy1.data_ptr(),
                                        gw2.data_ptr(), gb2.data_ptr(), f32, st), "head_out_backward")
        gw1 = _wgrad([hb_last, x], dy1)
        gb1 = tc.channel_sum(dy1)
        dh = z32(n_pad, 128)
        dx = z32(n_pad, cx)
        d_h = _dgrad(dy1, w14, 0, 128)
        d_x = _dgrad(dy1, w14, 128, 128 + cx)
        check(lib.dfb_acc_bf16(dh.data_ptr(), d_h.data_ptr(), None, dh.numel(), f32, st), "acc")
        check(lib.dfb_acc_bf16(dx.data_ptr(), d_x.data_ptr(), None, dx.numel(), f32, st), "acc")
        gwzr = gwq = gbzr = gbq = None
        for t in reversed(range(iters)):
            h, hb, zr, q, rh = saved["hs"][t], saved["hbs"][t], saved["zr"][t], saved["q"][t], saved["rh"][t]
            dq = torch.empty((n_pad, 128), dtype=DT, device=dev)
            dzr = torch.empty((n_pad, 256), dtype=DT, device=dev)
            dh_acc = torch.empty_like(dh)
            check(lib.dfb_gru_bwd1(zr.data_ptr(), q.data_ptr(), h.data_ptr(), dh.data_ptr(), n, n_pad, dq.data_ptr(),
                                   dzr.data_ptr(), dh_acc.data_ptr(), f32, st), "gru_bwd1")
            gwq = _wgrad([rh, x], dq, gwq)
            s = tc.channel_sum(dq)
            gbq = s if gbq is None else gbq + s
            d_rh = _dgrad(dq, wq4, 0, 128)
            d_xq = _dgrad(dq, wq4, 128, 192)
            check(lib.dfb_gru_bwd2(zr.data_ptr(), h.data_ptr(), d_rh.data_ptr(), n, n_pad, dzr.data_ptr(),
                                   dh_acc.data_ptr(), f32, st), "gru_bwd2")
            gwzr = _wgrad([hb, x], dzr, gwzr)
            s = tc.channel_sum(dzr)
            gbzr = s if gbzr is None else gbzr + s
            d_h2 = _dgrad(dzr, wzr4, 0, 128)
            d_xzr = _dgrad(dzr, wzr4, 128, 192)
            check(lib.dfb_acc_bf16(dh_acc.data_ptr(), d_h2.data_ptr(), None, dh_acc.numel(), f32, st), "acc")
            check(lib.dfb_acc_bf16(dx.data_ptr(), d_xq.data_ptr(), d_xzr.data_ptr(), dx.numel(), f32, st), "acc")
            dh = dh_acc
        gw_off, gb_off = z32(cx, 3), z32(cx)
        check(lib.dfb_offset_encode_backward(dx.data_ptr(), offsets.data_ptr(), n, cx, gw_off.data_ptr(),
                                             gb_off.data_ptr(), st), "offset_encode_backward")
        if iters > 0:
            gz, gr = gwzr[:128, :, 0], gwzr[128:, :, 0]   # [128,192,1] each (Conv1d k=1 layout)
            gq = gwq[:, :, 0]
            gbz, gbr = gbzr[:128], gbzr[128:]
        else:
            gz = gr = gq = gbz = gbr = gbq = None
        return (dh, None, None, None, None, gw_off, gb_off, gz, gbz, gr, gbr, gq, gbq, gw1[:, :, 0, 0], gb1, gw2, gb2)


def decode(h0, offsets, n, head, iters, parity=False):
    """head: ConvGRUDecoder (iters > 0) or LinearDecoder (iters == 0)."""
    if iters > 0:
        g = head.gru
        args = (g.convz.weight, g.convz.bias, g.convr.weight, g.convr.bias, g.convq.weight, g.convq.bias)
    else:
        args = (None,) * 6
    return _Decoder.apply(h0, offsets, n, iters, parity, head.offset_encoder.weight, head.offset_encoder.bias, *args,
                          head.decoder[0].weight, head.decoder[0].bias, head.decoder[2].weight, head.decoder[2].bias)


class _GRUStep(torch.autograd.Function):
    """ONE ConvGRU iteration h, x -> h' (decoder.py:184-193) in fp32 parity arithmetic: the gate GEMMs are 1x1
    split-precision tensor-core convolutions over the row list, the gate math is csrc/gru_elem.cu.  Serves the
    reference-signature ``ConvGRU.forward``; the decoders run all iterations through _Decoder / the fused kernels."""

    @staticmethod
    def forward(ctx, h, x, wz, bz, wr, br, wq, bq):
        lib = _lib.lib()
        if not h.is_cuda:
            raise RuntimeError("deflow_b200.ConvGRU runs on CUDA (sm_100a) only; there is no CPU path")
        if h.shape[1] != 128 or x.shape[1] not in (32, 64, 128):
            raise RuntimeError("ConvGRU: the tensor-core path covers hidden_dim 128 and input_dim 32 / 64 / 128")
        n = h.shape[0]
        n_pad = max((n + 7) // 8 * 8, 8)
        dev, st = h.device, _stream(h)
        f = lambda t: t.detach().float().contiguous()  # noqa: E731
        hp = torch.zeros((n_pad, 128), dtype=torch.float32, device=dev)
        xp = torch.zeros((n_pad, x.shape[1]), dtype=torch.float32, device=dev)
        hp[:n].copy_(h.detach())
        xp[:n].copy_(x.detach())
        cin = 128 + x.shape[1]
        wzr4 = torch.cat([f(wz)[:, :, 0], f(wr)[:, :, 0]], 0).reshape(256, cin, 1, 1).contiguous()
        wq4 = f(wq).reshape(128, cin, 1, 1)
        zr = _conv([hp, xp], wzr4, torch.cat([f(bz), f(br)]), 256)
        rh = torch.empty_like(hp)
        check(lib.dfb_gru_rh(zr.data_ptr(), hp.data_ptr(), n, n_pad, rh.data_ptr(), 1, st), "gru_rh")
        q = _conv([rh, xp], wq4, f(bq), 128)
        h_new = torch.empty_like(hp)
        check(lib.dfb_gru_update(zr.data_ptr(), q.data_ptr(), hp.data_ptr(), n, n_pad, h_new.data_ptr(), None, 1, st),
              "gru_update")
        ctx.saved = (hp, xp, zr, rh, q, wzr4, wq4)
        ctx.meta = (n, n_pad, x.shape[1])
        return h_new[:n]

    @staticmethod
    def backward(ctx, dh_out):
        lib = _lib.lib()
        hp, xp, zr, rh, q, wzr4, wq4 = ctx.saved
        n, n_pad, cx = ctx.meta
        dev, st = dh_out.device, _stream(dh_out)
        dh = torch.zeros((n_pad, 128), dtype=torch.float32, device=dev)
        dh[:n].copy_(dh_out)
        dq = torch.empty((n_pad, 128), dtype=torch.float32, device=dev)
        dzr = torch.empty((n_pad, 256), dtype=torch.float32, device=dev)
        dh_acc = torch.empty_like(dh)
        check(lib.dfb_gru_bwd1(zr.data_ptr(), q.data_ptr(), hp.data_ptr(), dh.data_ptr(), n, n_pad, dq.data_ptr(),
                               dzr.data_ptr(), dh_acc.data_ptr(), 1, st), "gru_bwd1")
        gwq = _wgrad([rh, xp], dq)
        gbq = tc.channel_sum(dq)
        d_rh = _dgrad(dq, wq4, 0, 128)
        d_xq = _dgrad(dq, wq4, 128, 128 + cx)
        check(lib.dfb_gru_bwd2(zr.data_ptr(), hp.data_ptr(), d_rh.data_ptr(), n, n_pad, dzr.data_ptr(), dh_acc.data_ptr(),
                               1, st), "gru_bwd2")
        gwzr = _wgrad([hp, xp], dzr)
        gbzr = tc.channel_sum(dzr)
        d_h2 = _dgrad(dzr, wzr4, 0, 128)
        d_xzr = _dgrad(dzr, wzr4, 128, 128 + cx)
        check(lib.dfb_acc_bf16(dh_acc.data_ptr(), d_h2.data_ptr(), None, dh_acc.numel(), 1, st), "acc")
        dx = torch.zeros((n_pad, cx), dtype=torch.float32, device=dev)
        check(lib.dfb_acc_bf16(dx.data_ptr(), d_xq.data_ptr(), d_xzr.data_ptr(), dx.numel(), 1, st), "acc")
        return (dh_acc[:n], dx[:n], gwzr[:128, :, 0], gbzr[:128], gwzr[128:, :, 0], gbzr[128:], gwq[:, :, 0], gbq)


def gru_step(h, x, gru_mod):
    """h[N,128], x[N,Cx] fp32 -> h'[N,128] (one iteration of ``gru_mod``: a decoder.ConvGRU)."""
    return _GRUStep.apply(h.float(), x.float(), gru_mod.convz.weight, gru_mod.convz.bias, gru_mod.convr.weight,
                          gru_mod.convr.bias, gru_mod.convq.weight, gru_mod.convq.bias)


class _FusedGRUDecoder(torch.autograd.Function):
    """ConvGRUDecoder on the persistent fused kernels (csrc/gru_fused.cu): h0 bf16 [n_pad,128], offsets [n,3] -> flow."""

    @staticmethod
    def forward(ctx, h0, offsets, n, iters, w_off, b_off, wz, bz, wr, br, wq, bq, w1, b1, w2, b2):
        lib = _lib.lib()
        dev = h0.device
        n_pad = h0.shape[0]
        st = _stream(h0)
        assert h0.dtype == BF and h0.is_contiguous() and w_off.shape[0] == 64 and iters > 0
        f32 = lambda t: t.detach().float().contiguous()  # noqa: E731
        wzr_b = torch.cat([f32(wz)[:, :, 0], f32(wr)[:, :, 0]], 0).to(BF).contiguous()       # [256,192]
        wq_b = f32(wq)[:, :, 0].to(BF).contiguous()                                           # [128,192]
        w1f = f32(w1)
        w1_b = w1f.to(BF).contiguous()                                                        # [32,192]
        w2f = f32(w2)
        par = torch.cat([f32(bz), f32(br), f32(bq), f32(b1), w2f.reshape(-1), f32(b2), f32(w_off).reshape(-1), f32(b_off)])
        offsets = offsets.contiguous()
        train = any(ctx.needs_input_grad)  # (grad mode is off inside Function.forward; inference passes no-grad inputs)
        hsave = torch.empty((iters + 1, n_pad, 128), dtype=BF, device=dev) if train else None
        xsave = torch.empty((n_pad, 64), dtype=BF, device=dev) if train else None
        y1 = torch.empty((n_pad, 32), dtype=BF, device=dev) if train else None
        flow = torch.empty((n, 3), dtype=torch.float32, device=dev)
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with tc._timed("k_gru_fused_fwd", 602688.0 * n * iters / 4, h0):
            check(lib.dfb_gru_fused_forward(h0.data_ptr(), offsets.data_ptr(), wzr_b.data_ptr(), wq_b.data_ptr(),
                                            w1_b.data_ptr(), par.data_ptr(), n, n_pad, iters, P(hsave), P(xsave), P(y1),
                                            flow.data_ptr(), st), "gru_fused_forward")
        ctx.saved = (hsave, xsave, y1, offsets, wzr_b, wq_b, w1f, w2f, par)
        ctx.meta = (n, n_pad, iters)
        return flow

    @staticmethod
    def backward(ctx, dflow):
        lib = _lib.lib()
        hsave, xsave, y1, offsets, wzr_b, wq_b, w1f, w2f, par = ctx.saved
        n, n_pad, iters = ctx.meta
        dev = dflow.device
        st = _stream(dflow)
        dflow = dflow.contiguous().float()
        z32 = lambda *s: tc.zeros(s, torch.float32, dev)  # noqa: E731
        # MLP head
        dy1 = torch.empty((n_pad, 32), dtype=BF, device=dev)
        gw2, gb2 = z32(3, 32), z32(3)
        check(lib.dfb_head_out_backward(y1.data_ptr(), w2f.data_ptr(), dflow.data_ptr(), n, n_pad, dy1.data_ptr(),
                                        gw2.data_ptr(), gb2.data_ptr(), 0, st), "head_out_backward")
        w14 = w1f.reshape(32, 192, 1, 1)
        fold = _WGRAD_BIAS
        gb1 = z32(32) if fold else None
        gw1 = _wgrad([hsave[iters], xsave], dy1, None, gb1)
        if not fold:
            gb1 = tc.channel_sum(dy1)
        d_h = _dgrad(dy1, w14, 0, 128)
        d_x = _dgrad(dy1, w14, 128, 192)
        # GRU iterations
        rh = torch.empty((iters, n_pad, 128), dtype=BF, device=dev)
        dq = torch.empty((iters, n_pad, 128), dtype=BF, device=dev)
        dzr = torch.empty((iters, n_pad, 256), dtype=BF, device=dev)
        dh0 = torch.empty((n_pad, 128), dtype=BF, device=dev)
        dx = torch.empty((n_pad, 64), dtype=torch.float32, device=dev)
        # algorithmic FLOPs of the launch: the data-gradient GEMMs of the GRU iterations (147 456 FLOP / point / iteration,
        # SURVEY 8a row a10); the gate recomputation it also does is not counted
        with tc._timed("k_gru_fused_bwd", 147456.0 * n * iters, dflow):
            check(lib.dfb_gru_fused_backward(hsave.data_ptr(), xsave.data_ptr(), d_h.data_ptr(), d_x.data_ptr(),
                                             wzr_b.data_ptr(), wq_b.data_ptr(), par.data_ptr(), n, n_pad, iters,
                                             rh.data_ptr(), dq.data_ptr(), dzr.data_ptr(), dh0.data_ptr(), dx.data_ptr(), st),
                  "gru_fused_backward")
        gwq = gwzr = None
        gbq, gbzr = (z32(128), z32(256)) if fold else (None, None)
        for t in range(iters):
            gwq = _wgrad([rh[t], xsave], dq[t], gwq, gbq)
            gwzr = _wgrad([hsave[t], xsave], dzr[t], gwzr, gbzr)
        if not fold:
            gbq = tc.channel_sum(dq.view(-1, 128))
            gbzr = tc.channel_sum(dzr.view(-1, 256))
        gw_off, gb_off = z32(64, 3), z32(64)
        check(lib.dfb_offset_encode_backward(dx.data_ptr(), offsets.data_ptr(), n, 64, gw_off.data_ptr(),
                                             gb_off.data_ptr(), st), "offset_encode_backward")
        return (dh0, None, None, None, gw_off, gb_off, gwzr[:128, :, 0], gbzr[:128], gwzr[128:, :, 0], gbzr[128:],
                gwq[:, :, 0], gbq, gw1[:, :, 0, 0], gb1, gw2, gb2)


def decode_fused(h0_bf16, offsets, n, head):
    g = head.gru
    return _FusedGRUDecoder.apply(h0_bf16, offsets, n, head.num_iters, head.offset_encoder.weight,
                                  head.offset_encoder.bias, g.convz.weight, g.convz.bias, g.convr.weight, g.convr.bias,
                                  g.convq.weight, g.convq.bias, head.decoder[0].weight, head.decoder[0].bias,
                                  head.decoder[2].weight, head.decoder[2].bias)
