"""Library comparators for the GPU tests: the UNet and the per-point decoders computed by torch's own strict-fp32 kernels
(cuDNN / cuBLAS through torch.nn.functional) from the SAME parameters as a deflow_b200 module.  Test infrastructure
only -- the product modules have no library backend (north_star: no multi-backend dispatch).

Arithmetic follows OSF/src/models/basic/unet.py:70-100, basic/__init__.py:66-79 and basic/decoder.py:184-237."""
import torch
import torch.nn.functional as F


def _cwn(layer, x):
    bn = layer.batchnorm
    y = F.conv2d(x, layer.conv.weight, layer.conv.bias, layer.conv.stride, layer.conv.padding)
    y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, layer.training, bn.momentum, bn.eps)
    if layer.training:
        bn.num_batches_tracked += 1
    return F.gelu(y)


def _seq(step, x):
    for layer in step:
        x = _cwn(layer, x)
    return x


def _up(block, a, b):
    c1 = block.u1_u2[0]
    u2 = F.interpolate(F.conv2d(a, c1.weight, c1.bias), scale_factor=2, mode="bilinear", align_corners=False)
    u3 = F.conv2d(b, block.u3.weight, block.u3.bias)
    u4 = F.conv2d(torch.cat([u2, u3], 1), block.u4_u5[0].weight, block.u4_u5[0].bias, padding=1)
    return F.conv2d(u4, block.u4_u5[1].weight, block.u4_u5[1].bias, padding=1)


def unet_forward(net, pc0_B, pc1_B):
    """deflow_b200.FastFlow3DUNet parameters, NCHW fp32 inputs -> NCHW output, cuDNN strict fp32."""
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        f0 = _seq(net.encoder_step_1, pc0_B); l0 = _seq(net.encoder_step_2, f0); r0 = _seq(net.encoder_step_3, l0)
        f1 = _seq(net.encoder_step_1, pc1_B); l1 = _seq(net.encoder_step_2, f1); r1 = _seq(net.encoder_step_3, l1)
        s = _up(net.decoder_step1, torch.cat([r0, r1], 1), torch.cat([l0, l1], 1))
        t = _up(net.decoder_step2, s, torch.cat([f0, f1], 1))
        u = _up(net.decoder_step3, t, torch.cat([pc0_B, pc1_B], 1))
        return F.conv2d(u, net.decoder_step4.weight, net.decoder_step4.bias, padding=1)


def gru_step(g, h, x):
    """decoder.py:184-193 on row matrices h[N,128], x[N,64]."""
    hx = torch.cat([h, x], 1)
    z = torch.sigmoid(F.linear(hx, g.convz.weight[:, :, 0], g.convz.bias))
    r = torch.sigmoid(F.linear(hx, g.convr.weight[:, :, 0], g.convr.bias))
    q = torch.tanh(F.linear(torch.cat([r * h, x], 1), g.convq.weight[:, :, 0], g.convq.bias))
    return (1 - z) * h + z * q


def decoder_forward(head, h0, offsets, rnd=None):
    """ConvGRUDecoder / LinearDecoder head on gathered rows h0[N,128] (decoder.py:226-237, 95-104), cuBLAS strict fp32.
    rnd: optional rounding applied to every tensor the fused bf16 kernel rounds to bf16 before it feeds a GEMM (the hidden
    state, x, r*h) -- with rnd = bf16 round-trip (straight-through gradient) this is the fused kernel's arithmetic with
    exact transcendental functions and fp32 everything else."""
    rnd = rnd or (lambda t: t)
    x = rnd(head.offset_encoder(offsets))
    h = h0
    if hasattr(head, "gru"):
        g = head.gru
        for _ in range(head.num_iters):
            hb = rnd(h)
            hx = torch.cat([hb, x], 1)
            z = torch.sigmoid(F.linear(hx, g.convz.weight[:, :, 0], g.convz.bias))
            r = torch.sigmoid(F.linear(hx, g.convr.weight[:, :, 0], g.convr.bias))
            q = torch.tanh(F.linear(torch.cat([rnd(r * h), x], 1), g.convq.weight[:, :, 0], g.convq.bias))
            h = (1 - z) * h + z * q
    y1 = F.gelu(F.linear(torch.cat([rnd(h), x], 1), head.decoder[0].weight, head.decoder[0].bias))
    return F.linear(y1, head.decoder[2].weight, head.decoder[2].bias)


class _RoundBF16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return g


def round_bf16(t):
    """bf16 round trip with a straight-through gradient."""
    return _RoundBF16.apply(t)
