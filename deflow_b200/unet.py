"""FastFlow3D UNet backbone with the reference's parameter tree (OSF/src/models/basic/unet.py,
ConvWithNorms in OSF/src/models/basic/__init__.py:61-79), computing on NHWC tensors.

state_dict keys are the reference's: backbone.encoder_step_{1,2,3}.{i}.{conv,batchnorm}.*,
backbone.decoder_step{1,2,3}.{u1_u2.0,u3,u4_u5.0,u4_u5.1}.*, backbone.decoder_step4.*.
"""
from __future__ import annotations

import torch
from torch import nn


def _to_nhwc(x: torch.Tensor, dtype) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("deflow_b200 modules run on CUDA (sm_100a) only; there is no CPU path")
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


class ConvWithNorms(nn.Module):
    """Conv2d(k, stride, pad) -> BatchNorm2d -> GELU (exact erf)  (basic/__init__.py:61-79), NCHW in / NCHW out like the
    reference; computed by the tcgen05 convolution + BN/GELU kernels (deflow_b200/conv.py) on an NHWC copy.  fp32 input
    -> split-precision ("bf16x3") parity arithmetic, bf16 input -> bf16 operands."""

    def __init__(self, in_num_channels: int, out_num_channels: int, kernel_size: int, stride: int, padding: int):
        super().__init__()
        self.conv = nn.Conv2d(in_num_channels, out_num_channels, kernel_size, stride, padding)
        self.batchnorm = nn.BatchNorm2d(out_num_channels)
        self.nonlinearity = nn.GELU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import conv as tc
        if self.conv.kernel_size != (3, 3) or self.conv.padding != (1, 1):
            raise RuntimeError("ConvWithNorms: the tensor-core path covers the 3x3 / pad 1 layers of FastFlow3DUNet")
        if (x.shape[2] - 1) // self.conv.stride[0] + 1 == 1 and (x.shape[3] - 1) // self.conv.stride[1] + 1 == 1:
            raise RuntimeError("ConvWithNorms: 1x1 outputs skip BatchNorm in the reference (basic/__init__.py:73-76); "
                               "that branch never triggers on the DeFlow path and is not built")
        dt = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        y = tc.conv_bn_gelu(_to_nhwc(x, dt), self.conv, self.batchnorm, self.training)
        tc.flush_batch_counters()
        return y.permute(0, 3, 1, 2)


class BilinearDecoder(nn.Module):
    """unet.py:8-18 (scale factor 2, bilinear, align_corners=False)."""

    def __init__(self, scale_factor: int):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from . import conv as tc
        if self.scale_factor != 2:
            raise RuntimeError("BilinearDecoder: the kernel covers scale_factor=2 (every use in FastFlow3DUNet)")
        dt = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        return tc.upsample_bilinear2x(_to_nhwc(x, dt)).permute(0, 3, 1, 2)


class UpsampleSkip(nn.Module):
    """unet.py:21-37: cat([bilinear2x(conv1x1(a)), conv1x1(b)]) -> conv3x3 -> conv3x3; bias, no norm/activation."""

    def __init__(self, skip_channels: int, latent_channels: int, out_channels: int):
        super().__init__()
        self.u1_u2 = nn.Sequential(nn.Conv2d(skip_channels, latent_channels, 1, 1, 0), BilinearDecoder(2))
        self.u3 = nn.Conv2d(latent_channels, latent_channels, 1, 1, 0)
        self.u4_u5 = nn.Sequential(nn.Conv2d(2 * latent_channels, out_channels, 3, 1, 1),
                                   nn.Conv2d(out_channels, out_channels, 3, 1, 1))

    def forward_nhwc(self, a, b):
        """a, b: tuples of 1-2 NHWC tensors (a channel concatenation is passed as its parts and never materialised)."""
        from . import conv as tc
        c1 = self.u1_u2[0]
        u1 = tc.conv_bias(c1.weight, c1.bias, *a)
        u2 = tc.upsample_bilinear2x(u1)
        u3 = tc.conv_bias(self.u3.weight, self.u3.bias, *b)
        u4 = tc.conv_bias(self.u4_u5[0].weight, self.u4_u5[0].bias, u2, u3)
        return tc.conv_bias(self.u4_u5[1].weight, self.u4_u5[1].bias, u4)

    def forward(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        dt = torch.bfloat16 if a.dtype == torch.bfloat16 else torch.float32
        return self.forward_nhwc((_to_nhwc(a, dt),), (_to_nhwc(b, dt),)).permute(0, 3, 1, 2)


class FastFlow3DUNet(nn.Module):
    """unet.py:40-100."""

    def __init__(self) -> None:
        super().__init__()
        self.encoder_step_1 = nn.Sequential(ConvWithNorms(32, 64, 3, 2, 1), *[ConvWithNorms(64, 64, 3, 1, 1) for _ in range(3)])
        self.encoder_step_2 = nn.Sequential(ConvWithNorms(64, 128, 3, 2, 1), *[ConvWithNorms(128, 128, 3, 1, 1) for _ in range(5)])
        self.encoder_step_3 = nn.Sequential(ConvWithNorms(128, 256, 3, 2, 1), *[ConvWithNorms(256, 256, 3, 1, 1) for _ in range(5)])
        self.decoder_step1 = UpsampleSkip(512, 256, 256)
        self.decoder_step2 = UpsampleSkip(256, 128, 128)
        self.decoder_step3 = UpsampleSkip(128, 64, 64)
        self.decoder_step4 = nn.Conv2d(64, 64, 3, 1, 1)
        self.compute_dtype = torch.float32

    def forward(self, pc0_B: torch.Tensor, pc1_B: torch.Tensor) -> torch.Tensor:
        """Reference signature: [B,32,H,W] x2 -> [B,64,H,W] (any memory format; channels-last is the fast one)."""
        return self.forward_nhwc(_to_nhwc(pc0_B, self.compute_dtype), _to_nhwc(pc1_B, self.compute_dtype)).permute(0, 3, 1, 2)

    def forward_nhwc(self, img0: torch.Tensor, img1: torch.Tensor, skip_imgs=None) -> torch.Tensor:
        """NHWC [B,H,W,32] x2 -> NHWC [B,H,W,64].  skip_imgs: optional aliases of (img0, img1) for the last skip connection
        (DeFlow.forward hands over separate autograd edges so that the two gradients of an image arrive separately)."""
        if skip_imgs is not None:
            skip_imgs = (skip_imgs[0].contiguous(), skip_imgs[1].contiguous())
        return self._forward_tensor_core(img0.contiguous(), img1.contiguous(), skip_imgs)

    # Every convolution is a tcgen05 implicit GEMM (csrc/conv_igemm.cu); BatchNorm statistics come from the convolution
    # epilogue; BN+GELU, bilinear x2 and their backward are HBM-bound passes (csrc/unet_elem.cu).  Channel
    # concatenations are never materialised: the consuming convolution reads two sources.
    #   bf16 images  -> perf mode: bf16 operands, fp32 accumulation;
    #   fp32 images  -> parity mode: every operand is a (hi, lo) bf16 pair and the K loop runs hi*hi + hi*lo + lo*hi
    #                   ("bf16x3": ~16 significant bits per operand, fp32 accumulation and fp32 activations in HBM).
    def _weight_bank(self):
        from . import conv as tc
        bank = self.__dict__.get("_dfb_bank")
        if bank is None:
            bank = tc.WeightBank([m.weight for m in self.modules() if isinstance(m, nn.Conv2d)])
            self.__dict__["_dfb_bank"] = bank
        return bank

    def _forward_tensor_core(self, img0, img1, skip_imgs=None):
        from . import conv as tc
        with tc.bank_scope(self._weight_bank(), img0.dtype == torch.float32):
            return self._forward_packed(img0, img1, skip_imgs)

    def _forward_packed(self, img0, img1, skip_imgs=None):
        from . import conv as tc

        def encoder(x):
            outs = []
            for step in (self.encoder_step_1, self.encoder_step_2, self.encoder_step_3):
                for layer in step:
                    x = tc.conv_bn_gelu(x, layer.conv, layer.batchnorm, self.training)
                outs.append(x)
            return outs

        f0, l0, r0 = encoder(img0)
        f1, l1, r1 = encoder(img1)
        s = self.decoder_step1.forward_nhwc((r0, r1), (l0, l1))
        t = self.decoder_step2.forward_nhwc((s,), (f0, f1))
        u = self.decoder_step3.forward_nhwc((t,), skip_imgs if skip_imgs is not None else (img0, img1))
        tc.flush_batch_counters()
        return tc.conv_bias(self.decoder_step4.weight, self.decoder_step4.bias, u)
