"""DeFlow / FastFlow3D with the reference's module surface (OSF/src/models/deflow.py:20-114,
OSF/src/models/fastflow3d.py:42-103): same constructor, same ``forward(batch) -> dict`` contract, same
``state_dict`` keys -- computed by one batched pass over all 2B frames instead of per-sample loops.
"""
from __future__ import annotations

import os
import time
from typing import Dict

import torch
from torch import nn

from . import ops
from .decoder import ConvGRUDecoder, LinearDecoder
from .encoder import DynamicEmbedder
from .unet import FastFlow3DUNet


class _Timing:
    """Stand-in for ``dztimer.Timing`` (a wall-clock tree with no arithmetic, OSF/src/models/deflow.py:38-39):
    the trainer calls ``model.timer[i].start(name) / .stop() / .print()`` (OSF/src/trainer.py:95-99,184)."""

    def __init__(self):
        self._t, self._children = {}, {}

    def __getitem__(self, i):
        return self._children.setdefault(i, _Timing())

    def start(self, name=""):
        self._name, self._t0 = name, time.perf_counter()

    def stop(self):
        if hasattr(self, "_t0"):
            self._t[self._name] = self._t.get(self._name, 0.0) + time.perf_counter() - self._t0

    def print(self, *a, **k):
        for name, t in self._t.items():
            print(f"{name}: {t:.3f}s")
        for c in self._children.values():
            c.print()


def weights_init(m):
    """OSF/src/utils/mics.py:98-105: Xavier-uniform on Conv2d / Linear with zero bias, BatchNorm2d 1/0;
    Conv1d and BatchNorm1d keep the torch defaults."""
    if isinstance(m, (nn.Conv2d, nn.Linear)):
        nn.init.xavier_uniform_(m.weight)
        if m.bias is not None:
            nn.init.zeros_(m.bias)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


class _SplitFrames(torch.autograd.Function):
    """image [2B,H,W,C] -> (image[:B], image[B:]) twice: one pair for the encoder, one for the last skip convolution -- the
    two consumers of each frame half -- so that the backward receives their gradients SEPARATELY and builds the gradient of
    the whole image in one pass: out = cat(g_enc0 + g_skip0, g_enc1 + g_skip1) (ops.add_cat2), then the rows of the decoder
    gather's gradient that belong to the image are added in place (``sink``, see ops._DecoderGather.backward).  Plain
    slicing + autograd made this two dense additions, a zero-filled dense gather gradient, a concatenation and a third dense
    addition (about 5 GB of HBM traffic per step at 512^2 x 32 frames)."""

    @staticmethod
    def forward(ctx, image, B, sink):
        ctx.meta = (B, image.shape, image.dtype, image.device)
        ctx.sink = sink
        base = image.detach()            # same storage, no autograd view relation to the input
        return base[:B], base[B:], base[:B], base[B:]

    @staticmethod
    def backward(ctx, ga0, ga1, gb0, gb1):
        B, shape, dtype, device = ctx.meta
        half = (B,) + tuple(shape[1:])
        pend = ctx.sink.pop("gather", None) if ctx.sink is not None else None
        if ga0 is None and ga1 is None and gb0 is None and gb1 is None and pend is None:
            return None, None, None
        z = lambda: torch.zeros(half, dtype=dtype, device=device)  # noqa: E731
        if ga0 is None and gb0 is not None:
            ga0, gb0 = gb0, None
        if ga1 is None and gb1 is not None:
            ga1, gb1 = gb1, None
        ga0 = z() if ga0 is None else ga0.contiguous()
        ga1 = z() if ga1 is None else ga1.contiguous()
        if (gb0 is None) != (gb1 is None):
            gb0 = z() if gb0 is None else gb0
            gb1 = z() if gb1 is None else gb1
        if gb0 is not None:
            gb0, gb1 = gb0.contiguous(), gb1.contiguous()
        out = ops.add_cat2(ga0, gb0, ga1, gb1)
        if pend is not None:
            rows, idx, Bp, H, W = pend
            ops.gather_img_rows_add(rows, idx, Bp, H, W, out)
        return out, None, None


class DeFlow(nn.Module):
    def __init__(self, voxel_size=[0.2, 0.2, 6], point_cloud_range=[-51.2, -51.2, -3, 51.2, 51.2, 3],
                 grid_feature_size=[512, 512], decoder_option="gru", num_iters=4, precision="fp32"):
        super().__init__()
        # constructor arguments as the reference's hydra target spells them (OSF/conf/model/deflow.yaml:4-9,
        # grid_feature_size added by ModelWrapper.__init__, OSF/src/trainer.py:42-55): goes into a checkpoint's
        # ``hyper_parameters`` (trainer.TrainStep.state_dict), which OSF/eval.py:41-46 reads back
        self.target_cfg = {"_target_": "src.models.DeFlow", "decoder_option": decoder_option, "num_iters": int(num_iters),
                           "voxel_size": [float(v) for v in voxel_size],
                           "point_cloud_range": [float(v) for v in point_cloud_range],
                           "grid_feature_size": [int(v) for v in grid_feature_size]}
        self.embedder = DynamicEmbedder(voxel_size=voxel_size, pseudo_image_dims=grid_feature_size,
                                        point_cloud_range=point_cloud_range, feat_channels=32)
        self.backbone = FastFlow3DUNet()
        if decoder_option == "gru":
            self.head = ConvGRUDecoder(num_iters=num_iters)
        elif decoder_option == "linear":
            self.head = LinearDecoder()
        else:
            raise ValueError(f"unknown decoder_option {decoder_option}")
        self.timer = _Timing()
        self.timer.start("Total")
        self.set_precision(precision)

    def set_precision(self, precision: str):
        """'fp32' = parity mode: fp32 tensors everywhere, the dense contractions on the tensor cores with split-precision
        ("bf16x3") operands; 'bf16' = perf mode (bf16 operands and activations, fp32 accumulation / statistics /
        voxelisation) -- SURVEY.md section 8(d).  Both run on the same tcgen05 kernels; there is no library backend."""
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32' or 'bf16'")
        self.precision = precision
        dt = torch.bfloat16 if precision == "bf16" else torch.float32
        for m in (self.backbone, self.head):
            m.compute_dtype = dt
        return self

    def load_from_checkpoint(self, ckpt_path):
        """deflow.py:41-47: Lightning checkpoint with 'model.'-prefixed keys, strict=False."""
        ckpt = torch.load(ckpt_path, map_location="cpu")["state_dict"]
        state_dict = {k[len("model."):]: v for k, v in ckpt.items() if k.startswith("model.")}
        print("\nLoading... model weight from: ", ckpt_path, "\n")
        return self.load_state_dict(state_dict=state_dict, strict=False)

    def forward(self, batch: Dict) -> Dict:
        pc0, pc1 = batch["pc0"], batch["pc1"]
        if not pc0.is_cuda:
            raise RuntimeError("deflow_b200.DeFlow runs on CUDA (sm_100a) only; there is no CPU path")
        B, N0, N1 = pc0.shape[0], pc0.shape[1], pc1.shape[1]
        assert len(batch["pose0"]) == B
        self.timer[0].start("Data Preprocess")
        Nm = max(N0, N1)
        pts = torch.empty((2 * B, Nm, 3), dtype=torch.float32, device=pc0.device)
        if N0 < Nm:
            pts[:B, N0:].fill_(float("nan"))
        if N1 < Nm:
            pts[B:, N1:].fill_(float("nan"))
        pts[B:, :N1].copy_(pc1)
        if "ego_motion" in batch:
            ego = batch["ego_motion"]
            ego = torch.stack(list(ego), 0) if not torch.is_tensor(ego) else ego
            pose_flow, _ = ops.ego_warp(pc0, None, None, ego.to(pc0.device), pts)
        else:
            p0 = torch.stack(list(batch["pose0"]), 0) if not torch.is_tensor(batch["pose0"]) else batch["pose0"]
            p1 = torch.stack(list(batch["pose1"]), 0) if not torch.is_tensor(batch["pose1"]) else batch["pose1"]
            pose_flow, _ = ops.ego_warp(pc0, p0.to(pc0.device), p1.to(pc0.device), None, pts)
        self.timer[0].stop()

        self.timer[1].start("Voxelization")
        img_dtype = torch.bfloat16 if self.precision == "bf16" else torch.float32
        image, idx = self.embedder.embed(pts, img_dtype)  # [2B,H,W,32]: frames 0..B-1 = pc0, B..2B-1 = pc1
        self.timer[1].stop()

        self.timer[2].start("Encoder")
        # the pseudo-image has three consumers (encoder, last skip convolution, decoder gather): their gradients meet in
        # _SplitFrames.backward (DFB_IMG_GRAD_FUSE=0: the gather returns its own dense gradient and autograd adds it)
        fuse = torch.is_grad_enabled() and image.requires_grad and os.environ.get("DFB_IMG_GRAD_FUSE", "1") != "0"
        sink = {} if fuse else None
        img0, img1, img0s, img1s = _SplitFrames.apply(image, B, sink)
        unet_out = self.backbone.forward_nhwc(img0, img1, skip_imgs=(img0s, img1s))
        self.timer[2].stop()

        self.timer[3].start("Decoder")
        n0 = idx.pt_off(B)  # the one host sync of the forward
        flow_flat = self.head.forward_flat(image, unet_out, idx, B, n0, sink=sink)
        self.timer[3].stop()

        infos = [idx.frame_info(f) for f in range(2 * B)]
        return {
            "flow": [flow_flat[idx.pt_off(b):idx.pt_off(b + 1)] for b in range(B)],
            "pose_flow": [pose_flow[b] for b in range(B)],
            "pc0_valid_point_idxes": [e["point_idxes"] for e in infos[:B]],
            "pc0_points_lst": [e["points"] for e in infos[:B]],
            "pc1_valid_point_idxes": [e["point_idxes"] for e in infos[B:]],
            "pc1_points_lst": [e["points"] for e in infos[B:]],
            "num_occupied_voxels": [unet_out.shape[1] * unet_out.shape[2]],  # dense size, deflow.py:112
            # flat handles for the fused loss (lossfuncs.training_step_loss); not part of the reference dict
            "_dfb": {"index": idx, "flow_flat": flow_flat, "pose_flow": pose_flow, "B": B,
                     "image": image, "unet": unet_out},
        }


class FastFlow3D(DeFlow):
    """OSF/src/models/fastflow3d.py:42-103 (plain forward; the unused cycle / symmetry passes are not built)."""

    def __init__(self, voxel_size=[0.2, 0.2, 6], point_cloud_range=[-51.2, -51.2, -3, 51.2, 51.2, 3],
                 grid_feature_size=[512, 512], precision="fp32"):
        super().__init__(voxel_size, point_cloud_range, grid_feature_size, decoder_option="linear",
                         precision=precision)
        self.target_cfg = {"_target_": "src.models.FastFlow3D", "voxel_size": self.target_cfg["voxel_size"],
                           "point_cloud_range": self.target_cfg["point_cloud_range"],
                           "grid_feature_size": self.target_cfg["grid_feature_size"]}
