"""Lightning-free training step with the arithmetic of ``ModelWrapper.training_step`` + ``configure_optimizers``
(OSF/src/trainer.py:94-175, OSF/train.py:119-142): forward, gt = flow[idx] - pose_flow[idx], per-sample losses
SUMMED over the batch, backward, gradient mean across ranks, clip-by-norm 5.0 (OSF/conf/config.yaml:26), Adam.

    step = TrainStep(model, lr=2e-4, loss_fn="deflowLoss")
    loss = step(batch)            # batch: collate_fn_pad layout (OSF/src/dataset.py:22-74), tensors on the GPU
    torch.save(step.state_dict(), "last.ckpt")        # reference checkpoint layout, see state_dict()
    step.load_state_dict(torch.load("last.ckpt"))     # resume: weights, Adam moments, step / epoch counters
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import dist as dd
from .lossfuncs import training_step_loss


class TrainStep:
    def __init__(self, model: torch.nn.Module, lr: float = 2e-4, loss_fn: str = "deflowLoss",
                 gradient_clip_val: float = 5.0, sync_bn: bool = False, overlap_allreduce: bool = False):
        self.model = model
        self.loss_fn = loss_fn
        self.lr = lr
        self.clip = gradient_clip_val
        self.sync_bn = bool(sync_bn)
        self.grads = dd.GradAverager(model.parameters())
        self.opt = torch.optim.Adam(self.grads.params, lr=lr, fused=self.grads.flat.is_cuda)  # trainer.py:173-175
        self.global_step = 0
        self.epoch = 0
        # Gradient mean: ONE ReduceOp.AVG collective on the flat buffer after the backward (default).  Opt-in
        # (overlap_allreduce=True or DFB_ALLREDUCE_OVERLAP=1): the head + UNet-decoder slice is all-reduced from a hook while
        # the encoder backward still runs.  Measured on 2 x B200 (profiles/r02_allreduce_overlap_ab.txt): 32.02-32.10 ms
        # with the overlap against 31.96-31.98 ms without (N = 1 on the same box: 31.83 ms) -- the 27.6 MB all-reduce costs
        # ~0.13 ms on NVSwitch, and the NCCL kernel that overlaps the backward takes SMs away from persistent kernels
        # that are sized one CTA per SM, so overlapping loses slightly more than it hides.
        want = overlap_allreduce or os.environ.get("DFB_ALLREDUCE_OVERLAP", "0") == "1"
        self.overlap = want and dd.world_size() > 1 and self.grads.flat.is_cuda
        if self.overlap:
            self.grads.plan_early_slice(model)
        self.stat_sync = dd.enable_sync_bn(model) if self.sync_bn else None
        # the UNet's weight gradients are unpacked once per step straight into the flat gradient buffer (conv.WeightBank);
        # not with the overlapped all-reduce, whose hooks wait for per-parameter gradients
        bb = getattr(model, "backbone", None)
        if (bb is not None and hasattr(bb, "_weight_bank") and self.grads.flat.is_cuda and not self.overlap
                and os.environ.get("DFB_DEFER_WGRAD", "1") != "0"):
            views = {id(p): v for p, v in zip(self.grads.params, self.grads.views)}
            bank = bb._weight_bank()
            bank.enable_deferred_grads({id(w): views[id(w)] for w in bank.weights if id(w) in views})

    def __call__(self, batch) -> torch.Tensor:
        self.grads.zero()
        # one forward + backward per call: the pseudo-image of a step is dead before the next step's embed(), so the
        # embedder may keep its canvas and clear only the previous step's pillar rows (DynamicEmbedder.reuse_canvas).
        # Scoped to this call: a validation forward between two training steps gets a canvas of its own and cannot
        # overwrite activations a live graph still needs.
        emb = getattr(self.model, "embedder", None)
        scoped = emb is not None and hasattr(emb, "reuse_canvas") and not emb.reuse_canvas
        if scoped:
            emb.reuse_canvas = True
        try:
            res = self.model(batch)
            loss = training_step_loss(batch, res, self.loss_fn)
            if self.overlap:
                self.grads.arm_early_slice()
            loss.backward()
        finally:
            if scoped:
                emb.reuse_canvas = False
        self.grads.average()
        if self.clip is not None and self.clip > 0:
            # torch.nn.utils.clip_grad_norm_ semantics on the flat buffer: scale by min(1, clip / (norm + 1e-6))
            norm = torch.linalg.vector_norm(self.grads.flat)
            self.grads.flat.mul_(torch.clamp(self.clip / (norm + 1e-6), max=1.0))
        self.opt.step()
        self.global_step += 1
        return loss.detach()

    # ------------------------------------------------------------------------------------------ checkpoints
    def hyper_parameters(self, extra: Optional[dict] = None) -> dict:
        """What ``ModelWrapper.save_hyperparameters()`` stores (OSF/src/trainer.py:92: the init arguments ``cfg`` and
        ``eval``), restricted to the keys the reference reads back: ``cfg.model`` / ``cfg.model.name`` / ``cfg.output`` /
        ``cfg.num_frames`` (OSF/eval.py:41-46, 58-60), plus the training settings of OSF/conf/config.yaml."""
        tgt = dict(getattr(self.model, "target_cfg", {}))
        name = "fastflow3d" if tgt.get("_target_", "").endswith("FastFlow3D") else "deflow"
        cfg = {"model": {"name": name, "target": tgt, "val_monitor": "val/Dynamic/Mean"},
               "voxel_size": tgt.get("voxel_size"), "point_cloud_range": tgt.get("point_cloud_range"),
               "num_frames": 2, "output": f"{name}-00000", "lr": self.lr, "loss_fn": self.loss_fn,
               "gradient_clip_val": self.clip, "sync_bn": self.sync_bn, "gpus": dd.world_size(), "seed": 42069}
        if extra:
            cfg.update(extra)
        return {"cfg": cfg, "eval": False}

    def state_dict(self, extra_cfg: Optional[dict] = None):
        """Checkpoint in the reference's (Lightning) layout: 'model.'-prefixed keys under 'state_dict'
        (OSF/src/models/deflow.py:41-47), 'hyper_parameters' (OSF/src/trainer.py:92, consumed by OSF/eval.py:41-46),
        'optimizer_states' (resume through ``trainer.fit(ckpt_path=...)``, OSF/train.py:142), 'epoch', 'global_step'."""
        return {"state_dict": {"model." + k: v for k, v in self.model.state_dict().items()},
                "optimizer_states": [self.opt.state_dict()], "lr_schedulers": [],
                "hyper_parameters": self.hyper_parameters(extra_cfg),
                "epoch": self.epoch, "global_step": self.global_step}

    def load_state_dict(self, ckpt: dict, strict: bool = True):
        """Resume from ``state_dict()`` output or from a reference Lightning checkpoint (same keys)."""
        sd = {k[len("model."):]: v for k, v in ckpt["state_dict"].items() if k.startswith("model.")}
        res = self.model.load_state_dict(sd, strict=strict)
        states = ckpt.get("optimizer_states") or []
        if states:
            self.opt.load_state_dict(states[0])
            for g in self.opt.param_groups:      # keep the hyper-parameters this TrainStep was built with consistent
                self.lr = g["lr"]
        self.epoch = int(ckpt.get("epoch", 0))
        self.global_step = int(ckpt.get("global_step", 0))
        return res
