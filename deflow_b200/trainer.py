"""Lightning-free training step with the arithmetic of ``ModelWrapper.training_step`` + ``configure_optimizers``
(OSF/src/trainer.py:94-175, OSF/train.py:119-142): forward, gt = flow[idx] - pose_flow[idx], per-sample losses
SUMMED over the batch, backward, gradient mean across ranks, clip-by-norm 5.0 (OSF/conf/config.yaml:26), Adam.

    step = TrainStep(model, lr=2e-4, loss_fn="deflowLoss")
    loss = step(batch)            # batch: collate_fn_pad layout (OSF/src/dataset.py:22-74), tensors on the GPU
"""
from __future__ import annotations

import torch

from . import dist as dd
from .lossfuncs import training_step_loss


class TrainStep:
    def __init__(self, model: torch.nn.Module, lr: float = 2e-4, loss_fn: str = "deflowLoss",
                 gradient_clip_val: float = 5.0):
        self.model = model
        self.loss_fn = loss_fn
        self.clip = gradient_clip_val
        self.grads = dd.GradAverager(model.parameters())
        # one forward + backward per call: the pseudo-image of a step is dead before the next step's embed(), so the
        # embedder may keep its canvas and clear only the previous step's pillar rows (encoder.DynamicEmbedder.reuse_canvas)
        emb = getattr(model, "embedder", None)
        if emb is not None and hasattr(emb, "reuse_canvas"):
            emb.reuse_canvas = True
        self.opt = torch.optim.Adam(self.grads.params, lr=lr, fused=self.grads.flat.is_cuda)  # trainer.py:173-175

    def __call__(self, batch) -> torch.Tensor:
        self.grads.zero()
        res = self.model(batch)
        loss = training_step_loss(batch, res, self.loss_fn)
        loss.backward()
        self.grads.average()
        if self.clip is not None and self.clip > 0:
            # torch.nn.utils.clip_grad_norm_ semantics on the flat buffer: scale by min(1, clip / (norm + 1e-6))
            norm = torch.linalg.vector_norm(self.grads.flat)
            self.grads.flat.mul_(torch.clamp(self.clip / (norm + 1e-6), max=1.0))
        self.opt.step()
        return loss.detach()

    def state_dict(self):
        """Checkpoint in the reference's layout: 'model.'-prefixed keys under 'state_dict' (deflow.py:41-47)."""
        return {"state_dict": {"model." + k: v for k, v in self.model.state_dict().items()},
                "optimizer_states": [self.opt.state_dict()]}
