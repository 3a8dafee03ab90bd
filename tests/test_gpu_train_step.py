"""The Lightning-free training step (deflow_b200/trainer.py: arithmetic of OSF/src/trainer.py:94-175) and the
double-buffered input feed, on the GPU."""
import numpy as np
import pytest
import torch

import deflow_b200 as d
from deflow_b200 import synth
from deflow_b200.feed import DeviceFeeder
from deflow_b200.trainer import TrainStep
from oracle import deflow_oracle as orc
from helpers import load_fixture, batch_to

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _small_model(precision):
    m = d.DeFlow([0.2, 0.2, 6], [-6.4, -6.4, -3, 6.4, 6.4, 3], [64, 64], "gru", 4, precision=precision)
    m.load_state_dict(orc.random_state(11, "gru"), strict=True)
    return m.to(DEV).train()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_first_step_matches_reference_loss_and_training_reduces_it(precision):
    fx, batch, cfg = load_fixture("deflow_small_gru")
    m = _small_model(precision)
    step = TrainStep(m, lr=1e-4, loss_fn="deflowLoss", gradient_clip_val=5.0)
    gb = batch_to(batch, DEV)
    losses = [float(step(gb)) for _ in range(10)]
    tol = 2e-4 if precision == "fp32" else 2e-2
    assert abs(losses[0] - float(fx["loss_total"])) <= tol * max(1.0, abs(float(fx["loss_total"])))  # reference loss at step 0
    assert min(losses[1:]) < losses[0]                                                                    # Adam makes progress
    assert all(np.isfinite(losses))
    # gradients live in ONE flat buffer (the all-reduce operand) and were clipped to norm <= 5
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(step.grads.params, step.grads.views))
    assert float(torch.linalg.vector_norm(step.grads.flat)) <= 5.0 + 1e-3
    ck = step.state_dict()
    assert all(k.startswith("model.") for k in ck["state_dict"]) and len(ck["state_dict"]) == 156


def test_device_feeder_double_buffering():
    host = synth.make_batch(2, 3000, seed=4, pin=True)
    host["pose0"] = torch.stack(host["pose0"]).pin_memory()
    host["pose1"] = torch.stack(host["pose1"]).pin_memory()
    f = DeviceFeeder(DEV)
    f.submit(host)
    for _ in range(3):
        b = f.get()
        f.submit(host)
        assert b["pc0"].is_cuda and torch.equal(b["pc0"].cpu().nan_to_num(1e9), host["pc0"].nan_to_num(1e9))
        assert torch.equal(b["flow_category_indices"].cpu(), host["flow_category_indices"])


def test_checkpoint_round_trip(tmp_path):
    """Reference checkpoint layout: Lightning dict with 'model.'-prefixed keys (OSF/src/models/deflow.py:41-47)."""
    m = _small_model("fp32")
    step = TrainStep(m)
    path = tmp_path / "ck.ckpt"
    torch.save(step.state_dict(), path)
    m2 = d.DeFlow([0.2, 0.2, 6], [-6.4, -6.4, -3, 6.4, 6.4, 3], [64, 64], "gru", 4)
    res = m2.load_from_checkpoint(str(path))
    assert not res.missing_keys and not res.unexpected_keys
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
