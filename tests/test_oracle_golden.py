"""The CPU oracle against the fixtures produced by the reference's own modules
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import deflow_oracle as orc
from helpers import load_fixture

CASES = ["deflow_small_gru", "deflow_small_linear", "deflow_small_gru_eval"]


def _run(name, need_grad):
    fx, batch, cfg = load_fixture(name)
    state = orc.random_state(cfg["seed_state"], cfg["decoder"])
    if need_grad:
        for k, v in state.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    res = orc.deflow_forward(batch, state, cfg["voxel_size"], cfg["range"], cfg["grid"], cfg["decoder"], 4,
                             training=cfg["training"], buffers=buffers if cfg["training"] else None)
    return fx, batch, cfg, state, buffers, res


@pytest.mark.parametrize("name", CASES)
def test_forward_matches_reference(name):
    fx, batch, cfg, state, buffers, res = _run(name, False)
    for b in range(len(batch["pose0"])):
        assert np.array_equal(res["pc0_valid_point_idxes"][b].numpy(), fx[f"pc0_idx_{b}"])
        assert np.array_equal(res["pc1_valid_point_idxes"][b].numpy(), fx[f"pc1_idx_{b}"])
        np.testing.assert_allclose(res["pose_flow"][b].numpy(), fx[f"pose_flow_{b}"], atol=1e-6)
        np.testing.assert_allclose(res["flow"][b].detach().numpy(), fx[f"flow_{b}"], atol=2e-5)
    loss = orc.training_step_loss(batch, res, cfg["loss"])
    assert abs(float(loss) - float(fx["loss_total"])) < 1e-5
    assert res["num_occupied_voxels"] == list(fx["num_occupied_voxels"])


@pytest.mark.parametrize("name", ["deflow_small_gru", "deflow_small_linear"])
def test_backward_and_running_stats_match_reference(name):
    fx, batch, cfg, state, buffers, res = _run(name, True)
    loss = orc.training_step_loss(batch, res, cfg["loss"])
    loss.backward()
    for k in fx:
        if k.startswith("grad::"):
            g = state[k[6:]].grad.numpy()
            ref = fx[k]
            assert np.abs(g - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max()), k
        if k.startswith("buf::") and "num_batches" not in k:
            np.testing.assert_allclose(buffers[k[5:]].numpy(), fx[k], rtol=1e-5, atol=1e-6, err_msg=k)


def test_config1_20k_and_real_sweep():
    for name in ("deflow_cfg1_20k", "deflow_av2_real_20k"):
        fx, batch, cfg, state, buffers, res = _run(name, False)
        assert np.array_equal(res["pc0_valid_point_idxes"][0].numpy(), fx["pc0_idx_0"])
        np.testing.assert_allclose(res["flow"][0].detach().numpy(), fx["flow_0"], atol=5e-5)


@pytest.mark.parametrize("name,tol", [("train4_small_gru_lr2e-5", 2e-5), ("train3_small_gru_lr2e-4", 3e-4)])
def test_training_trajectory_matches_reference(name, tol):
    """K Adam steps of the oracle (clip 5.0) against the trajectory the reference modules produced
    (make_golden.run_train_case; OSF/src/trainer.py:94-175).  The lr 2e-4 trajectory overshoots and amplifies
    arithmetic noise (1 vs 8 threads of the reference itself: 2e-5), hence its wider tolerance."""
    tz = np.load(f"{__import__('helpers').GOLDEN}/{name}.npz")
    fx, batch, cfg = load_fixture(str(tz["fixture"]))
    state = orc.random_state(int(tz["seed_state"]), str(tz["decoder"]))
    params = [v for k, v in state.items() if v.is_floating_point() and "running" not in k]
    for v in params:
        v.requires_grad_(True)
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    opt = torch.optim.Adam(params, lr=float(tz["lr"]))
    losses = []
    for _ in range(int(tz["steps"])):
        opt.zero_grad(set_to_none=True)
        res = orc.deflow_forward(batch, state, cfg["voxel_size"], cfg["range"], cfg["grid"], cfg["decoder"], 4,
                                 training=True, buffers=buffers)
        loss = orc.training_step_loss(batch, res, str(tz["loss_name"]))
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, float(tz["clip"]))
        opt.step()
        losses.append(float(loss))
    np.testing.assert_allclose(losses, tz["losses"], rtol=tol)
    for k in tz.files:
        if k.startswith("weight::"):
            w0 = orc.random_state(int(tz["seed_state"]), str(tz["decoder"]))[k[8:]].numpy()
            d_ref, d = tz[k] - w0, state[k[8:]].detach().numpy() - w0
            assert np.linalg.norm(d - d_ref) <= 50 * tol * np.linalg.norm(d_ref), k


def test_loss_restatements_equal_reference_functions():
    """deflowLoss / zeroflowLoss / ff3dLoss restatements against the reference's own functions (AST-extracted from
    OSF/src/lossfuncs.py:102-157)."""
    from oracle import ref_modules
    if ref_modules.root() is None:
        pytest.skip("reference sources not available")
    ns = ref_modules.extract_functions("src/lossfuncs.py", ("deflowLoss", "ff3dLoss", "zeroflowLoss"))
    g = torch.Generator().manual_seed(0)
    n = 4000
    gt = torch.randn(n, 3, generator=g) * torch.tensor([0.0, 0.03, 0.3])[torch.randint(0, 3, (n, 1), generator=g)]
    est = gt + 0.05 * torch.randn(n, 3, generator=g)
    gt[5] = float("nan")
    cls = torch.randint(0, 3, (n,), generator=g).to(torch.uint8)
    assert float(ns["deflowLoss"]({"est_flow": est, "gt_flow": gt})["loss"]) == pytest.approx(float(orc.deflow_loss(est, gt)), rel=1e-6)
    assert float(ns["zeroflowLoss"]({"est_flow": est, "gt_flow": gt})["loss"]) == pytest.approx(float(orc.zeroflow_loss(est, gt)), rel=1e-6)
    gt[5] = 0.0
    assert float(ns["ff3dLoss"]({"est_flow": est, "gt_flow": gt, "gt_classes": cls})["loss"]) == pytest.approx(
        float(orc.ff3d_loss(est, gt, cls)), rel=1e-6)


def test_seflow_loss_restatement_equals_reference_function():
    """seflowLoss (OSF/src/lossfuncs.py:22-100): the reference's own function (AST-extracted, its MyCUDAChamferDis bound to the
    CPU stand-in of the chamfer op) against oracle/seflow_oracle.seflow_loss, values and gradient."""
    from oracle import ref_modules, seflow_oracle as so
    if ref_modules.root() is None:
        pytest.skip("reference sources not available")
    ns = ref_modules.extract_functions("src/lossfuncs.py", ("seflowLoss",),
                                       {"MyCUDAChamferDis": so.ChamferStandIn(), "TRUNCATED_DIST": 4})
    for n0, n1, seed in ((3000, 2800, 1), (900, 1000, 2), (400, 300, 3)):      # the last has no dynamic cluster (<= 256 points)
        sc = so.make_scene(n0, n1, seed)
        e1 = sc["est_flow"].clone().requires_grad_(True)
        e2 = sc["est_flow"].clone().requires_grad_(True)
        ref = ns["seflowLoss"]({**sc, "est_flow": e1})
        got = so.seflow_loss({**sc, "est_flow": e2})
        assert set(ref) == set(got)
        for k in ref:
            assert float(ref[k]) == pytest.approx(float(got[k]), rel=1e-6, abs=1e-7), k
        sum(ref.values()).backward()
        sum(got.values()).backward()
        np.testing.assert_allclose(e2.grad.numpy(), e1.grad.numpy(), rtol=1e-5, atol=1e-8)
