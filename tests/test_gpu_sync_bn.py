"""SyncBatchNorm semantics (the reference's default sync_bn: true, OSF/conf/config.yaml:23, OSF/train.py:128) of the
opt-in statistics exchange: two data-parallel ranks -- two processes sharing cuda:0, gloo collectives, so it runs on a
one-GPU box -- against the CPU oracle evaluated on the pooled batch (SyncBN over R ranks == BatchNorm over the
concatenated batch; the per-sample BatchNorm1d of the pillar feature net pools sample b of every rank)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
SMALL = dict(vs=[0.2, 0.2, 6], rg=[-6.4, -6.4, -3, 6.4, 6.4, 3], grid=[64, 64])
KEYS = ["backbone.encoder_step_1.0.conv.weight", "backbone.encoder_step_1.0.batchnorm.weight",
        "backbone.encoder_step_3.5.batchnorm.bias", "backbone.decoder_step4.weight", "head.gru.convz.weight",
        "embedder.feature_net.pfn_layers.0.0.weight", "embedder.feature_net.pfn_layers.0.1.weight",
        "embedder.feature_net.pfn_layers.0.1.bias"]
BUFS = ["embedder.feature_net.pfn_layers.0.1.running_mean", "embedder.feature_net.pfn_layers.0.1.running_var",
        "backbone.encoder_step_1.0.batchnorm.running_mean", "backbone.encoder_step_2.3.batchnorm.running_var"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, sync_bn, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": "0"})
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    import deflow_b200 as d
    from deflow_b200 import dist as dd
    from oracle import deflow_oracle as orc
    from helpers import load_fixture, batch_to
    torch.cuda.set_device(0)
    dd.init("gloo")
    fx, batch, cfg = load_fixture("deflow_small_gru")
    mine = {k: (v[rank:rank + 1] if torch.is_tensor(v) else v[rank:rank + 1]) for k, v in batch.items()}   # sample `rank`
    m = d.DeFlow(cfg["voxel_size"], cfg["range"], cfg["grid"], "gru", 4, precision="fp32")
    m.load_state_dict(orc.random_state(11, "gru"), strict=True)
    m = m.to("cuda:0").train()
    sync = dd.enable_sync_bn(m) if sync_bn else None
    avg = dd.GradAverager(m.parameters())
    gb = batch_to(mine, "cuda:0")
    res = m(gb)
    loss = d.training_step_loss(gb, res, "deflowLoss")
    loss.backward()
    avg.average()
    sd = m.state_dict()
    named = dict(m.named_parameters())
    out[rank] = {"loss": float(loss), "flow": res["flow"][0].detach().cpu(),
                 "grads": {k: named[k].grad.detach().cpu().clone() for k in KEYS},
                 "bufs": {k: sd[k].detach().cpu().clone() for k in BUFS},
                 "exchanges": (sync.calls, sync.bytes) if sync else (0, 0)}
    dd.barrier()
    dist.destroy_process_group()


def _oracle(sync_world):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle import deflow_oracle as orc
    from helpers import load_fixture
    fx, batch, cfg = load_fixture("deflow_small_gru")
    state = orc.random_state(11, "gru")
    for k, v in state.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    buffers = {k: v.clone() for k, v in state.items() if "running" in k}
    res = orc.deflow_forward(batch, state, cfg["voxel_size"], cfg["range"], cfg["grid"], "gru", 4, training=True,
                             buffers=buffers, sync_world=sync_world)
    per = []
    for b in range(2):
        one = {k: (v[b:b + 1] if torch.is_tensor(v) else v[b:b + 1]) for k, v in batch.items()}
        r1 = {k: (v[b:b + 1] if isinstance(v, list) and len(v) == 2 else v) for k, v in res.items()}
        per.append(orc.training_step_loss(one, r1, "deflowLoss"))
    (per[0] + per[1]).backward()
    return res, per, state, buffers


def test_sync_bn_two_ranks_equal_oracle_on_the_pooled_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, True, out), nprocs=world, join=True)
    res, per, state, buffers = _oracle(sync_world=2)
    for r in range(2):
        got = out[r]
        assert float((got["flow"] - res["flow"][r].detach()).abs().max()) <= 1e-3
        assert abs(got["loss"] - float(per[r])) <= 2e-4 * max(1.0, abs(float(per[r])))
        for k in KEYS:      # averaged over the two ranks == half the gradient of the summed loss
            ref = 0.5 * state[k].grad
            assert float((got["grads"][k] - ref).norm()) <= 3e-3 * float(ref.norm()) + 1e-6, k
        for k in BUFS:
            np.testing.assert_allclose(got["bufs"][k].numpy(), buffers[k].numpy(), rtol=2e-4, atol=2e-5, err_msg=k)
    # 32 BatchNorm2d layer calls + 1 fused BatchNorm1d call (2 tensors), forward and backward
    assert out[0]["exchanges"][0] == 32 + 2 + 32 + 1


def test_without_sync_bn_ranks_use_their_own_statistics():
    """The default (per-rank statistics, what SCALE measures): each rank equals the single-sample oracle."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, False, out), nprocs=world, join=True)
    res, per, state, buffers = _oracle(sync_world=2)
    # pooled statistics give a measurably different flow than per-rank statistics
    d0 = float((out[0]["flow"] - res["flow"][0].detach()).abs().max())
    assert d0 > 1e-3, d0
