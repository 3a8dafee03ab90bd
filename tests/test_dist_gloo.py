"""World-size-2 gloo test (CPU) of the data-parallel plumbing: flat-buffer gradient averaging equals the
full-batch gradient of a single process, parameters are replicated from rank 0, timing is the max over ranks."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    import torch.distributed as dist
    from deflow_b200 import dist as dd
    dd.init("gloo")
    torch.manual_seed(100 + rank)  # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    dd.broadcast_module(net)
    avg = dd.GradAverager(net.parameters())
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]   # shard by sample, no data-path collective
    avg.zero()
    ((net(xs) - ys) ** 2).sum().backward()      # per-rank SUM loss, like the trainer's summed per-sample losses
    avg.average()
    t = dd.max_over_ranks(float(rank + 1), "cpu")
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(avg.params, avg.views))  # grads are views of the flat buffer
    out[rank] = ([p.detach().clone() for p in net.parameters()], [p.grad.clone() for p in net.parameters()], t)
    dd.barrier()
    dist.destroy_process_group()


def test_grad_average_matches_full_batch():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    p0, g0, t0 = out[0]
    p1, g1, t1 = out[1]
    for a, b in zip(p0, p1):
        assert torch.equal(a, b)                      # replicated from rank 0
    for a, b in zip(g0, g1):
        assert torch.allclose(a, b, atol=1e-7)        # identical averaged gradients on both ranks
    assert t0 == t1 == 2.0                            # max over ranks
    # single-process reference: mean over ranks of the per-rank summed loss == 0.5 * full-batch summed loss
    torch.manual_seed(100)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.GELU(), torch.nn.Linear(5, 3))
    g = torch.Generator().manual_seed(7)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    (0.5 * ((net(x) - y) ** 2).sum()).backward()
    for a, p in zip(g0, net.parameters()):
        assert torch.allclose(a, p.grad, atol=1e-6)


class _Toy(torch.nn.Module):
    """Parameter tree shaped like DeFlow's: embedder / backbone.encoder_step_* (late in the backward) and
    backbone.decoder_* / head (early)."""

    def __init__(self):
        super().__init__()
        self.embedder = torch.nn.Linear(6, 6)
        self.backbone = torch.nn.ModuleDict({"encoder_step_1": torch.nn.Linear(6, 5), "decoder_step1": torch.nn.Linear(5, 5)})
        self.head = torch.nn.Linear(5, 3)

    def forward(self, x):
        x = torch.tanh(self.embedder(x))
        x = torch.tanh(self.backbone["encoder_step_1"](x))
        return self.head(torch.tanh(self.backbone["decoder_step1"](x)))


def _worker_early(rank, world, port, out):
    os.environ.update({"MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port), "RANK": str(rank), "WORLD_SIZE": str(world),
                       "LOCAL_RANK": str(rank)})
    import torch.distributed as dist
    from deflow_b200 import dist as dd
    dd.init("gloo")
    torch.manual_seed(5)
    net = _Toy()
    g = torch.Generator().manual_seed(11)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    xs, ys = x[rank * 4:(rank + 1) * 4], y[rank * 4:(rank + 1) * 4]
    res = {}
    for mode in ("single", "early"):
        avg = dd.GradAverager(net.parameters())
        if mode == "early":
            avg.plan_early_slice(net)
            assert avg.early_from == 4                  # embedder.{w,b}, encoder.{w,b} are late; decoder + head early
            avg.arm_early_slice()
        avg.zero()
        ((net(xs) - ys) ** 2).sum().backward()
        if mode == "early":
            assert avg._early_done                      # the slice was reduced from inside the backward
        avg.average()
        res[mode] = avg.flat.clone()
        for h in avg._hooks:
            h.remove()
    out[rank] = (res["single"], res["early"])
    dd.barrier()
    dist.destroy_process_group()


def test_early_slice_allreduce_equals_single_collective():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker_early, args=(world, port, out), nprocs=world, join=True)
    for r in range(world):
        single, early = out[r]
        assert torch.equal(single, early)
    assert torch.equal(out[0][1], out[1][1])
