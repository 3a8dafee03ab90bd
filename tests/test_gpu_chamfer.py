"""Chamfer nearest-neighbour kernels (csrc/chamfer.cu, deflow_b200/chamfer3D.py) and seflowLoss:
* bit-exact (distances AND indices) against the REFERENCE's own compiled chamfer3D extension, live on the GPU, for every query
  that lies in a full 256-thread block of the reference launch (its partial last block reads target tiles the exited threads
  never staged -- undefined behaviour there);
* against the CPU oracle for ragged sizes, empty clouds, exact ties;
* the reference's own assets/cuda/chamfer3D/__init__.py on top of install_as_chamfer3D();
* seflowLoss against the oracle restatement (itself pinned against the reference function)."""
import numpy as np
import pytest
import torch

import deflow_b200 as d
from deflow_b200 import chamfer3D as ch
from oracle import build_ref, seflow_oracle as so

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _clouds(n0, n1, seed, scale=20.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n0, 3, generator=g) * scale).to(torch.float16).float(), (torch.randn(n1, 3, generator=g) * scale).to(torch.float16).float()


@pytest.mark.parametrize("n0,n1,seed", [(2048, 3072, 1), (20480, 20480, 2), (80128, 79872, 3), (256, 5000, 4)])
def test_equals_reference_extension_bit_exact(n0, n1, seed):
    ref = build_ref.load_chamfer_ref()
    if ref is None:
        pytest.skip("oracle/_ref/chamfer3D_ref.so not built")
    a, b = _clouds(n0, n1, seed)
    a, b = a.to(DEV), b.to(DEV)
    rd0, rd1 = torch.zeros(n0, device=DEV), torch.zeros(n1, device=DEV)
    ri0, ri1 = torch.zeros(n0, dtype=torch.int32, device=DEV), torch.zeros(n1, dtype=torch.int32, device=DEV)
    ref.forward(a, b, rd0, rd1, ri0, ri1)
    d0, d1, i0, i1 = ch.ChamferDis.apply(a, b)
    f0, f1 = n0 // 256 * 256, n1 // 256 * 256            # queries in full blocks of the reference launch
    assert torch.equal(d0[:f0], rd0[:f0]) and torch.equal(i0[:f0], ri0[:f0])
    assert torch.equal(d1[:f1], rd1[:f1]) and torch.equal(i1[:f1], ri1[:f1])
    # backward against the reference kernel (float atomics on both sides: tolerance)
    g0, g1 = torch.randn(n0, device=DEV), torch.randn(n1, device=DEV)
    rg0, rg1 = torch.zeros(n0, 3, device=DEV), torch.zeros(n1, 3, device=DEV)
    if f0 == n0 and f1 == n1:
        ref.backward(a, b, ri0, ri1, g0, g1, rg0, rg1)
        gp0, gp1 = torch.empty(n0, 3, device=DEV), torch.empty(n1, 3, device=DEV)
        ch.backward(a, b, i0, i1, g0, g1, gp0, gp1)
        np.testing.assert_allclose(gp0.cpu().numpy(), rg0.cpu().numpy(), rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(gp1.cpu().numpy(), rg1.cpu().numpy(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("n0,n1,seed", [(1000, 777, 5), (1, 1, 6), (5, 3000, 7), (3001, 2, 8), (0, 50, 9), (40, 0, 10)])
def test_equals_oracle_ragged_and_empty(n0, n1, seed):
    a, b = _clouds(n0, n1, seed)
    od0, od1, oi0, oi1 = so.chamfer_forward(a.numpy(), b.numpy())
    d0, d1, i0, i1 = ch.ChamferDis.apply(a.to(DEV), b.to(DEV))
    assert np.array_equal(i0.cpu().numpy(), oi0) and np.array_equal(i1.cpu().numpy(), oi1)
    assert np.array_equal(d0.cpu().numpy(), od0) and np.array_equal(d1.cpu().numpy(), od1)


def test_ties_take_the_lowest_index():
    b = torch.tensor([[1.0, 0, 0], [-1.0, 0, 0], [0, 1.0, 0], [1.0, 0, 0]]).repeat(700, 1)     # 2800 targets, 4 distinct
    a = torch.zeros(300, 3)
    d0, d1, i0, i1 = ch.ChamferDis.apply(a.to(DEV), b.to(DEV))
    assert bool((i0 == 0).all()) and bool((d0 == 1.0).all()) and bool((i1 == 0).all())


def test_reference_python_wrapper_on_the_dropin_module_and_autograd():
    """The reference's own assets/cuda/chamfer3D/__init__.py bound to install_as_chamfer3D(), and the gradient of
    nnChamferDis against the oracle."""
    import importlib, sys
    from oracle import ref_modules
    a, b = _clouds(3000, 2500, 11, 5.0)
    pa, pb = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref_loss = so.ChamferStandIn()(pa, pb, truncate_dist=4)
    ref_loss.backward()
    ga, gb = a.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    loss = ch.nnChamferDis()(ga, gb, truncate_dist=4)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    np.testing.assert_allclose(ga.grad.cpu().numpy(), pa.grad.numpy(), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(gb.grad.cpu().numpy(), pb.grad.numpy(), rtol=1e-4, atol=1e-6)
    root = ref_modules.root()
    if root is None:
        pytest.skip("reference modules not staged")
    ch.install_as_chamfer3D()
    for k in [k for k in sys.modules if k.startswith("assets")]:
        del sys.modules[k]
    if root not in sys.path:
        sys.path.insert(0, root)
    refmod = importlib.import_module("assets.cuda.chamfer3D")
    l2 = refmod.nnChamferDis(truncate_dist=False)(a.to(DEV), b.to(DEV), truncate_dist=4)
    assert abs(float(l2) - float(loss)) <= 1e-6 * abs(float(loss))


@pytest.mark.parametrize("n0,n1,seed", [(6000, 5600, 1), (900, 1000, 2), (400, 300, 3)])
def test_seflow_loss_equals_oracle(n0, n1, seed):
    sc = so.make_scene(n0, n1, seed)
    e1 = sc["est_flow"].clone().requires_grad_(True)
    want = so.seflow_loss({**sc, "est_flow": e1})
    sum(want.values()).backward()
    g = {k: v.to(DEV) for k, v in sc.items()}
    e2 = g["est_flow"].clone().requires_grad_(True)
    got = d.seflowLoss({**g, "est_flow": e2})
    assert set(got) == set(want)
    for k in want:
        assert abs(float(got[k]) - float(want[k])) <= 1e-5 * max(1.0, abs(float(want[k]))), k
    sum(got.values()).backward()
    np.testing.assert_allclose(e2.grad.cpu().numpy(), e1.grad.numpy(), rtol=1e-4, atol=1e-7)
