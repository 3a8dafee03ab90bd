"""The REFERENCE's own Python modules (staged copy oracle/_ref/osf or /root/reference; oracle/ref_modules.py) running on
the product's plugin boundary on the GPU:

* the reference's assets/cuda/mmcv/{scatter_points,voxelize}.py bound to ``deflow_b200.mmcv_ext.install_as_mmcv_ext()``
  -- the drop-in exactly as INTEGRATION.md tells a maintainer to install it -- against the same wrappers bound to the
  reference's own CUDA extension;
* the reference DeFlow (cuDNN/cuBLAS fp32) on the reference extension against deflow_b200.DeFlow in parity mode, same
  weights, same batch: flows within the north-star bound, identical indices.
"""
import numpy as np
import pytest
import torch

import deflow_b200 as d
from deflow_b200 import synth
from oracle import deflow_oracle as orc
from oracle import ref_modules
from helpers import load_fixture, batch_to

pytestmark = pytest.mark.gpu
DEV = "cuda"
VS, RG = [0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]


def _need_ref():
    if ref_modules.root() is None:
        pytest.skip("reference modules not staged (python oracle/build_ref.py in the build container)")
    from oracle import build_ref
    if build_ref.load_ref() is None:
        pytest.skip("oracle/_ref/mmcv_ref_ext.so not built")


def _scatter_through(mmcv, pts, red_mean=True):
    vox = mmcv.Voxelization(VS, RG, max_num_points=-1)
    coors = vox(pts)
    ok = (coors != -1).all(1)
    coors, p = coors[ok].contiguous(), pts[ok].contiguous().requires_grad_(True)
    sc = mmcv.DynamicScatter(VS, RG, red_mean)
    feats, vc = sc(p, coors)
    feats.square().sum().backward()
    return coors, feats.detach(), vc, p.grad


@pytest.mark.parametrize("mean", [True, False])
def test_reference_mmcv_wrappers_on_the_dropin_ext(mean):
    _need_ref()
    pts = synth.make_batch(1, 50000, seed=9)["pc0"][0]
    pts = pts[~torch.isnan(pts).any(1)].to(DEV).contiguous()
    out = {}
    for ext in ("cuda", "dfb"):
        mmcv = ref_modules.load_mmcv_wrappers(ext)        # the reference's own scatter_points.py / voxelize.py
        assert mmcv.scatter_points.ext_module is __import__("sys").modules["mmcv._ext"]
        out[ext] = _scatter_through(mmcv, pts, mean)
    a, b = out["cuda"], out["dfb"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2])                       # voxel coords, unique pillars: bit-exact
    np.testing.assert_allclose(b[1].cpu().numpy(), a[1].cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(b[3].cpu().numpy(), a[3].cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_dropin_ext_error_contract():
    """scatter_points.cpp:32 / pytorch_device_registry.hpp:116-122: unknown reduce type and CPU tensors raise RuntimeError."""
    from deflow_b200 import mmcv_ext
    e = mmcv_ext.install_as_mmcv_ext()
    f, c = torch.zeros(4, 3, device=DEV), torch.zeros(4, 3, dtype=torch.int32, device=DEV)
    with pytest.raises(RuntimeError):
        e.dynamic_point_to_voxel_forward(f, c, "median")
    with pytest.raises(RuntimeError):
        e.dynamic_point_to_voxel_forward(f.cpu(), c.cpu(), "mean")
    with pytest.raises(RuntimeError):
        e.hard_voxelize_forward(f.cpu(), torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.zeros(1), torch.zeros(1), 5, 10, 3, True)


@pytest.mark.parametrize("ext", ["cuda", "dfb"])
def test_reference_deflow_on_gpu_vs_parity_mode(ext):
    """The reference model itself (strict fp32: TF32 off) on the GPU -- with its own CUDA extension, and with the product's
    drop-in extension under it -- against deflow_b200.DeFlow(precision='fp32')."""
    _need_ref()
    DeFlow, _, lossns = ref_modules.load_reference(ext)
    batch = synth.make_batch(2, 20000, seed=12)
    state = orc.random_state(41, "gru")
    gb = batch_to(batch, DEV)
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = DeFlow(VS, RG, [512, 512], "gru", 4)
        ref.load_state_dict(state, strict=True)
        ref = ref.to(DEV).train()
        r = ref(gb)
        rl = 0.0
        for b in range(2):
            i = r["pc0_valid_point_idxes"][b]
            rl = rl + lossns["deflowLoss"]({"est_flow": r["flow"][b], "gt_flow": gb["flow"][b][i] - r["pose_flow"][b][i]})["loss"]
        rl.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    m = d.DeFlow(VS, RG, [512, 512], "gru", 4, precision="fp32")
    m.load_state_dict(orc.random_state(41, "gru"), strict=True)
    m = m.to(DEV).train()
    o = m(gb)
    ol = d.training_step_loss(gb, o, "deflowLoss")
    ol.backward()
    for b in range(2):
        assert torch.equal(o["pc0_valid_point_idxes"][b], r["pc0_valid_point_idxes"][b])
        assert torch.equal(o["pc1_valid_point_idxes"][b], r["pc1_valid_point_idxes"][b])
        err = float((o["flow"][b] - r["flow"][b]).abs().max())
        assert err <= 1e-3, err
    assert abs(float(ol) - float(rl)) <= 2e-4 * max(1.0, abs(float(rl)))
    rg = dict(ref.named_parameters())
    for k in ("backbone.encoder_step_1.0.conv.weight", "backbone.decoder_step4.weight", "head.gru.convz.weight",
              "head.gru.convq.weight", "backbone.decoder_step2.u4_u5.0.weight", "embedder.feature_net.pfn_layers.0.0.weight"):
        g, w = dict(m.named_parameters())[k].grad, rg[k].grad
        assert float((g - w).norm()) <= 2e-3 * float(w.norm()), k
