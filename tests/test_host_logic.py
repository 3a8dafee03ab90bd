"""Host-side logic that needs no GPU: parameter tree / state_dict contract, init, synthetic generator."""
import numpy as np
import torch

import deflow_b200 as d
from deflow_b200 import synth
from oracle import deflow_oracle as orc


def test_state_dict_matches_reference_layout():
    for dec in ("gru", "linear"):
        sd = d.DeFlow(decoder_option=dec).state_dict()
        ref = orc.param_shapes(dec)
        assert set(sd) == set(ref)
        for k, shp in ref.items():
            assert tuple(sd[k].shape) == tuple(shp), k
    assert len(d.DeFlow().state_dict()) == 156  # SURVEY.md section 5, checkpoint row
    assert sum(p.numel() for p in d.DeFlow().parameters()) == 6891939


def test_load_reference_style_state_dict():
    m = d.DeFlow()
    state = orc.random_state(3, "gru")
    m.load_state_dict(state, strict=True)
    assert torch.equal(m.head.gru.convz.weight, state["head.gru.convz.weight"])


def test_weights_init_rules():
    m = d.DeFlow()
    m.apply(d.weights_init)
    assert float(m.backbone.decoder_step4.bias.abs().max()) == 0.0
    assert float(m.head.offset_encoder.bias.abs().max()) == 0.0
    w = m.backbone.encoder_step_1[0].conv.weight
    bound = (6.0 / (32 * 9 + 64 * 9)) ** 0.5
    assert float(w.abs().max()) <= bound + 1e-6
    # Conv1d keeps torch's default init: bias is not zeroed (mics.py:98-105 only touches Conv2d / Linear)
    assert float(m.head.gru.convz.bias.abs().max()) > 0.0


def test_synthetic_batch_layout():
    b = synth.make_batch(2, 3000, seed=5)
    assert b["pc0"].shape[0] == 2 and b["pc0"].shape[2] == 3 and b["pc0"].dtype == torch.float32
    assert b["flow"].shape == b["pc0"].shape and b["flow_category_indices"].dtype == torch.uint8
    pc = b["pc0"][0]
    ok = ~torch.isnan(pc).any(1)
    assert torch.equal(pc[ok], pc[ok].half().float())  # fp16-exact like AV2
    # occupancy statistics in the AV2 range: 0.03 .. 0.5 pillars per valid point, some points out of range
    info = orc.voxelize_frame(pc)
    from oracle import mmcv_ext_oracle as ext
    m = ext.unique_pillars(info["voxel_coords"].numpy())[0].shape[0]
    n = info["points"].shape[0]
    assert 0.03 < m / n < 0.5 and n < int(ok.sum())


def test_zero_pool_hands_out_disjoint_zeroed_views():
    from deflow_b200 import conv
    ts = [conv.zeros(s, dt, "cpu") for s, dt in [((2, 64), torch.float64), ((3, 32), torch.float32), ((5,), torch.float32),
                                                 ((2, 256), torch.float64), ((1 << 19,), torch.float32)]]
    for i, t in enumerate(ts):
        assert float(t.abs().sum()) == 0.0
        t += i + 1                       # writing one view must not touch the others
    for i, t in enumerate(ts):
        assert float(t.min()) == float(t.max()) == i + 1
    assert ts[0].dtype == torch.float64 and ts[0].shape == (2, 64) and ts[0].data_ptr() % 8 == 0


def test_batch_counters_flush_once_per_forward():
    from deflow_b200 import conv
    a, b = torch.zeros((), dtype=torch.long), torch.zeros((), dtype=torch.long)
    conv._pending_nbt.extend([a, b, a])   # an encoder BatchNorm is called twice per forward (pc0, pc1), a decoder-side one once
    conv.flush_batch_counters()
    assert int(a) == 2 and int(b) == 1 and not conv._pending_nbt
    conv.flush_batch_counters()
    assert int(a) == 2 and int(b) == 1
