"""Device-side collate_fn_pad / ground strip (deflow_b200/feed.py, csrc/collate.cu) against the oracle restatement of
OSF/src/dataset.py:22-74 -- bit-exact (stable compaction, NaN / zero padding), ragged and degenerate batches."""
import pytest
import torch

import deflow_b200 as d
from deflow_b200.feed import DeviceCollator, DeviceFeeder
from oracle import feed_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _same(dev_batch, ref):
    assert set(dev_batch) == set(ref)
    for k, r in ref.items():
        g = dev_batch[k]
        if isinstance(r, list):
            assert len(g) == len(r) and all(torch.equal(x.cpu(), y.float()) for x, y in zip(g, r)), k
            continue
        assert g.dtype == r.dtype and tuple(g.shape) == tuple(r.shape), (k, g.dtype, r.dtype, g.shape, r.shape)
        if r.is_floating_point():
            assert torch.equal(torch.isnan(g.cpu()), torch.isnan(r)), k
            assert torch.equal(g.cpu().nan_to_num(0.0), r.nan_to_num(0.0)), k
        else:
            assert torch.equal(g.cpu(), r), k


@pytest.mark.parametrize("B,lo,hi,flow,seed", [(4, 2000, 9000, True, 1), (3, 1, 40, True, 2), (1, 5000, 5000, False, 3),
                                                (5, 1000, 1100, True, 4), (2, 100000, 120000, True, 5)])
def test_device_collate_equals_reference_collate(B, lo, hi, flow, seed):
    samples = feed_oracle.make_samples(B, lo, hi, seed, flow)
    col = DeviceCollator(DEV)
    _same(col(samples), feed_oracle.collate_fn_pad(samples))
    _same(col(samples), feed_oracle.collate_fn_pad(samples))       # staging buffers are reused


def test_device_collate_degenerate_masks():
    samples = feed_oracle.make_samples(3, 300, 2500, 7, True)
    samples[0]["gm0"][:] = True          # every point of pc0 is ground: the row is all padding
    samples[1]["gm1"][:] = False         # nothing dropped
    samples[2]["gm0"][:] = False
    col = DeviceCollator(DEV)
    _same(col(samples), feed_oracle.collate_fn_pad(samples))
    for s in samples:                    # a whole batch without kept points: Nmax = 0
        s["gm0"][:] = True
    _same(col(samples), feed_oracle.collate_fn_pad(samples))


def test_strip_ground_and_model_consumes_collated_batch():
    """run_model_wo_ground_data (OSF/src/trainer.py:268-282): pc[~gm].unsqueeze(0); and the collated batch feeds
    DeFlow.forward + the fused loss directly."""
    samples = feed_oracle.make_samples(2, 3000, 4000, 11, True)
    for s in samples:
        s["pc0"][:, :2] *= 0.3; s["pc1"][:, :2] *= 0.3
        s["pose0"], s["pose1"] = torch.eye(4), torch.eye(4)
    col = DeviceCollator(DEV)
    one = col.strip_ground(samples[0])
    assert torch.equal(one["pc0"][0].cpu(), samples[0]["pc0"][~samples[0]["gm0"]])
    assert torch.equal(one["pc1"][0].cpu(), samples[0]["pc1"][~samples[0]["gm1"]])
    assert torch.equal(one["origin_pc0"].cpu(), samples[0]["pc0"]) and torch.equal(one["gm0"].cpu(), samples[0]["gm0"])
    feeder = DeviceFeeder(DEV)
    feeder.submit_samples(samples)
    batch = feeder.get()
    m = d.DeFlow([0.2, 0.2, 6], [-6.4, -6.4, -3, 6.4, 6.4, 3], [64, 64], "gru", 4, precision="fp32").to(DEV).train()
    res = m(batch)
    loss = d.training_step_loss(batch, res, "deflowLoss")
    loss.backward()
    assert torch.isfinite(loss)
    ref = feed_oracle.collate_fn_pad(samples)
    host = {k: ([t.to(DEV) for t in v] if isinstance(v, list) else v.to(DEV)) for k, v in ref.items()}
    m.zero_grad()
    res2 = m(host)
    for a, b in zip(res["pc0_valid_point_idxes"], res2["pc0_valid_point_idxes"]):
        assert torch.equal(a, b)
