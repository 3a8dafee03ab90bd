"""Input-feed oracle (oracle/feed_oracle.py) against the reference's own collate_fn_pad, and the host-side pieces of
deflow_b200.feed.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import feed_oracle, ref_modules


def _same(a, b):
    assert a.keys() == b.keys()
    for k in a:
        if isinstance(a[k], list):
            assert all(torch.equal(x, y) for x, y in zip(a[k], b[k])), k
        else:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            if a[k].is_floating_point():
                assert torch.equal(a[k].nan_to_num(1e9), b[k].nan_to_num(1e9)), k
            else:
                assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("with_flow", [True, False])
def test_collate_oracle_equals_reference_function(with_flow):
    if ref_modules.root() is None:
        pytest.skip("reference sources not available (neither /root/reference nor oracle/_ref/osf)")
    ref = ref_modules.extract_functions("src/dataset.py", ("collate_fn_pad",))["collate_fn_pad"]
    samples = feed_oracle.make_samples(3, 50, 400, 1, with_flow)
    _same(feed_oracle.collate_fn_pad(samples), ref(samples))
    # degenerate: a sample whose points are all ground
    samples[1]["gm0"][:] = True
    _same(feed_oracle.collate_fn_pad(samples), ref(samples))


def test_sample_from_h5_layout():
    """The HDF5 record layout of the reference (OSF/src/dataset.py:131-205) as a mapping of arrays."""
    from deflow_b200.feed import sample_from_h5
    rng = np.random.default_rng(0)
    cur = {"lidar": rng.normal(size=(100, 4)).astype(np.float32), "ground_mask": rng.random(100) < 0.3,
           "pose": np.eye(4, dtype=np.float32), "flow": rng.normal(size=(100, 3)).astype(np.float32),
           "flow_is_valid": rng.random(100) < 0.9, "flow_category_indices": rng.integers(0, 31, 100).astype(np.uint8),
           "ego_motion": np.eye(4, dtype=np.float32)}
    nxt = {"lidar": rng.normal(size=(90, 4)).astype(np.float32), "ground_mask": rng.random(90) < 0.3,
           "pose": np.eye(4, dtype=np.float32)}
    s = sample_from_h5(cur, nxt, "sc", 123, eval_index=True)
    assert s["pc0"].shape == (100, 3) and s["pc1"].shape == (90, 3) and s["gm0"].dtype == torch.bool
    assert s["timestamp"] == "123" and s["eval_mask"].all() and s["flow_category_indices"].dtype == torch.uint8
    assert set(s) >= {"pc0", "gm0", "pose0", "pc1", "gm1", "pose1", "flow", "flow_is_valid", "flow_category_indices", "ego_motion"}
