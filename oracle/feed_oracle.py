"""CPU restatement of the reference's batch collation.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

collate_fn_pad (OpenSceneFlow/src/dataset.py:22-74, two-frame case): per sample ``pc[~gm]``, then
``torch.nn.utils.rnn.pad_sequence(batch_first=True)`` with NaN for the points and the default 0 for flow,
flow_is_valid and flow_category_indices (all three follow the pc0 ground mask).  Pinned against the reference's own
function (AST-extracted from dataset.py, which imports h5py at module level) in tests/test_feed_oracle.py."""
import torch


def collate_fn_pad(batch):
    pad = torch.nn.utils.rnn.pad_sequence
    keep0 = [~b["gm0"] for b in batch]          # dataset.py:32
    keep1 = [~b["gm1"] for b in batch]          # dataset.py:33
    res = {
        "pc0": pad([b["pc0"][k] for b, k in zip(batch, keep0)], batch_first=True, padding_value=float("nan")),   # :37
        "pc1": pad([b["pc1"][k] for b, k in zip(batch, keep1)], batch_first=True, padding_value=float("nan")),   # :38
        "pose0": [b["pose0"] for b in batch],   # :45
        "pose1": [b["pose1"] for b in batch],   # :46
    }
    if "flow" in batch[0]:                      # :53-59
        res["flow"] = pad([b["flow"][k] for b, k in zip(batch, keep0)], batch_first=True)
        res["flow_is_valid"] = pad([b["flow_is_valid"][k] for b, k in zip(batch, keep0)], batch_first=True)
        res["flow_category_indices"] = pad([b["flow_category_indices"][k] for b, k in zip(batch, keep0)], batch_first=True)
    if "ego_motion" in batch[0]:                # :61-62
        res["ego_motion"] = [b["ego_motion"] for b in batch]
    return res


def make_samples(B, n_lo, n_hi, seed, with_flow=True, ground_frac=0.3):
    """Ragged raw samples in the layout HDF5Dataset.__getitem__ returns (dataset.py:131-205)."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(B):
        n0 = int(torch.randint(n_lo, n_hi + 1, (1,), generator=g))
        n1 = int(torch.randint(n_lo, n_hi + 1, (1,), generator=g))
        s = {"scene_id": f"scene{b}", "timestamp": str(1000 + b),
             "pc0": torch.randn(n0, 3, generator=g) * 20, "gm0": torch.rand(n0, generator=g) < ground_frac,
             "pose0": torch.eye(4) + 0.01 * torch.randn(4, 4, generator=g),
             "pc1": torch.randn(n1, 3, generator=g) * 20, "gm1": torch.rand(n1, generator=g) < ground_frac,
             "pose1": torch.eye(4) + 0.01 * torch.randn(4, 4, generator=g)}
        if with_flow:
            s["flow"] = torch.randn(n0, 3, generator=g)
            s["flow_is_valid"] = torch.rand(n0, generator=g) < 0.9
            s["flow_category_indices"] = torch.randint(0, 31, (n0,), generator=g, dtype=torch.uint8)
        out.append(s)
    return out
