"""Input feed: everything between the dataset's per-sweep records and ``DeFlow.forward``.

* ``sample_from_h5`` -- the reference's HDF5 record layout (OSF/src/dataset.py:131-205: ``lidar f32[N,3+]``,
  ``ground_mask bool[N]``, ``pose f32[4,4]``, optional ``flow``, ``flow_is_valid``, ``flow_category_indices``,
  ``ego_motion``, ``eval_mask``) -> the sample dict ``HDF5Dataset.__getitem__`` returns.  Works on any mapping of
  array-likes (an ``h5py.Group`` of a real file, or a dict of numpy arrays).
* ``DeviceCollator`` -- ``collate_fn_pad`` (OSF/src/dataset.py:22-74) on the GPU: the raw sweeps of a batch are
  concatenated into pinned staging buffers (one H2D copy per field), and the ground-mask strip, the NaN padding of the
  points and the zero padding of flow / validity / class run as one launch sequence (csrc/collate.cu) instead of
  boolean indexing + pad_sequence in the DataLoader workers.  ``strip_ground`` is the single-sample form used by
  validation (``ModelWrapper.run_model_wo_ground_data``, OSF/src/trainer.py:268-282).
* ``DeviceFeeder`` -- double buffering: the copy + collate of step i+1 runs on a side stream under the compute of step i.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check
from .synth import batch_to


def _t(a, dtype=None):
    t = a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(np.asarray(a)))
    return t if dtype is None else t.to(dtype)


def sample_from_h5(cur, nxt, scene_id="", timestamp="", eval_index: bool = False) -> Dict:
    """One training / validation sample from two consecutive sweep records (OSF/src/dataset.py:131-205)."""
    s = {"scene_id": scene_id, "timestamp": str(timestamp),
         "pc0": _t(cur["lidar"][:])[:, :3], "gm0": _t(cur["ground_mask"][:]), "pose0": _t(cur["pose"][:]),
         "pc1": _t(nxt["lidar"][:])[:, :3], "gm1": _t(nxt["ground_mask"][:]), "pose1": _t(nxt["pose"][:])}
    if "flow" in cur:
        s["flow"] = _t(cur["flow"][:])
        s["flow_is_valid"] = _t(cur["flow_is_valid"][:])
        s["flow_category_indices"] = _t(cur["flow_category_indices"][:])
    if "ego_motion" in cur:
        s["ego_motion"] = _t(cur["ego_motion"][:])
    if eval_index:
        s["eval_mask"] = _t(cur["eval_mask"][:]) if "eval_mask" in cur else torch.ones(s["pc0"].shape[0], dtype=torch.bool)
    return s


class _Stage:
    """Pinned host staging buffer + device twin, grown geometrically."""

    def __init__(self, dtype, width, device):
        self.dtype, self.width, self.device = dtype, width, device
        self.host = self.dev = self.copied = None

    def fill(self, parts: Sequence[torch.Tensor], total: int):
        if self.copied is not None:
            self.copied.synchronize()           # the previous async copy out of the pinned buffer has finished
        if self.host is None or self.host.shape[0] < total:
            cap = max(int(total * 1.25), 1024)
            shape = (cap, self.width) if self.width else (cap,)
            self.host = torch.empty(shape, dtype=self.dtype).pin_memory()
            self.dev = torch.empty(shape, dtype=self.dtype, device=self.device)
        off = 0
        for p in parts:
            n = p.shape[0]
            self.host[off:off + n].copy_(p)     # converts dtype (bool -> u8, f16/f64 -> f32) while staging
            off += n
        self.dev[:total].copy_(self.host[:total], non_blocking=True)
        self.copied = torch.cuda.Event()
        self.copied.record(torch.cuda.current_stream(self.device))
        return self.dev


class DeviceCollator:
    """``collate_fn_pad`` on the device.  ``__call__(samples) -> batch`` with the reference's keys and layouts:
    ``pc0, pc1 f32[B,Nmax,3]`` NaN-padded, ``pose0, pose1`` lists of ``[4,4]``, ``flow f32[B,Nmax0,3]``,
    ``flow_is_valid bool``, ``flow_category_indices u8`` zero-padded, optional ``ego_motion``."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("deflow_b200.feed.DeviceCollator runs on CUDA only; there is no CPU path")
        mk = lambda dt, w: _Stage(dt, w, self.device)  # noqa: E731
        self.st = {k: mk(torch.float32, 3) for k in ("pc0", "pc1", "flow")}
        self.st.update({k: mk(torch.uint8, 0) for k in ("gm0", "gm1", "valid", "cls")})

    def _frame(self, pcs, gms, flow=None, valid=None, cls=None, tag="0"):
        lib = _lib.lib()
        dev = self.device
        B = len(pcs)
        ns = [int(p.shape[0]) for p in pcs]
        total, max_n = sum(ns), max(ns + [0])
        keep = [int(n - int(g.sum())) for n, g in zip(ns, gms)]    # host count: the masks are host-resident anyway
        Nmax = max(keep + [0])
        st = torch.cuda.current_stream(dev).cuda_stream
        pts = self.st["pc" + tag].fill([p[:, :3] for p in pcs], total)
        gm = self.st["gm" + tag].fill(gms, total)
        offs = torch.tensor(np.concatenate([[0], np.cumsum(ns)]).astype(np.int32)).pin_memory().to(dev, non_blocking=True)
        out = {"pc": torch.empty((B, Nmax, 3), dtype=torch.float32, device=dev)}
        f = v = c = None
        if flow is not None:
            f = self.st["flow"].fill(flow, total)
            v = self.st["valid"].fill(valid, total)
            c = self.st["cls"].fill(cls, total)
            out["flow"] = torch.empty((B, Nmax, 3), dtype=torch.float32, device=dev)
            out["valid"] = torch.empty((B, Nmax), dtype=torch.uint8, device=dev)
            out["cls"] = torch.empty((B, Nmax), dtype=torch.uint8, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        if Nmax == 0 or total == 0:          # nothing survives the ground mask: an all-padding batch of width 0
            counts.zero_()
            out["counts"], out["keep_host"] = counts, keep
            return out
        ws = torch.empty(int(lib.dfb_collate_workspace(B, max_n)), dtype=torch.int32, device=dev)
        P = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        check(lib.dfb_collate_pad(pts.data_ptr(), gm.data_ptr(), offs.data_ptr(), B, max_n, Nmax, P(f), P(v), P(c),
                                  out["pc"].data_ptr(), P(out.get("flow")), P(out.get("valid")), P(out.get("cls")),
                                  counts.data_ptr(), ws.data_ptr(), st), "collate_pad")
        out["counts"], out["keep_host"] = counts, keep
        return out

    def __call__(self, samples: List[Dict]) -> Dict:
        dev = self.device
        has_flow = "flow" in samples[0]
        gm0 = [_t(s["gm0"]).to(torch.bool) for s in samples]
        gm1 = [_t(s["gm1"]).to(torch.bool) for s in samples]
        a = self._frame([_t(s["pc0"]) for s in samples], gm0,
                        [_t(s["flow"]) for s in samples] if has_flow else None,
                        [_t(s["flow_is_valid"]) for s in samples] if has_flow else None,
                        [_t(s["flow_category_indices"]) for s in samples] if has_flow else None, "0")
        b = self._frame([_t(s["pc1"]) for s in samples], gm1, tag="1")
        pose0 = torch.stack([_t(s["pose0"], torch.float32) for s in samples]).pin_memory().to(dev, non_blocking=True)
        pose1 = torch.stack([_t(s["pose1"], torch.float32) for s in samples]).pin_memory().to(dev, non_blocking=True)
        batch = {"pc0": a["pc"], "pc1": b["pc"], "pose0": list(pose0.unbind(0)), "pose1": list(pose1.unbind(0))}
        if has_flow:
            batch["flow"] = a["flow"]
            batch["flow_is_valid"] = a["valid"].view(torch.bool)
            batch["flow_category_indices"] = a["cls"]
        if "ego_motion" in samples[0]:
            ego = torch.stack([_t(s["ego_motion"], torch.float32) for s in samples]).pin_memory().to(dev, non_blocking=True)
            batch["ego_motion"] = list(ego.unbind(0))
        return batch

    def strip_ground(self, sample: Dict) -> Dict:
        """``run_model_wo_ground_data`` (OSF/src/trainer.py:268-282): keeps ``origin_pc0`` and replaces pc0 / pc1 by
        ``pc[~gm].unsqueeze(0)`` -- the B = 1 case of the same kernels; gm0 / gm1 go to the device for the scatter-back."""
        dev = self.device
        out = self([sample])
        out["origin_pc0"] = _t(sample["pc0"], torch.float32)[:, :3].to(dev, non_blocking=True)
        out["gm0"] = _t(sample["gm0"]).to(torch.bool).to(dev, non_blocking=True)
        out["gm1"] = _t(sample["gm1"]).to(torch.bool).to(dev, non_blocking=True)
        for k in ("eval_mask", "scene_id", "timestamp"):
            if k in sample:
                out[k] = sample[k].to(dev) if torch.is_tensor(sample[k]) else sample[k]
        return out


class DeviceFeeder:
    """Double buffering: ``submit`` starts the H2D copy of a pinned, already collated host batch -- or, with
    ``submit_samples``, the copy + device-side collate of raw samples -- on a side stream; ``get`` hands the result to
    the current stream.  The copy of step i+1 overlaps the compute of step i."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None
        self._collator: Optional[DeviceCollator] = None

    def submit(self, host_batch):
        """Start the H2D copy of a (pinned) host batch."""
        with torch.cuda.stream(self.stream):
            dev = batch_to(host_batch, self.device, non_blocking=True)
            evt = torch.cuda.Event()
            evt.record(self.stream)
        self._pending = (dev, evt)

    def submit_samples(self, samples: List[Dict]):
        """Start copy + ground strip + padding of raw samples (``HDF5Dataset.__getitem__`` records)."""
        if self._collator is None:
            self._collator = DeviceCollator(self.device)
        with torch.cuda.stream(self.stream):
            dev = self._collator(samples)
            evt = torch.cuda.Event()
            evt.record(self.stream)
        self._pending = (dev, evt)

    def get(self):
        """The batch submitted last, ready for the current stream."""
        dev, evt = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(evt)
        for v in dev.values():
            for t in (v if isinstance(v, list) else [v]):
                if torch.is_tensor(t):
                    t.record_stream(cur)
        return dev
