"""Input feed: pinned host batches (``collate_fn_pad`` layout, OSF/src/dataset.py:22-74) copied to the device on a
side stream, double-buffered so the copy of step i+1 overlaps the compute of step i."""
from __future__ import annotations

import torch

from .synth import batch_to


class DeviceFeeder:
    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def submit(self, host_batch):
        """Start the H2D copy of a (pinned) host batch."""
        with torch.cuda.stream(self.stream):
            dev = batch_to(host_batch, self.device, non_blocking=True)
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
                t.record_stream(cur)
        return dev
