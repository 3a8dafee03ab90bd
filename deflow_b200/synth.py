"""Synthetic Argoverse-2-shaped frame pairs ("Synthetic input generator").

There is no dataset on the build or GPU boxes, so benches and tests draw scenes here.  The
shape statistics follow the one real sweep the reference ships (OpenSceneFlow/assets/tests/
test_pc0.npy: ~20 % of points outside +-51.2 m, ~0.07 pillars per point at 0.2 m, median 4 /
max ~750 points per pillar, coordinates exactly representable in fp16):

* 2-D "column sites" with 1/r radial density on [2, 75] m, uniform azimuth;
* a heavy-tailed (log-normal) number of points per site, xy jitter inside the column, z
  spread along the column;
* coordinates rounded to fp16-representable values;
* pc1 = independent re-sample of the same sites after a small SE(3) (yaw 0.01 rad, 1 m in x)
  plus per-site velocities; gt flow = site velocity * 0.1 s (+ ego motion), classes u8;
* batches are NaN-padded to a common Nmax exactly like ``collate_fn_pad``
  (OpenSceneFlow/src/dataset.py:22-74).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

SEED_BASE = 42069  # OpenSceneFlow/conf/config.yaml:34


def _round_fp16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float16).to(torch.float32)


def make_scene(n_points: int, gen: torch.Generator, sites_per_point: float = 0.085,
               r_min: float = 2.0, r_max: float = 100.0, jitter: float = 0.1):
    """Draw column sites and a per-point site assignment for one frame pair."""
    n_sites = max(8, int(n_points * sites_per_point))
    # 1/r density on [r_min, r_max]  <=>  log-uniform radius
    u = torch.rand(n_sites, generator=gen)
    r = r_min * (r_max / r_min) ** u
    th = torch.rand(n_sites, generator=gen) * (2 * math.pi)
    site_xy = torch.stack([r * torch.cos(th), r * torch.sin(th)], dim=1)
    # heavy-tailed weights: log-normal, median ~4 points, tail to several hundred
    w = torch.exp(torch.randn(n_sites, generator=gen) * 1.5)
    site_z0 = torch.randn(n_sites, generator=gen) * 0.6 + 0.2
    site_h = torch.rand(n_sites, generator=gen) * 2.5 + 0.2
    # ~12 % of the sites move (vehicles / pedestrians), up to ~20 m/s
    moving = torch.rand(n_sites, generator=gen) < 0.12
    speed = torch.rand(n_sites, generator=gen) * 20.0
    heading = torch.rand(n_sites, generator=gen) * (2 * math.pi)
    vel = torch.stack([speed * torch.cos(heading), speed * torch.sin(heading), torch.zeros(n_sites)], dim=1)
    vel = vel * moving.unsqueeze(1)
    cls = (moving.to(torch.uint8) * (1 + (torch.rand(n_sites, generator=gen) * 20).to(torch.uint8)))
    return {"xy": site_xy, "w": w, "z0": site_z0, "h": site_h, "vel": vel, "cls": cls, "jitter": jitter}


def sample_frame(scene: dict, n_points: int, gen: torch.Generator, dt: float = 0.0,
                 pose: Optional[torch.Tensor] = None):
    """Sample ``n_points`` returns from the scene at time ``dt`` seen from ``pose`` (sensor<-world is
    inv(pose)).  Returns (points[n,3] fp32 fp16-exact, site index[n])."""
    sid = torch.multinomial(scene["w"], n_points, replacement=True, generator=gen)
    j = scene["jitter"]
    xy = scene["xy"][sid] + (torch.rand(n_points, 2, generator=gen) * 2 - 1) * j
    z = scene["z0"][sid] + torch.rand(n_points, generator=gen) * scene["h"][sid]
    # a few % of returns far above / below the pillar layer to exercise the z range test
    far = torch.rand(n_points, generator=gen) < 0.03
    z = torch.where(far, z + 6.0 * torch.sign(torch.randn(n_points, generator=gen)), z)
    p = torch.cat([xy, z.unsqueeze(1)], dim=1) + scene["vel"][sid] * dt
    if pose is not None:
        rot, t = pose[:3, :3], pose[:3, 3]
        p = (p - t) @ rot  # world -> sensor: R^T (p - t)
    return _round_fp16(p), sid


def make_pair(n_points: int, seed: int, yaw: float = 0.01, tx: float = 1.0) -> Dict[str, torch.Tensor]:
    """One frame pair with ground-truth flow in the pc0 sensor frame convention of the dataset:
    ``flow = p(t+dt in sensor-1 frame) - p(t in sensor-0 frame)`` which contains ego motion;
    the trainer subtracts pose_flow again (OpenSceneFlow/src/trainer.py:123-126)."""
    gen = torch.Generator().manual_seed(seed)
    scene = make_scene(n_points, gen)
    pose0 = torch.eye(4)
    pose1 = torch.eye(4)
    c, s = math.cos(yaw), math.sin(yaw)
    pose1[:3, :3] = torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
    pose1[0, 3] = tx
    n0 = n_points
    n1 = max(1, int(n_points * (1.0 + 0.005 * (2 * torch.rand((), generator=gen).item() - 1))))
    pc0, sid0 = sample_frame(scene, n0, gen, 0.0, pose0)
    pc1, _ = sample_frame(scene, n1, gen, 0.1, pose1)
    # gt: where each pc0 point is at t+0.1 s, expressed in sensor-1 frame
    moved = pc0 + scene["vel"][sid0] * 0.1
    in1 = (moved - pose1[:3, 3]) @ pose1[:3, :3]
    flow = in1 - pc0
    return {"pc0": pc0, "pc1": pc1, "pose0": pose0, "pose1": pose1, "flow": flow,
            "flow_is_valid": torch.ones(n0, dtype=torch.bool),
            "flow_category_indices": scene["cls"][sid0]}


def make_batch(batch_size: int, n_points: int, seed: int = SEED_BASE, device: str = "cpu",
               pin: bool = False) -> Dict:
    """NaN-padded batch in the ``collate_fn_pad`` layout."""
    pairs = [make_pair(n_points, seed * 1000 + b) for b in range(batch_size)]
    pad = torch.nn.utils.rnn.pad_sequence
    batch = {
        "pc0": pad([p["pc0"] for p in pairs], batch_first=True, padding_value=float("nan")),
        "pc1": pad([p["pc1"] for p in pairs], batch_first=True, padding_value=float("nan")),
        "pose0": [p["pose0"] for p in pairs],
        "pose1": [p["pose1"] for p in pairs],
        "flow": pad([p["flow"] for p in pairs], batch_first=True),
        "flow_is_valid": pad([p["flow_is_valid"] for p in pairs], batch_first=True),
        "flow_category_indices": pad([p["flow_category_indices"] for p in pairs], batch_first=True),
    }
    if pin:
        for k, v in batch.items():
            batch[k] = [t.pin_memory() for t in v] if isinstance(v, list) else v.pin_memory()
    if device != "cpu":
        batch = batch_to(batch, device)
    return batch


def batch_to(batch: Dict, device, non_blocking: bool = True) -> Dict:
    out = {}
    for k, v in batch.items():
        if isinstance(v, list):
            out[k] = [t.to(device, non_blocking=non_blocking) for t in v]
        else:
            out[k] = v.to(device, non_blocking=non_blocking)
    return out
