"""Generate tests/golden/ext_gpu_ref.npz by running the REFERENCE's own CUDA extension (compiled from
/root/reference into oracle/_ref by oracle/build_ref.py) on a B200:

    gpurun -- python tests/golden/make_ext_golden_gpu.py gpurun_out/ext_gpu_ref.npz   # then copy into tests/golden/

Seeded inputs (regenerated identically by tests/test_oracle_ext_golden.py) -> outputs of the reference kernels
dynamic_voxelize_forward / dynamic_point_to_voxel_forward / _backward.  These pin oracle/mmcv_ext_oracle.py."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import build_ref  # noqa: E402

VOX_CASES = [([0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]), ([0.1, 0.1, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]),
             ([0.3, 0.25, 0.2], [-10, -7, -3, 11.1, 8.3, 3.4])]


def vox_points(n, seed, rg):
    rng = np.random.default_rng(seed)
    p = rng.uniform(-1.2, 1.2, size=(n, 3)).astype(np.float32) * np.array([rg[3], rg[4], 4.0], np.float32)
    k = n // 4
    p[:k, :2] = (np.round(p[:k, :2] / 0.2) * 0.2).astype(np.float32)      # exactly on voxel boundaries
    p[k:2 * k, 0] = np.nextafter(p[k:2 * k, 0], np.float32(np.inf))
    p[0] = [rg[0], rg[1], rg[2]]
    p[1] = [rg[3], rg[4], rg[5]]
    p[2] = [np.nextafter(np.float32(rg[3]), np.float32(-np.inf)), 0, 0]
    return p


def scatter_case(n, c, span, seed):
    rng = np.random.default_rng(seed)
    coors = np.stack([rng.integers(0, 2, n), rng.integers(0, span, n), rng.integers(0, span, n)], 1).astype(np.int32)
    bad = rng.random(n) < 0.1
    coors[bad, rng.integers(0, 3, int(bad.sum()))] = -1
    feats = (np.round(rng.normal(size=(n, c)) * 8) / 8).astype(np.float32)   # exactly summable: sums are order-free
    gv_seed = seed + 1
    return coors, feats, gv_seed


SCATTER_CASES = [(513, 3, 6, 11), (5000, 32, 20, 12), (60000, 32, 300, 13), (3000, 7, 3, 14)]
RECORDED = (0, 1, 3)   # case 2 is compared live on the GPU only (tests/test_gpu_vs_reference_ext.py): 24 MB of fixtures


def main(out):
    ext = build_ref.load_ref()
    assert ext is not None, "oracle/_ref/mmcv_ref_ext.so is missing: run python oracle/build_ref.py in the build container"
    dev = "cuda"
    fix = {}
    for i, (vs, rg) in enumerate(VOX_CASES):
        pts = vox_points(20000, 100 + i, rg)
        coors = torch.zeros((pts.shape[0], 3), dtype=torch.int32, device=dev)
        ext.dynamic_voxelize_forward(torch.from_numpy(pts).to(dev), torch.tensor(vs, dtype=torch.float32),
                                     torch.tensor(rg, dtype=torch.float32), coors, 3)
        fix[f"vox{i}_coors"] = coors.cpu().numpy()
    for i, (n, c, span, seed) in enumerate(SCATTER_CASES):
        if i not in RECORDED:
            continue
        coors, feats, gseed = scatter_case(n, c, span, seed)
        for red in ("sum", "mean", "max"):
            vf, vc, cmap, cnt = ext.dynamic_point_to_voxel_forward(torch.from_numpy(feats).to(dev), torch.from_numpy(coors).to(dev), red)
            fix[f"sc{i}_{red}_feats"] = vf.cpu().numpy()
            if red == "sum":
                fix[f"sc{i}_coors"], fix[f"sc{i}_map"], fix[f"sc{i}_count"] = vc.cpu().numpy(), cmap.cpu().numpy(), cnt.cpu().numpy()
            gv = np.random.default_rng(gseed).normal(size=tuple(vf.shape)).astype(np.float32)
            g = torch.zeros((n, c), device=dev)
            ext.dynamic_point_to_voxel_backward(g, torch.from_numpy(gv).to(dev), torch.from_numpy(feats).to(dev), vf, cmap, cnt, red)
            fix[f"sc{i}_{red}_grad"] = g.cpu().numpy()
    np.savez_compressed(out, **fix)
    print("wrote", out, {k: v.shape for k, v in list(fix.items())[:6]})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "ext_gpu_ref.npz"))
