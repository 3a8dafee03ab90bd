#!/usr/bin/env python
"""BASELINE configs[4] sweep: the scatter path over N in {25k .. 400k} points / frame on a 1024 x 1024
grid, 32 frames per launch sequence -- (a) the fused path DeFlow uses (pillar index + fused PFN forward + backward, C = 32),
(b) the mmcv._ext drop-ins called the way the reference calls them, one frame per call, mean reduce, forward + backward,
C = 3 (cluster_scatter on xyz) and C = 32 (pfn_scatter on point features).  GB/s on SURVEY 8d bytes.

    python tools/scatter_sweep.py [--out gpurun_out/scatter_sweep.txt]"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from deflow_b200 import ops, synth  # noqa: E402

RG = [-51.2, -51.2, -3, 51.2, 51.2, 3]


def ext_path(dev, n, C, frames=8, iters=5):
    """dynamic_point_to_voxel_forward + backward per frame (scatter_points.py:25-67), the reference's call pattern."""
    vs = [0.1, 0.1, 6]
    b = synth.make_batch(frames, n, seed=11)
    idx = ops.pillar_index(b["pc0"].to(dev), vs, RG)
    per = []
    for f in range(frames):
        a, e = idx.pt_off(f), idx.pt_off(f + 1)
        coors = idx.pt_coor[a:e].contiguous()
        feats = idx.pt_xyz[a:e].contiguous() if C == 3 else torch.randn(e - a, C, device=dev)
        per.append((feats, coors))
    tot_bytes = 0
    evs = []
    for it in range(iters + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for feats, coors in per:
            vf, vc, cmap, cnt = ops.dynamic_point_to_voxel_forward(feats, coors, "mean")
            g = torch.empty_like(feats)
            ops.dynamic_point_to_voxel_backward(g, vf, feats, vf, cmap, cnt, "mean")
            if it == 0:
                N, M = feats.shape[0], vf.shape[0]
                tot_bytes += (4 * N * C + 12 * N) + (4 * M * C + 12 * M + 4 * N + 4 * M) + (4 * M * C + 4 * N + 4 * M) + 4 * N * C
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs[1:]) / iters
    return ms, tot_bytes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peaks = bench.load_peaks()
    lines = ["# scatter sweep, 1024x1024 grid (voxel 0.1 m), B200; fused = pillar index + fused PFN fwd + bwd over 32 frames per call; "
             "ext = mmcv._ext drop-in scatter fwd+bwd (mean), one frame per call, 8 frames",
             f"# HBM peak {peaks['hbm_gbs']:.0f} GB/s ({peaks['source']})",
             f"{'pts/frame':>9s} | {'fused ms':>8s} {'GB/s 8d':>8s} {'frac':>5s} {'GB/s -canvas':>12s} {'frac':>5s} {'M/N':>5s} | "
             f"{'ext C=3 ms':>10s} {'GB/s':>6s} | {'ext C=32 ms':>11s} {'GB/s':>6s}"]
    rows = []
    bench.scatter_microbench(dev, peaks, n=100000, grid=1024, iters=3)     # global warm-up (allocator, function attributes)
    ext_path(dev, 100000, 32, iters=2)
    for n in (25000, 50000, 100000, 200000, 400000):
        bench.scatter_microbench(dev, peaks, n=n, grid=1024, iters=2)      # per-size warm-up: buffers of this size exist
        sc = bench.scatter_microbench(dev, peaks, n=n, grid=1024, iters=8)
        ms = sum(sc["ms"].values())
        ext_path(dev, n, 3, iters=1)
        m3, b3 = ext_path(dev, n, 3)
        m32, b32 = ext_path(dev, n, 32)
        rows.append({"n": n, "fused": sc, "ext_c3_ms": m3, "ext_c3_gbs": b3 / m3 / 1e6, "ext_c32_ms": m32, "ext_c32_gbs": b32 / m32 / 1e6})
        lines.append(f"{n:9d} | {ms:8.3f} {sc['gbs_8d_reference_layout']:8.0f} {sc['frac_8d_reference_layout']:5.2f} "
                     f"{sc['achieved_gbs']:12.0f} {sc['frac_of_hbm_peak']:5.2f} {sc['pillars'] / max(sc['valid_points'], 1):5.2f} | "
                     f"{m3:10.3f} {b3 / m3 / 1e6:6.0f} | {m32:11.3f} {b32 / m32 / 1e6:6.0f}")
        print(lines[-1], flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("\n".join(lines) + "\n")
        with open(args.out.replace(".txt", ".json"), "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
