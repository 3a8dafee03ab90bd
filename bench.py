#!/usr/bin/env python
"""bench.py -- frame-pairs/s (fwd + loss + bwd + Adam) of the DeFlow hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

N>1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`.
Prints ONE JSON line on rank 0 (contract: task prompt section 4 / "bench.py").

Workload (BASELINE.json configs[1]): DeFlow-GRU (pillar encoder + UNet + 4-iteration GRU decoder),
synthetic AV2-shaped frame pairs, 80 000 points / frame, 512x512 pillars, batch 16 per GPU, bf16
operands in the dense contractions (fp32 accumulation / statistics / voxelisation), full training step.

Only the `cpu_baseline` leg and `--impl reference` execute anything under oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frame_pairs_per_sec_fwd_bwd"
UNIT = "frame-pairs/s"
VS, RG = [0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="frame pairs per GPU")
    ap.add_argument("--points", type=int, default=80000, help="points per frame")
    ap.add_argument("--grid", type=int, default=512)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--decoder", default="gru", choices=["gru", "linear"])
    ap.add_argument("--loss", default="deflowLoss", choices=["deflowLoss", "ff3dLoss"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scatter", action="store_true", help="skip the config-5 scatter microbench")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scatter-only", action="store_true",
                    help="run only the config-5 scatter microbench (BASELINE configs[4]) and print its JSON line")
    ap.add_argument("--cpu-points", type=int, default=None)
    ap.add_argument("--sync-bn", action="store_true",
                    help="BatchNorm statistics over all ranks (the reference's default sync_bn: true, OSF/conf/config.yaml:23)")
    ap.add_argument("--min-seconds", type=float, default=3.0,
                    help="repeat the K-step timed region until this much device time has been measured (median reported)")
    ap.add_argument("--no-flow-err", action="store_true", help="skip the bf16-vs-parity-mode flow error on the bench batch")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, reasons, smax, pw = [], set(), None, []
        for r in rows:
            try:
                sm.append(float(r[0])); smax = float(r[1]); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.strip().lower() == "active":
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w_max": max(pw) if pw else None, "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def cpu_reference_pairs_per_sec(points, grid, decoder, loss, steps=1, warmup=0):
    """The reference's CPU path for the same step.  kind "reference": the reference's OWN Python modules (oracle/_ref/osf,
    staged by oracle/build_ref.py; /root/reference in the build container) on torch-CPU, with the numpy restatement of the
    three mmcv._ext functions under them (the reference's extension registers CUDA kernels only).  kind "port": the
    functional restatement oracle/deflow_oracle.py, when the staged modules are absent.  One step = fwd + loss + bwd of ONE
    frame pair on all host threads."""
    from deflow_b200 import synth
    from oracle import ref_modules
    torch.set_num_threads(os.cpu_count())
    scale = grid / 512.0
    vs = [0.2 / scale, 0.2 / scale, 6]
    batch = synth.make_batch(1, points, seed=synth.SEED_BASE)
    kind = "port"
    if ref_modules.root() is not None:
        try:
            DeFlow, _, lossns = ref_modules.load_reference("numpy")
            model = DeFlow(vs, RG, [grid, grid], decoder, 4)
            model.apply(ref_modules.load_weights_init())
            model.train()
            kind = "reference"

            def one():
                res = model(batch)
                idx = res["pc0_valid_point_idxes"][0]
                l = lossns[loss]({"est_flow": res["flow"][0], "gt_flow": batch["flow"][0][idx] - res["pose_flow"][0][idx],
                                  "gt_classes": batch["flow_category_indices"][0][idx]})["loss"]
                l.backward()
                model.zero_grad(set_to_none=True)
        except Exception as ex:  # noqa: BLE001
            print(f"[bench] reference modules unusable ({ex!r}); timing the oracle port", file=sys.stderr)
            kind = "port"
    if kind == "port":
        from oracle import deflow_oracle as orc
        state = orc.random_state(1, decoder)
        for k, v in state.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        buffers = {k: v.clone() for k, v in state.items() if "running" in k}

        def one():
            res = orc.deflow_forward(batch, state, vs, RG, (grid, grid), decoder, 4, training=True, buffers=buffers)
            l = orc.training_step_loss(batch, res, loss)
            l.backward()
            for v in state.values():
                v.grad = None
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = sum(times) / len(times)
    return 1.0 / t, t, os.cpu_count(), kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # every step is one frame pair (fwd + loss + bwd); ~1.2 s each on 16 threads, so the driver's K fits as given
    steps = max(1, min(args.steps, 40))
    warm = 1 if args.warmup > 0 else 0
    pts = args.cpu_points or args.points
    v, t, cores, kind = cpu_reference_pairs_per_sec(pts, args.grid, args.decoder, args.loss, steps, warm)
    what = ("the reference's own Python modules on torch-CPU + numpy stand-in for its CUDA-only mmcv._ext" if kind == "reference"
            else "fp32 torch-CPU oracle port")
    sample = f"{steps} step(s) of 1 frame pair, {pts} pts/frame, {args.grid}x{args.grid}, fwd+loss+bwd, fp32, {what}, {cores} threads"
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def baseline_config_label(args):
    """Which BASELINE.json `configs` entry this command line is (labels the workload; the default is configs[1])."""
    if args.decoder == "linear" or args.loss == "ff3dLoss":
        return "BASELINE configs[3]: fastflow3d ablation" if args.points == 80000 else "fastflow3d ablation, non-BASELINE size"
    if args.points == 80000 and args.grid == 512 and args.batch == 16:
        return "BASELINE configs[1]"
    if args.points == 120000 and args.grid == 512 and args.batch == 16:
        return "BASELINE configs[2]"
    if args.points == 20000 and args.grid == 512 and args.batch == 1:
        return "BASELINE configs[0]"
    return "non-BASELINE size"


def workload_config(args, world):
    return {"workload": f"DeFlow-{args.decoder.upper()} training step (pillar encoder + UNet + "
                        f"{'4-iter GRU' if args.decoder == 'gru' else 'linear'} decoder + {args.loss} + Adam), "
                        f"{args.points} pts/frame, {args.grid}x{args.grid} pillars, batch {args.batch}/GPU ({baseline_config_label(args)})",
            "batch_per_gpu": args.batch, "global_batch": args.batch * world, "points_per_frame": args.points,
            "grid": [args.grid, args.grid], "precision": args.precision, "parallelism": f"dp{world}",
            "l2": "per-step working set (>2 GB of activations) exceeds the 126 MB L2; no explicit flush",
            "sync_bn": bool(getattr(args, "sync_bn", False))}


# ------------------------------------------------------------------------------------------ native arm
def scatter_microbench(dev, peaks, frames=32, n=200000, grid=1024, iters=5, channels=32):
    """BASELINE configs[4]: voxelisation stress -- 200k pts/frame, 1024x1024 grid, 32 frames per launch,
    scatter-only forward + backward (pillar index + fused PFN fwd + PFN bwd).  HBM GB/s on algorithmic bytes, three ways
    (see the comment at the return)."""
    from deflow_b200 import ops, synth
    import deflow_b200 as d
    vs = [0.1 * 1024 / grid, 0.1 * 1024 / grid, 6]
    b = synth.make_batch(4, n, seed=7)
    pts = torch.cat([b["pc0"][:, :n], b["pc1"][:, :n]], 0)
    reps = (frames + pts.shape[0] - 1) // pts.shape[0]
    pts = pts.repeat(reps, 1, 1)[:frames].contiguous().to(dev)
    emb = d.DynamicEmbedder(vs, [grid, grid], RG, 32).to(dev).train()
    emb.reuse_canvas = True   # as in a training step: the canvas persists, only the previous call's pillar rows are cleared
    gimg = None
    # all iterations are queued back to back (one synchronize at the end): the events time the device, not the host's
    # launch latency after an idle GPU.  embed = pillar index + fused PFN forward in ONE call (the dense canvas zero-fill
    # runs on a side stream under the index kernels, exactly as in DeFlow.forward).
    evs = []
    for it in range(iters + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        image, idx = emb.embed(pts, torch.bfloat16)
        ev[1].record()
        if gimg is None:
            gimg = torch.randn_like(image)
        image.backward(gimg)
        ev[2].record()
        evs.append(ev)
        del image
    torch.cuda.synchronize()
    t_fwd = t_bwd = 0.0
    for ev in evs[2:]:
        t_fwd += ev[0].elapsed_time(ev[1]); t_bwd += ev[1].elapsed_time(ev[2])
    t_fwd, t_bwd = t_fwd / iters, t_bwd / iters
    N = idx.pt_off(frames)
    M = idx.pil_off(frames)
    Nin = frames * pts.shape[1]
    C = 32
    # voxelise 24*N_in; scatter fwd 4NC+12N read + 4MC+12M+4N+4M written, x2 scatters (C=3, C=32);
    # pillar->image 4MC+12M + 2*C*H*W (bf16); bwd 4MC+4N+4M read + 4NC written.
    by_idx = 24 * Nin
    by_fwd = (4 * N * 3 + 12 * N + 4 * M * 3 + 12 * M + 4 * N + 4 * M) + (4 * N * C + 12 * N + 4 * M * C + 12 * M + 4 * N + 4 * M) \
        + (4 * M * C + 12 * M + 2 * C * grid * grid * frames)
    by_bwd = 4 * M * C + 4 * N + 4 * M + 4 * N * C
    by_canvas = 2 * C * grid * grid * frames       # the dense zero canvas of PointPillarsScatter (encoder.py:135-141)
    tot_ms = t_fwd + t_bwd
    by_8d = by_idx + by_fwd + by_bwd
    gbs_8d = by_8d / (tot_ms * 1e-3) / 1e9
    # (1) SURVEY 8d bytes of the REFERENCE's algorithm; (2) the same minus the dense canvas that the persistent pseudo-image
    # never rewrites (only the previous call's M pillar rows are cleared: 2*C*M bytes) -- the bytes THIS implementation has
    # to move, and the fraction reported as the roofline; (3) DRAM bytes actually moved per iteration, from the committed
    # ncu capture of this command (profiles/*_scatter_traffic.json: sum over the scatter kernels of
    # dram__bytes_read.sum + dram__bytes_write.sum), null when no capture is committed.
    by_impl = by_8d - by_canvas + 2 * C * M
    gbs_impl = by_impl / (tot_ms * 1e-3) / 1e9
    actual = None
    for name in ("r02_scatter_traffic.json", "r01_scatter_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("frames") == frames and tj.get("points_per_frame") == n and tj.get("grid") == grid:
                actual = {"dram_bytes": tj["dram_bytes_per_iteration"], "source": "profiles/" + name}
                break
    out = {"workload": f"{frames} frames x {n} pts, {grid}x{grid} grid, index + fused PFN fwd + bwd (BASELINE configs[4])",
           "canvas": "persistent pseudo-image: the previous call's pillar rows are cleared, the dense zero canvas is not rewritten",
           "valid_points": N, "pillars": M, "max_points_per_pillar": int(idx.pil_cnt[:M].max()) if M > 0 else 0,
           "ms": {"embed(index+pfn_fwd)": t_fwd, "pfn_bwd": t_bwd},
           "bytes_8d_reference_layout": by_8d, "bytes_8d_minus_elided_canvas": by_impl,
           "gbs_8d_reference_layout": gbs_8d, "frac_8d_reference_layout": gbs_8d / peaks["hbm_gbs"],
           "achieved_gbs": gbs_impl, "frac_of_hbm_peak": gbs_impl / peaks["hbm_gbs"],
           "algorithmic_bytes": by_impl}
    if actual:
        out["dram_bytes_actual"] = actual["dram_bytes"]
        out["gbs_actual_dram"] = actual["dram_bytes"] / (tot_ms * 1e-3) / 1e9
        out["frac_actual_dram"] = out["gbs_actual_dram"] / peaks["hbm_gbs"]
        out["dram_source"] = actual["source"]
    return out


def run_native(args):
    import torch.distributed as dist
    import deflow_b200 as d
    from deflow_b200 import _lib, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from deflow_b200 import dist as dd
    dd.init("nccl", dev)
    _lib.lib()  # fail loudly if the CUDA library is missing
    peaks = load_peaks()
    if args.scatter_only:
        # BASELINE configs[4]: every rank runs an independent replica of the microbench ("replicas only"); the
        # aggregate is world x bytes over the slowest rank's time
        n_pts = args.points if args.points != 80000 else 200000
        grid = args.grid if args.grid != 512 else 1024
        if world > 1:
            dist.barrier()
        sc = scatter_microbench(dev, peaks, n=n_pts, grid=grid, iters=max(args.steps, 5))
        tot = torch.tensor([sum(sc["ms"].values())], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        tot = float(tot)
        if rank == 0:
            agg = world * sc["algorithmic_bytes"] / (tot * 1e-3) / 1e9
            print(json.dumps({"metric": "scatter_path_hbm_gbs", "value": agg, "unit": "GB/s", "n_gpus": world,
                              "steps": max(args.steps, 5), "warmup": 2, "ms_per_step": tot, "higher_is_better": True,
                              "scaling": "weak (independent replicas)", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                              "config": {"workload": sc["workload"]},
                              "roofline": {"bound": "hbm", "achieved": agg / world, "peak": peaks["hbm_gbs"],
                                           "unit": "GB/s", "frac": agg / world / peaks["hbm_gbs"],
                                           "traffic": sc.get("dram_bytes_actual"),
                                           "note": "achieved = SURVEY 8d bytes minus the dense canvas the implementation does not "
                                                   "rewrite, per GPU; scatter.frac_8d_reference_layout / frac_actual_dram give the "
                                                   "other two readings"},
                              "scatter": sc}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    scale = args.grid / 512.0
    vs = [0.2 / scale, 0.2 / scale, 6]
    torch.manual_seed(synth.SEED_BASE + rank)
    model = d.DeFlow(vs, RG, [args.grid, args.grid], args.decoder, 4, precision=args.precision)
    model.apply(d.weights_init)
    model = model.to(dev).train()
    dd.broadcast_module(model)
    from deflow_b200.trainer import TrainStep
    # lr: the reference's configured default (OSF/conf/config.yaml:29, lr 2e-6); clip 5.0 (:26)
    step = TrainStep(model, lr=2e-6, loss_fn=args.loss, gradient_clip_val=5.0, sync_bn=args.sync_bn)

    host = synth.make_batch(args.batch, args.points, seed=synth.SEED_BASE + rank, pin=True)
    host["pose0"] = torch.stack(host["pose0"]).pin_memory()
    host["pose1"] = torch.stack(host["pose1"]).pin_memory()
    resident = synth.batch_to(host, dev, non_blocking=False)
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps, out

    n_warm = args.warmup if os.environ.get("DFB_PROFILE") else max(args.warmup, 3)  # profiling runs may warm up less
    # the sampler starts BEFORE the warm-up: nvidia-smi's start-up takes a driver lock for tens of milliseconds, which
    # must not land inside the timed region (it samples every 200 ms through warm-up and the timed steps, all under load)
    clocks = ClockSampler(local) if rank == 0 else None
    for _ in range(n_warm):
        loss = step(resident)
    barrier()
    l0 = _lib.launch_count()
    from deflow_b200 import conv as tcconv
    ms, loss = timed(lambda: step(resident), args.steps)
    launches = _lib.launch_count() - l0
    # the K-step region is repeated (each repeat times EXACTLY K steps, barrier + synchronize on both sides) until
    # --min-seconds of device time has been measured; the median region is reported, all of them are listed
    regions = [ms]
    if not os.environ.get("DFB_PROFILE"):
        while sum(regions) * args.steps * 1e-3 < args.min_seconds and len(regions) < 12:
            m2, loss = timed(lambda: step(resident), args.steps)
            regions.append(m2)
    ms = sorted(regions)[len(regions) // 2]
    clk = clocks.stop() if clocks else None
    pairs = args.batch * world
    value = pairs / (ms * 1e-3)
    # the same K steps once more with CUDA events around every tensor-core convolution launch (per-kernel roofline);
    # kept out of the `value` region because ~100 event pairs per step cost ~2 % of the step
    conv_records = None
    if rank == 0 and not os.environ.get("DFB_PROFILE"):
        tcconv.TIMING = []
    ms_ev, _ = timed(lambda: step(resident), args.steps)
    conv_records, tcconv.TIMING = tcconv.TIMING, None

    # stage breakdown (one extra step with events; not part of `value`)
    stages = {}
    # (with SyncBatchNorm every forward / backward contains collectives, so the extra step has to run on all ranks)
    if rank == 0 or args.sync_bn:
        evs = {}
        orig = {}

        def wrap(name, obj, attr):
            f = getattr(obj, attr)
            orig[(obj, attr)] = f

            def g(*a, **k):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); r = f(*a, **k); e.record()
                evs.setdefault(name, []).append((s, e))
                return r
            setattr(obj, attr, g)
        wrap("embed(index+pfn)", model.embedder, "embed")
        wrap("unet_fwd", model.backbone, "forward_nhwc")
        wrap("decoder_fwd", model.head, "forward_flat")
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        step.grads.zero()
        s0.record()
        res = model(resident)
        lo = d.training_step_loss(resident, res, args.loss)
        s1.record()
        lo.backward()
        s2.record()
        torch.cuda.synchronize()
        for (obj, attr), f in orig.items():
            setattr(obj, attr, f)
        stages = {k: sum(s.elapsed_time(e) for s, e in v) for k, v in evs.items()}
        stages["forward_total"] = s0.elapsed_time(s1)
        stages["backward_total"] = s1.elapsed_time(s2)
        idx = res["_dfb"]["index"]
        stages["valid_points_pc0"] = idx.pt_off(args.batch)
        stages["pillars_total"] = idx.pil_off(2 * args.batch)
    if world > 1:
        dist.barrier()

    # end to end through the public API with HOST (pinned) inputs: H2D inside the timed region, loss read back
    e2e = None
    if not args.no_e2e:
        from deflow_b200.feed import DeviceFeeder
        feeder = DeviceFeeder(dev)
        feeder.submit(host)

        def e2e_step():
            b = feeder.get()          # H2D copy of THIS step's inputs (issued during the previous step, side stream)
            feeder.submit(host)       # start the copy for the next step
            l = step(b)
            return float(l)           # D2H read of the step's result (the summed loss)
        for _ in range(0 if os.environ.get("DFB_PROFILE") else 2):
            e2e_step()
        eregions = []
        while not eregions or (sum(eregions) * args.steps * 1e-3 < args.min_seconds and len(eregions) < 12
                               and not os.environ.get("DFB_PROFILE")):
            ems, _ = timed(e2e_step, args.steps)
            eregions.append(ems)
        ems = sorted(eregions)[len(eregions) // 2]
        e2e = {"value": pairs / (ems * 1e-3), "unit": UNIT, "ms_per_step": ems, "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": 4 * world, "ms_per_step_regions": eregions}

    # flow error of THIS precision mode against the parity mode (split-precision fp32 arithmetic, the mode that meets the
    # north-star 1e-3 bound against the oracle: tests/test_gpu_parity.py) on the bench batch, same weights, train-mode
    # BatchNorm -- SURVEY 8d: "the measured flow error printed beside every perf-mode number"
    flow_err = None
    if (rank == 0 or args.sync_bn) and not args.no_flow_err and not os.environ.get("DFB_PROFILE"):
        try:
            par = d.DeFlow(vs, RG, [args.grid, args.grid], args.decoder, 4, precision="fp32").to(dev).train()
            par.load_state_dict(model.state_dict())
            if args.sync_bn:
                dd.enable_sync_bn(par, step.stat_sync)
            with torch.no_grad():
                fa = model(resident)["_dfb"]["flow_flat"].float()
                fb = par(resident)["_dfb"]["flow_flat"].float()
            df = (fa - fb).abs()
            flow_err = {"vs": "parity mode (precision=fp32, bf16x3 split operands) on the same batch and weights",
                        "max": float(df.max()), "mean": float(df.mean()), "flow_abs_max": float(fb.abs().max()),
                        "flow_abs_mean": float(fb.abs().mean()), "points": int(df.shape[0])}
            del par, fa, fb, df
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            flow_err = {"error": repr(ex)}

    line = None
    if rank == 0:
        scatter = None
        if not args.no_scatter:
            try:
                scatter = scatter_microbench(dev, peaks)
            except Exception as ex:  # noqa: BLE001
                scatter = {"error": repr(ex)}
        # roofline: UNet fwd+bwd FLOPs (343.06 GFLOP fwd / pair @512^2, x3 for training) against the
        # sustained bf16 tensor peak; every dense contraction runs in the library's own tcgen05 kernels.
        flop_pair = 343.06e9 * (args.grid / 512.0) ** 2
        n0 = stages.get("valid_points_pc0", 0)
        dec_flop = (602688 if args.decoder == "gru" else 17344) * n0
        flops_step = 3 * (flop_pair * args.batch + dec_flop)
        whole = flops_step / (ms * 1e-3) / 1e12
        # per-kernel roofline of the tensor-core convolution families, measured live with CUDA events over the timed region
        fam = {}
        for name, fl, e0, e1 in conv_records or []:
            d = fam.setdefault(name, [0, 0.0, 0.0])
            d[0] += 1; d[1] += fl; d[2] += e0.elapsed_time(e1)
        kernels = sorted(({"kernel": k, "launches_per_step": v[0] / args.steps, "ms_per_step": v[2] / args.steps,
                           "avg_launch_ms": v[2] / v[0], "tflops": v[1] / (v[2] * 1e-3) / 1e12,
                           "frac_of_sustained_peak": v[1] / (v[2] * 1e-3) / 1e12 / peaks["bf16_tflops_sustained"]}
                          for k, v in fam.items()), key=lambda r: -r["ms_per_step"])
        if kernels:
            top = kernels[0]
            # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this command
            # (profiles/r0N_traffic.json, written by tools/traffic_table.py: dram__bytes_read.sum + dram__bytes_write.sum,
            # averaged over the captured launches of that kernel)
            traffic = None
            for tname in ("r02_traffic.json", "r01_traffic.json"):
                tpath = os.path.join(ROOT, "profiles", tname)
                if traffic is None and os.path.exists(tpath):
                    traffic = _traffic_lookup(json.load(open(tpath)), top["kernel"])
            roof = {"bound": "tensor", "kernel": top["kernel"], "achieved": top["tflops"], "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": top["frac_of_sustained_peak"], "traffic": traffic,
                    "avg_launch_ms": top["avg_launch_ms"], "share_of_step": top["ms_per_step"] / ms,
                    "note": f"dominant kernel family by time in the timed region (CUDA events around every launch); achieved = "
                            f"algorithmic conv FLOPs / launch time; peak = sustained bf16 ({peaks['source']}). Whole step: "
                            f"{whole:.0f} TFLOP/s on algorithmic FLOPs (UNet 343.06 GFLOP/pair + decoder 602688 FLOP/pt, x3) = "
                            f"{whole / peaks['bf16_tflops_sustained']:.3f} of sustained peak",
                    "whole_step_tflops": whole, "kernels": kernels[:8]}
        else:
            roof = {"bound": "tensor", "achieved": whole, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": whole / peaks["bf16_tflops_sustained"], "traffic": None,
                    "note": "whole-step algorithmic FLOPs / step time (no per-kernel records in this run)"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:   # the CPU baseline is timed at N = 1 only (rank 0 would stall the others)
            pts = args.cpu_points or args.points
            v, t, cores, kind = cpu_reference_pairs_per_sec(pts, args.grid, args.decoder, args.loss, 10, 1)
            what = ("the reference's own Python modules on torch-CPU (numpy stand-in for its CUDA-only mmcv._ext)"
                    if kind == "reference" else "fp32 torch-CPU oracle port")
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": f"10 steps (+1 warm-up) of 1 frame pair ({pts} pts/frame, {args.grid}x{args.grid}) fwd+loss+bwd, "
                             f"{what}, {cores} threads, {t:.2f} s per step"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
                "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": workload_config(args, world), "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "scatter": scatter, "stages_ms": stages, "loss": float(loss),
                "flow_err": flow_err, "ms_per_step_regions": regions, "lr": 2e-6,
                "sync_bn_exchange": ({"collectives_per_step": step.stat_sync.calls / max(step.global_step, 1),
                                      "bytes_per_step": step.stat_sync.bytes / max(step.global_step, 1)}
                                     if step.stat_sync is not None else None),
                "grad_allreduce": {"bytes": step.grads.nbytes, "op": "AVG", "overlapped_early_slice_params": (
                    len(step.grads.params) - step.grads.early_from if step.grads.early_from is not None else 0)},
                "timed_seconds": sum(regions) * args.steps * 1e-3}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def _traffic_lookup(table: dict, kernel: str):
    """DRAM bytes per launch of a kernel family from a profiles/r0N_traffic.json table.  The table is keyed by the kernel's
    name as ncu prints it (with template arguments, e.g. 'k_gru_fused_bwd<16>'); the timing records name the family
    ('k_gru_fused_bwd', 'k_conv_igemm_halo<128>'): exact key first, then the family's instantiations (mean)."""
    hit = table.get(kernel)
    if isinstance(hit, dict) and "dram_bytes_per_launch" in hit:
        return hit["dram_bytes_per_launch"]
    vals = [v["dram_bytes_per_launch"] for k, v in table.items()
            if isinstance(v, dict) and "dram_bytes_per_launch" in v and k.startswith(kernel + "<")]
    return sum(vals) / len(vals) if vals else None


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
