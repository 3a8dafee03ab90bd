#!/usr/bin/env python
"""GPU bar-to-beat (SURVEY.md 8d "GPU bar-to-beat"): the REFERENCE's own modules + the reference's
own CUDA extension (oracle/_ref/mmcv_ref_ext.so) + cuDNN / cuBLAS through torch, timed on the same B200 with the same
synthetic batch and the same training-step arithmetic as bench.py (OSF/src/trainer.py:94-175: forward, per-sample loss
loop, backward, clip 5.0, Adam).  Precision as the reference configures it: fp32 tensors, TF32 matmuls
(`torch.set_float32_matmul_precision('medium')`, OSF/src/trainer.py:36) and cuDNN's default allow_tf32=True.
`--amp` additionally wraps the step in bf16 autocast (NOT something the reference does; a generous extra data point).

    python tools/ref_gpu_bench.py [--batch 16] [--points 80000] [--steps 5] [--warmup 2] [--amp] [--decoder gru]

Test / measurement infrastructure: runs reference code from oracle/_ref/osf (staged by oracle/build_ref.py)."""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--points", type=int, default=80000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--decoder", default="gru")
    ap.add_argument("--loss", default="deflowLoss")
    ap.add_argument("--amp", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from oracle import ref_modules
    from deflow_b200 import synth
    DeFlow, FastFlow3D, lossns = ref_modules.load_reference(ext="cuda")
    weights_init = ref_modules.load_weights_init()
    torch.set_float32_matmul_precision("medium")     # OSF/src/trainer.py:36
    dev = torch.device("cuda", 0)
    torch.manual_seed(synth.SEED_BASE)
    model = DeFlow(decoder_option=args.decoder, num_iters=4).to(dev)
    model.apply(weights_init)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    host = synth.make_batch(args.batch, args.points, seed=synth.SEED_BASE)
    batch = synth.batch_to(host, dev, non_blocking=False)
    loss_fn = lossns[args.loss]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.amp):
            res = model(batch)
            total = 0.0
            for b in range(args.batch):                       # OSF/src/trainer.py:120-142
                idx = res["pc0_valid_point_idxes"][b]
                d = {"est_flow": res["flow"][b].float(), "gt_flow": batch["flow"][b][idx] - res["pose_flow"][b][idx],
                     "gt_classes": batch["flow_category_indices"][b][idx]}
                total = total + loss_fn(d)["loss"]
        total.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        return total

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / args.steps
    line = {"what": "reference modules + reference CUDA ext + cuDNN/cuBLAS on this GPU (bar to beat)",
            "impl": "reference-gpu", "metric": "frame_pairs_per_sec_fwd_bwd", "value": args.batch / (ms * 1e-3),
            "unit": "frame-pairs/s", "ms_per_step": ms, "wall_ms_per_step": wall / args.steps * 1e3, "steps": args.steps,
            "warmup": args.warmup, "batch": args.batch, "points_per_frame": args.points, "decoder": args.decoder,
            "precision": "bf16 autocast (not a reference setting)" if args.amp else "fp32 tensors, TF32 matmul/conv (reference setting)",
            "loss": float(loss), "gpu": torch.cuda.get_device_name(0),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30, "torch": torch.__version__,
            "cudnn": torch.backends.cudnn.version()}
    print(json.dumps(line), flush=True)
    if args.out:
        with open(args.out, "a") as f:
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
