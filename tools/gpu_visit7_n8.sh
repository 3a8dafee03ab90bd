#!/bin/bash
# Round 2, 8-GPU visit: what the driver's own scaling run does not cover -- BASELINE configs[2] (120 k points, DDP on 8 GPUs),
# configs[3] (fastflow3d ablation on 8 GPUs), the SyncBatchNorm line and the config-5 scatter replicas at N = 8.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513"
$T 300 $R bench.py --gpus 8 --steps 10 --warmup 3 --no-scatter --points 120000 --no-flow-err > gpurun_out/bench_n8_cfg3.log 2> gpurun_out/bench_n8_cfg3.err; echo "n8 cfg3 rc=$?"; tail -2 gpurun_out/bench_n8_cfg3.err | cut -c1-200
$T 300 $R bench.py --gpus 8 --steps 10 --warmup 3 --no-scatter --decoder linear --loss ff3dLoss --no-flow-err > gpurun_out/bench_n8_cfg4.log 2> gpurun_out/bench_n8_cfg4.err; echo "n8 cfg4 rc=$?"
$T 300 $R bench.py --gpus 8 --steps 10 --warmup 3 --no-scatter --sync-bn --no-flow-err > gpurun_out/bench_n8_syncbn.log 2> gpurun_out/bench_n8_syncbn.err; echo "n8 syncbn rc=$?"
$T 300 $R bench.py --gpus 8 --steps 10 --warmup 3 --no-scatter --no-flow-err > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err; echo "n8 rc=$?"
$T 200 $R bench.py --gpus 8 --scatter-only --steps 5 > gpurun_out/bench_n8_scatter.log 2> gpurun_out/bench_n8_scatter.err; echo "n8 scatter rc=$?"
python - <<'PY'
import json
for f in ["bench_n8", "bench_n8_syncbn", "bench_n8_cfg3", "bench_n8_cfg4", "bench_n8_scatter"]:
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ["value", "ms_per_step", "n_gpus", "loss", "sync_bn_exchange"]}, (d.get("e2e") or {}).get("value"), d["config"].get("sync_bn"), d["config"]["workload"][-40:])
    except Exception as e:
        print(f, "parse failed", e)
PY
