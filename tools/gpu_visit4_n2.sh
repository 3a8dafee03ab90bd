#!/bin/bash
# Round 2, second 2-GPU visit: SyncBatchNorm bench (its cost against the per-rank-statistics line), config 3 / config 4 /
# config-5 replicas at N = 2.  TERM first so that torchrun takes its workers down with it.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
$T 400 $R bench.py --gpus 2 --steps 20 --warmup 3 --no-scatter --no-flow-err > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "n2 rc=$?"
$T 400 $R bench.py --gpus 2 --steps 20 --warmup 3 --no-scatter --sync-bn > gpurun_out/bench_n2_syncbn.log 2> gpurun_out/bench_n2_syncbn.err; echo "n2 syncbn rc=$?"; tail -3 gpurun_out/bench_n2_syncbn.err | cut -c1-300
nvidia-smi --query-compute-apps=pid,used_memory --format=csv
$T 400 $R bench.py --gpus 2 --steps 10 --warmup 3 --no-scatter --points 120000 --no-flow-err > gpurun_out/bench_n2_cfg3.log 2> gpurun_out/bench_n2_cfg3.err; echo "n2 cfg3 rc=$?"
$T 400 $R bench.py --gpus 2 --steps 10 --warmup 3 --no-scatter --decoder linear --loss ff3dLoss --no-flow-err > gpurun_out/bench_n2_cfg4.log 2> gpurun_out/bench_n2_cfg4.err; echo "n2 cfg4 rc=$?"
$T 300 $R bench.py --gpus 2 --scatter-only --steps 5 > gpurun_out/bench_n2_scatter.log 2> gpurun_out/bench_n2_scatter.err; echo "n2 scatter rc=$?"
python - <<'PY'
import json
for f in ["bench_n2", "bench_n2_syncbn", "bench_n2_cfg3", "bench_n2_cfg4", "bench_n2_scatter"]:
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ["value", "ms_per_step", "n_gpus", "loss", "sync_bn_exchange", "flow_err"]}, (d.get("e2e") or {}).get("value"), d["config"].get("sync_bn"), d["config"]["workload"][-40:])
    except Exception as e:
        print(f, "parse failed", e)
PY
