#!/bin/bash
# Round 2, visit 20 (2 GPUs): the validated build at N = 2 (plain and SyncBatchNorm) next to N = 1 on the same box.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --min-seconds 2 > gpurun_out/n1.log 2> gpurun_out/n1.err; echo "n1 rc=$?"
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --min-seconds 2 > gpurun_out/n2.log 2> gpurun_out/n2.err; echo "n2 rc=$?"
$T 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --min-seconds 2 --sync-bn > gpurun_out/n2_syncbn.log 2> gpurun_out/n2_syncbn.err; echo "n2 syncbn rc=$?"
python - <<'PY'
import json
for f in ("n1", "n2", "n2_syncbn"):
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        print(f, round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], "MHz", d["config"].get("sync_bn"))
    except Exception as e:
        print(f, "FAILED", e)
PY
