#!/bin/bash
# A/B at N = 2: early all-reduce slice (overlapped with the encoder backward) against one collective after the backward.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514"
for rep in 1 2; do
  for ov in 1 0; do
    DFB_ALLREDUCE_OVERLAP=$ov $T 300 $R bench.py --gpus 2 --steps 20 --warmup 3 --no-scatter --no-flow-err --no-e2e --min-seconds 4 > gpurun_out/ab_ov${ov}_$rep.log 2> gpurun_out/ab_ov${ov}_$rep.err
    python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/ab_ov${ov}_$rep.log") if x.startswith("{")][-1])
print("overlap=$ov rep=$rep", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", [round(x, 2) for x in d["ms_per_step_regions"]], d["grad_allreduce"])
PY
  done
done
$T 200 python bench.py --steps 20 --warmup 3 --no-scatter --no-flow-err --no-e2e --no-cpu-baseline --min-seconds 4 > gpurun_out/ab_n1.log 2> gpurun_out/ab_n1.err
python - <<'PY'
import json
d = json.loads([x for x in open("gpurun_out/ab_n1.log") if x.startswith("{")][-1])
print("N=1 on the same box", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", [round(x, 2) for x in d["ms_per_step_regions"]])
PY
