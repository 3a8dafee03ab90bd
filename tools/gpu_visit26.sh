#!/bin/bash
# Round 2, visit 26 (1 GPU): the other BASELINE configs on the final build at N = 1 (configs[2] 120 k points, configs[3] linear + ff3dLoss).
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 200 python bench.py --steps 20 --warmup 3 --points 120000 --no-scatter --no-cpu-baseline --no-flow-err --min-seconds 1.5 > gpurun_out/cfg3.log 2> gpurun_out/cfg3.err; echo "cfg3 rc=$?"
$T 200 python bench.py --steps 20 --warmup 3 --decoder linear --loss ff3dLoss --no-scatter --no-cpu-baseline --no-flow-err --min-seconds 1.5 > gpurun_out/cfg4.log 2> gpurun_out/cfg4.err; echo "cfg4 rc=$?"
python - <<'PY'
import json
for f in ("cfg3", "cfg4"):
    try:
        d = json.loads([x for x in open(f"gpurun_out/{f}.log") if x.startswith("{")][-1])
        print(f, d["config"]["workload"][:70], round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "FAILED", e)
PY
