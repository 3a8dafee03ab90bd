#!/bin/bash
# Round 2, visit 13 (1 GPU): weight-gradient kernels at one CTA per SM with floor-rounded splits: tests of the wgrad paths,
# sweep around the new default.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train_step.py -m gpu -q -rf > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_sub.log | cut -c1-250 | tail -12
for x in 10 20 9 10 8; do
  DFB_WGRAD_CTAS_X10=$x $T 300 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 3 > gpurun_out/sw_$x.log 2> gpurun_out/sw_$x.err
  python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/sw_$x.log") if x.startswith("{")][-1])
k = {r["kernel"]: round(r["ms_per_step"], 3) for r in d["roofline"]["kernels"]}
print("ctas_x10=$x", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], {n: v for n, v in k.items() if "wgrad" in n})
PY
done
