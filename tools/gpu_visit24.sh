#!/bin/bash
# Round 2, visit 24 (1 GPU): dual-tile variant of the 128-channel halo kernel: tests, single launches, A/B on the step.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q -rf -x -k "dual or forward_dgrad_wgrad or split_precision" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_sub.log | cut -c1-250 | tail -8
echo "--- h128 single launch: default / dual"; $T 120 python tools/prof_general_igemm.py h128 2>&1 | tail -1; DFB_HALO_DUAL=1 $T 120 python tools/prof_general_igemm.py h128 2>&1 | tail -1
run() {
  local label=$1; shift
  env "$@" $T 300 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 2 > gpurun_out/ab_$label.log 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d = json.loads([x for x in open("gpurun_out/ab_$label.log") if x.startswith("{")][-1])
    k = {r["kernel"]: round(r["ms_per_step"], 3) for r in d["roofline"]["kernels"]}
    print("$label", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz loss", round(d["loss"], 5), {n: v for n, v in k.items() if "halo<" in n})
except Exception as e:
    print("$label FAILED", e)
PY
}
run single DFB_HALO_DUAL=0
run dual DFB_HALO_DUAL=1
run single2 DFB_HALO_DUAL=0
run dual2 DFB_HALO_DUAL=1
