#!/bin/bash
# Round 2, visit 11 (1 GPU): suite at HEAD; A/B of (a) the fused pseudo-image gradient with compact deferred gather rows
# (DFB_IMG_GRAD_FUSE) and (b) the deferred weight-gradient unpack (DFB_DEFER_WGRAD); launch list.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -30
for rep in 1 2; do
  for v in "1 1" "0 1" "1 0" "0 0"; do
    set -- $v
    DFB_IMG_GRAD_FUSE=$1 DFB_DEFER_WGRAD=$2 $T 400 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 4 > gpurun_out/ab_$1$2_$rep.log 2> gpurun_out/ab_$1$2_$rep.err
    python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/ab_$1$2_$rep.log") if x.startswith("{")][-1])
print("img_fuse=$1 defer_wgrad=$2 rep=$rep", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", [round(x, 2) for x in d["ms_per_step_regions"]], d["clocks"]["sm_mhz"], d["gpu_launches"], d["loss"])
PY
  done
done
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 4 > gpurun_out/launch_summary.txt 2>&1; head -40 gpurun_out/launch_summary.txt
