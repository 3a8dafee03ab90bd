#!/bin/bash
# Round 2, visit 17 (1 GPU): general implicit-GEMM kernel with per-tile latency chains cut (no divisions in the tile decode,
# whole-warp MMA issue, alternate-tile epilogue groups): tests, single-launch timings, A/B on the step.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train_step.py tests/test_gpu_parity.py -m gpu -q -rf -x > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_sub.log | cut -c1-250 | tail -16
echo "--- single launches, default"; $T 120 python tools/prof_general_igemm.py 2>&1 | tail -4
echo "--- single launches, DFB_EPI_ALT=0"; DFB_EPI_ALT=0 $T 120 python tools/prof_general_igemm.py 2>&1 | tail -4
echo "--- single launches, DFB_EPI_WIDE=0"; DFB_EPI_WIDE=0 $T 120 python tools/prof_general_igemm.py 2>&1 | tail -4
echo "--- single launches, DFB_IGEMM_B_RESIDENT=0"; DFB_IGEMM_B_RESIDENT=0 $T 120 python tools/prof_general_igemm.py 2>&1 | tail -4
run() {  # label, env assignments...
  local label=$1; shift
  env "$@" $T 300 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 2 > gpurun_out/ab_$label.log 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d = json.loads([x for x in open("gpurun_out/ab_$label.log") if x.startswith("{")][-1])
    print("$label", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz loss", round(d["loss"], 5))
except Exception as e:
    print("$label FAILED", e)
PY
}
run noalt DFB_EPI_ALT=0
run default DFB_X=0
run noalt2 DFB_EPI_ALT=0
run default2 DFB_X=0
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-scatter --no-e2e --no-flow-err > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 5 > gpurun_out/launch_summary.txt 2>&1; grep "k_conv_igemm<\|sum of kernel" gpurun_out/launch_summary.txt
