#!/bin/bash
# Round 2, visit 15 (1 GPU): fewer TMA requests per tile in the general implicit-GEMM kernel (64-channel stores, resident
# weights), one launch for the four parity planes of the stride-2 data gradients, bias gradients of the point decoder from
# idle rows of its weight-gradient launches, 64-channel stores in the row-pair kernel: tests, then one switch at a time.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 900 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train_step.py tests/test_gpu_parity.py -m gpu -q -rf > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_sub.log | cut -c1-250 | tail -16
run() {  # label, env assignments...
  local label=$1; shift
  env "$@" $T 300 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 2 > gpurun_out/ab_$label.log 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d = json.loads([x for x in open("gpurun_out/ab_$label.log") if x.startswith("{")][-1])
    k = {r["kernel"]: round(r["ms_per_step"], 3) for r in d["roofline"]["kernels"]}
    print("$label", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz loss", round(d["loss"], 5) if "loss" in d else None, k)
except Exception as e:
    print("$label FAILED", e)
PY
}
OLD="DFB_WGRAD_BIAS=0 DFB_DGRAD_PAR_MERGE=0 DFB_EPI_WIDE=0 DFB_IGEMM_B_RESIDENT=0"
run old $OLD
run bias DFB_DGRAD_PAR_MERGE=0 DFB_EPI_WIDE=0 DFB_IGEMM_B_RESIDENT=0
run bias_par DFB_EPI_WIDE=0 DFB_IGEMM_B_RESIDENT=0
run bias_par_wide DFB_IGEMM_B_RESIDENT=0
run default DFB_X=0
run default_pairwide DFB_PAIR_WIDE=1
run old2 $OLD
run default2 DFB_X=0
run default_pairwide2 DFB_PAIR_WIDE=1
DFB_PAIR_WIDE=1 DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-scatter --no-e2e --no-flow-err > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 5 > gpurun_out/launch_summary.txt 2>&1; head -40 gpurun_out/launch_summary.txt
