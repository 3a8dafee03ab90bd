#!/bin/bash
# Round 2, visit 9 (1 GPU): suite at HEAD, A/B of the border-sum bias gradients for the 64-channel 3x3 data gradients
# (DFB_BORDER_COLSUM=0 keeps the column sums in the epilogue of the row-pair kernel), launch list.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=5 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -30
for rep in 1 2; do
  for v in 1 0; do
    DFB_BORDER_COLSUM=$v $T 400 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 4 > gpurun_out/ab_bc${v}_$rep.log 2> gpurun_out/ab_bc${v}_$rep.err
    python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/ab_bc${v}_$rep.log") if x.startswith("{")][-1])
k = {r["kernel"]: round(r["ms_per_step"], 3) for r in d["roofline"]["kernels"]}
print("border_colsum=$v rep=$rep", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", [round(x, 2) for x in d["ms_per_step_regions"]], k.get("k_conv_igemm_halo_pair"), d["clocks"]["sm_mhz"])
PY
  done
done
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 4 > gpurun_out/launch_summary.txt 2>&1; head -30 gpurun_out/launch_summary.txt
