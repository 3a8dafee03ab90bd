#!/bin/bash
# Round 2, visit 19 (1 GPU): decoder gather backward with block-shared heavy pillars: tests, kernel time in the step, step time.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_train_step.py tests/test_gpu_reference_modules.py -m gpu -q -rf -x > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_sub.log | cut -c1-250 | tail -8
for i in 1 2; do
$T 300 python bench.py --steps 20 --warmup 3 --no-scatter --no-cpu-baseline --no-flow-err --no-e2e --min-seconds 2 > gpurun_out/ab_$i.log 2> gpurun_out/ab_$i.err
python - <<PY
import json
d = json.loads([x for x in open("gpurun_out/ab_$i.log") if x.startswith("{")][-1])
print("run$i", round(d["value"], 1), "pairs/s", round(d["ms_per_step"], 3), "ms", d["clocks"]["sm_mhz"], "MHz loss", round(d["loss"], 5))
PY
done
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_decoder_gather|k_channel_sum|k_offset_encode" --csv --log-file gpurun_out/launches_gather.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-scatter --no-e2e --no-flow-err > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
grep -o '"dfb::k_[a-z_]*[^"]*".*' gpurun_out/launches_gather.csv | awk -F'","' '{print $1, $NF}' | tail -12
