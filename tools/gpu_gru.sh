#!/bin/bash
# GRU worker-warp variants: decoder parity test + bench line for DFB_GRU_WARPS = 8 and 16.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
for w in 8 16; do
  DFB_GRU_WARPS=$w $T 300 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train_step.py -q -k "decoder or model or step" > gpurun_out/pytest_gru$w.log 2>&1; echo "pytest gru$w rc=$?"
  tail -3 gpurun_out/pytest_gru$w.log
  DFB_GRU_WARPS=$w $T 400 python bench.py --no-scatter --no-cpu-baseline --no-e2e > gpurun_out/bench_gru$w.log 2> gpurun_out/bench_gru$w.err; echo "bench gru$w rc=$?"
  python - <<PY
import json
d = json.loads([x for x in open('gpurun_out/bench_gru$w.log') if x.startswith('{')][-1])
print('GRU warps $w:', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms', d['stages_ms'])
PY
done
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
