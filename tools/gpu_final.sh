#!/bin/bash
# Final check of HEAD: GPU parity tests (as the driver runs them) and the default bench line.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 600 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
$T 400 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads([x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks']})
print('e2e', d['e2e']); print('roofline', {k: d['roofline'][k] for k in ['kernel', 'achieved', 'peak', 'frac', 'traffic', 'share_of_step']})
print('cpu', d['cpu_baseline']); print('scatter', d['scatter']['achieved_gbs'], d['scatter']['frac_of_hbm_peak'])
PY
