#!/bin/bash
# Round 2, final 1-GPU validation of HEAD: the GPU suite as the driver runs it, smoke, the default bench line, the parity-mode
# line, the scatter-only line, the reference arm, a launch list.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
$T 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
$T 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads([x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1])
print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks', 'flow_err', 'ms_per_step_regions', 'loss', 'timed_seconds']})
print('e2e', d['e2e']); print('roofline', {k: d['roofline'].get(k) for k in ['kernel', 'achieved', 'peak', 'frac', 'traffic', 'share_of_step', 'whole_step_tflops']})
for k in d['roofline'].get('kernels', []): print('   ', k['kernel'], round(k['ms_per_step'], 3), round(k['tflops']), round(k['frac_of_sustained_peak'], 3))
print('cpu', d['cpu_baseline']); print('scatter', {k: v for k, v in d['scatter'].items() if 'frac' in k or 'gbs' in k or k == 'ms'})
print('stages', d['stages_ms'])
PY
$T 400 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline --no-scatter --min-seconds 1 > gpurun_out/bench_fp32.log 2> gpurun_out/bench_fp32.err; echo "fp32 rc=$?"; cut -c1-160 gpurun_out/bench_fp32.log
$T 200 python bench.py --scatter-only --steps 5 > gpurun_out/bench_scatter.log 2> gpurun_out/bench_scatter.err; echo "scatter rc=$?"; cut -c1-200 gpurun_out/bench_scatter.log
$T 400 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-200 gpurun_out/bench_reference.log
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 5 > gpurun_out/launch_summary.txt 2>&1; head -44 gpurun_out/launch_summary.txt
