#!/bin/bash
# Round 2, first GPU visit: the whole GPU suite at HEAD (stale-weight fix, reference-trajectory tests, reference modules on the
# drop-in ext), smoke, the k_conv_wgrad_x validation + A/B, bench lines (bf16 + fp32 parity), ncu launch list, the reference
# on the GPU (bar to beat) and the per-layer cuDNN table.  Outputs: gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -45 gpurun_out/pytest_gpu.log
$T 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
$T 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1])
    print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks', 'flow_err', 'ms_per_step_regions', 'loss']})
    print('e2e', d['e2e']); print('roofline', {k: d['roofline'].get(k) for k in ['kernel', 'achieved', 'peak', 'frac', 'traffic', 'share_of_step']})
    for k in d['roofline'].get('kernels', []): print('   ', k)
    print('cpu', d['cpu_baseline']); print('scatter', d['scatter'])
    print('stages', d['stages_ms'])
except Exception as e: print('bench parse failed', e)
PY
$T 600 python bench.py --steps 5 --warmup 3 --precision fp32 --no-cpu-baseline --no-scatter --min-seconds 1 > gpurun_out/bench_fp32.log 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?"; cut -c1-400 gpurun_out/bench_fp32.log
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 4 > gpurun_out/launch_summary.txt 2>&1; head -40 gpurun_out/launch_summary.txt
$T 700 python tools/ref_gpu_bench.py --batch 16 --out gpurun_out/ref_gpu.json > gpurun_out/ref_gpu.log 2>&1 || $T 500 python tools/ref_gpu_bench.py --batch 8 --out gpurun_out/ref_gpu.json >> gpurun_out/ref_gpu.log 2>&1
echo "ref gpu rc=$?"; tail -2 gpurun_out/ref_gpu.log | cut -c1-600
$T 500 python tools/ref_gpu_bench.py --batch 16 --amp --out gpurun_out/ref_gpu.json >> gpurun_out/ref_gpu.log 2>&1; tail -1 gpurun_out/ref_gpu.log | cut -c1-400
$T 600 python tools/conv_layer_table.py --out gpurun_out/conv_layer_table.txt > gpurun_out/conv_layer_table.log 2>&1; echo "conv table rc=$?"; cat gpurun_out/conv_layer_table.log | tail -25
