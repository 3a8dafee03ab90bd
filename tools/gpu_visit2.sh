#!/bin/bash
# Round 2, GPU visit 2: full suite (feed / eval-metric / SyncBN / reference-module tests added), default bench with the
# cross-shift weight-gradient kernel on, complete per-layer cuDNN table, warm-cache ncu metrics of the tensor-core kernels,
# scatter sweep + DRAM traffic of the scatter microbench.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -40
$T 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1])
    print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks', 'flow_err', 'ms_per_step_regions', 'loss']})
    print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
    for k in d['roofline'].get('kernels', []): print('   ', k['kernel'], round(k['ms_per_step'], 3), round(k['tflops']), round(k['frac_of_sustained_peak'], 3))
    print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind']); print('scatter', {k: v for k, v in d['scatter'].items() if 'frac' in k or 'gbs' in k or k == 'ms'})
except Exception as e: print('bench parse failed', e)
PY
$T 600 python tools/conv_layer_table.py --out gpurun_out/conv_layer_table.txt > gpurun_out/conv_layer_table.log 2>&1; echo "conv table rc=$?"; tail -22 gpurun_out/conv_layer_table.log
$T 600 python tools/scatter_sweep.py --out gpurun_out/scatter_sweep.txt > gpurun_out/scatter_sweep.log 2>&1; echo "sweep rc=$?"; tail -8 gpurun_out/scatter_sweep.log
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
DFB_PROFILE=1 $T 900 ncu --cache-control none --clock-control none --metrics $M -k regex:'k_conv_igemm|k_conv_wgrad|k_gru_fused' \
  --launch-skip 200 --launch-count 220 --csv --log-file gpurun_out/warm_metrics.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_warm.log 2>&1; echo "ncu warm rc=$?"
python tools/ncu_metrics_summary.py gpurun_out/warm_metrics.csv > gpurun_out/warm_metrics_summary.txt 2>&1; head -60 gpurun_out/warm_metrics_summary.txt
$T 400 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --csv --log-file gpurun_out/scatter_metrics.csv python bench.py --scatter-only --steps 1 > gpurun_out/ncu_scatter.log 2>&1; echo "ncu scatter rc=$?"
python tools/ncu_metrics_summary.py gpurun_out/scatter_metrics.csv > gpurun_out/scatter_metrics_summary.txt 2>&1; head -50 gpurun_out/scatter_metrics_summary.txt
