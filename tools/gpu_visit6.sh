#!/bin/bash
# Round 2, visit 6 (1 GPU): suite at HEAD (chamfer contraction fix, 1x1 bias-gradient sums by matvec), bench, scatter sweep,
# warm --set full capture of the main tensor-core kernels exported to text on the box (gpurun_out must stay < 64 MiB).
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -30
$T 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads([x for x in open('gpurun_out/bench.log') if x.startswith('{')][-1])
    print({k: d[k] for k in ['value', 'ms_per_step', 'gpu_launches', 'clocks', 'ms_per_step_regions', 'loss']})
    print('e2e', d['e2e']['value'], 'flow_err', d['flow_err'])
    for k in d['roofline'].get('kernels', []): print('   ', k['kernel'], round(k['ms_per_step'], 3), round(k['tflops']), round(k['frac_of_sustained_peak'], 3))
except Exception as e: print('bench parse failed', e)
PY
$T 600 python tools/scatter_sweep.py --out gpurun_out/scatter_sweep.txt > gpurun_out/scatter_sweep.log 2>&1; echo "sweep rc=$?"
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
python tools/summarize_launches.py gpurun_out/launches.csv 4 > gpurun_out/launch_summary.txt 2>&1; head -34 gpurun_out/launch_summary.txt
DFB_PROFILE=1 $T 900 ncu --set full --cache-control none --clock-control none -k regex:'k_conv_igemm_halo|k_conv_wgrad_x|k_conv_wgrad_halo|k_gru_fused' \
  --launch-skip 120 --launch-count 36 -o /tmp/main_warm_full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_main_warm.log 2>&1; echo "ncu main rc=$?"
ncu -i /tmp/main_warm_full.ncu-rep --page raw --csv > /tmp/main_warm_full.csv 2>/dev/null
python - <<'PY'
import csv, re
rows = list(csv.reader(open('/tmp/main_warm_full.csv')))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum"]
idx = {h: i for i, h in enumerate(hdr)}
with open('gpurun_out/ncu_full_warm_main.txt', 'w') as f:
    f.write("# ncu --set full --cache-control none --clock-control none (warm caches), bench.py --steps 1 --warmup 1 (config 2), r02 visit 6\n")
    for r in rows[2:]:
        f.write(f"---- {re.sub(r'^void ', '', r[idx['Kernel Name']])} grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n")
        for m in want:
            if m in idx:
                f.write(f"   {m:82s} {r[idx[m]]:>14s} {units[idx[m]]}\n")
print(open('gpurun_out/ncu_full_warm_main.txt').read()[:3000])
PY
du -sh gpurun_out
