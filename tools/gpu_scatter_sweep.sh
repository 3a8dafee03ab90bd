#!/bin/bash
# Scatter-path variant sweep (diagnostic env switches) + ncu captures.  Outputs under gpurun_out/.
# Historical: run at commit 60d604d (results: profiles/r01_scatter_variant_sweep.txt); the DFB_COMPACT_LOOP / DFB_FILL_U /
# DFB_MARK_PRECHECK / DFB_PFN_POINTS switches it sets selected variants that lost and have been removed since.
mkdir -p gpurun_out
T="timeout -s KILL"
run() {  # name, env...
  name=$1; shift
  env "$@" $T 200 python bench.py --scatter-only --steps 10 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$name', round(d['value']), 'GB/s', d['scatter']['ms'])
"
}
$T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_ext.py -x -q > gpurun_out/pytest_idx.log 2>&1; echo "pytest idx rc=$?"; tail -2 gpurun_out/pytest_idx.log
DFB_COMPACT_LOOP=1 DFB_PFN_POINTS=tile DFB_FILL_U=4 DFB_MARK_PRECHECK=0 $T 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_ext.py -x -q > gpurun_out/pytest_idx2.log 2>&1; echo "pytest idx2 rc=$?"; tail -2 gpurun_out/pytest_idx2.log
run default DFB_X=0
run compact_loop DFB_COMPACT_LOOP=1
run no_precheck DFB_MARK_PRECHECK=0
run fill_u2 DFB_FILL_U=2
run fill_u4 DFB_FILL_U=4
run points_tile DFB_PFN_POINTS=tile
$T 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/scatter_launches.csv python bench.py --scatter-only --steps 1 > gpurun_out/scatter_ncu.log 2>&1; echo "ncu scatter rc=$?"
$T 400 ncu --set full --import-source on --clock-control none -k regex:'k_compact|k_pfn_points|k_fill_csr|k_mark_points|k_pfn_bwd$|k_pillar_mean' --launch-skip 12 --launch-count 6 \
  -o gpurun_out/scatter_full -f python bench.py --scatter-only --steps 1 > gpurun_out/scatter_full.log 2>&1; echo "ncu full scatter rc=$?"
DFB_PROFILE=1 $T 500 ncu --set full --import-source on --clock-control none -k regex:'k_conv_wgrad_halo|k_gru_fused' --launch-skip 30 --launch-count 6 \
  -o gpurun_out/wgrad_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/wgrad_full.log 2>&1; echo "ncu full wgrad rc=$?"
ls -la gpurun_out/*.ncu-rep
