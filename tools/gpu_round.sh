#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench line, scatter microbench, ncu launch list + full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
$T 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
$T 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
head -c 1500 gpurun_out/bench.log; echo
$T 300 python bench.py --scatter-only > gpurun_out/scatter.log 2> gpurun_out/scatter.err; echo "scatter rc=$?"
cut -c1-300 gpurun_out/scatter.log
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
DFB_PROFILE=1 $T 500 ncu --set full --import-source on --clock-control none -k regex:'k_conv_igemm_halo_pair|k_gru_fused_bwd|k_conv_igemm_halo<128>|k_bn_gelu_bwd_reduce_s' --launch-skip 60 --launch-count 8 \
  -o gpurun_out/top_full -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/top_full.log 2>&1; echo "ncu full rc=$?"
$T 300 ncu --set full --clock-control none -k regex:'k_pfn_points|k_compact|k_fill_csr|k_pfn_bwd$' --launch-skip 8 --launch-count 4 \
  -o gpurun_out/scatter_full2 -f python bench.py --scatter-only --steps 1 > gpurun_out/scatter_full2.log 2>&1; echo "ncu full scatter rc=$?"
ls -la gpurun_out/*.ncu-rep
