#!/bin/bash
# One GPU-box visit: parity tests, bench line, scatter microbench (+ ncu launch lists).  Outputs under gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 300 python -m pytest tests/test_gpu_tensorcore.py -x -q -k "conv_forward" > gpurun_out/pytest_conv.log 2>&1; rc=$?; echo "pytest conv rc=$rc"
tail -15 gpurun_out/pytest_conv.log
if [ $rc -ne 0 ]; then export DFB_HALO_PAIR=0; echo "row-pair kernel disabled for the rest of this visit"; fi
$T 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
$T 300 python bench.py --scatter-only > gpurun_out/scatter.log 2> gpurun_out/scatter.err; echo "scatter rc=$?"
cut -c1-300 gpurun_out/scatter.log
$T 600 python bench.py --no-scatter > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
head -c 1800 gpurun_out/bench.log; echo
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
