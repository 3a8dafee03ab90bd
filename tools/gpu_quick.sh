#!/bin/bash
# Short GPU-box visit: tensor-core tests, bench line, launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
$T 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train_step.py -q > gpurun_out/pytest_tc.log 2>&1; echo "pytest tc rc=$?" | tee -a gpurun_out/pytest_tc.log
tail -6 gpurun_out/pytest_tc.log
$T 600 python bench.py --no-scatter --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench.log; echo
DFB_PROFILE=1 $T 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
