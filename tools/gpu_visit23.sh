#!/bin/bash
# Round 2, visit 23 (1 GPU): warm --set full captures of one launch of each 3x3 tensor-core kernel at benchmark shapes
# (tensor-pipe utilisation, DRAM bytes, duration against the CUDA-event time printed by the same script).
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
K="h128 h256 wh128 wh256 wx pair"
$T 120 python tools/prof_general_igemm.py $K > gpurun_out/prof_tc_times.txt 2>&1; tail -6 gpurun_out/prof_tc_times.txt
$T 900 ncu --set full --clock-control none --cache-control none -k regex:"k_conv_igemm_halo|k_conv_wgrad" -s 12 -c 6 -o /tmp/prof_tc \
  python tools/prof_general_igemm.py $K > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_tc.ncu-rep --page details > gpurun_out/tc_details.txt 2> gpurun_out/export.err
ncu -i /tmp/prof_tc.ncu-rep --page raw --csv > gpurun_out/tc_raw.csv 2>> gpurun_out/export.err
grep -E "^  [a-z].*k_conv|Duration|highest-utilized|DRAM Throughput|Memory Throughput  " gpurun_out/tc_details.txt | cut -c1-160
