#!/bin/bash
# Round 2, visit 16 (1 GPU): where do the general implicit-GEMM kernel's warps wait?  One --set full capture with source
# counters of four launches at benchmark shapes; the report is exported to text on the box (the .ncu-rep stays in /tmp).
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 120 python tools/prof_general_igemm.py > gpurun_out/prof_general_times.txt 2>&1; cat gpurun_out/prof_general_times.txt | tail -5
# launches of interest: third repetition (skip 2 x 4 igemm launches)
$T 900 ncu --set full --clock-control none --import-source on -k regex:k_conv_igemm -s 8 -c 4 -o /tmp/prof_general \
  python tools/prof_general_igemm.py > gpurun_out/ncu_general.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_general.ncu-rep --page source --print-source sass --csv > gpurun_out/general_source_sass.csv 2> gpurun_out/export.err
ncu -i /tmp/prof_general.ncu-rep --page source --print-source cuda --csv > gpurun_out/general_source_cuda.csv 2>> gpurun_out/export.err
ncu -i /tmp/prof_general.ncu-rep --page raw --csv > gpurun_out/general_raw.csv 2>> gpurun_out/export.err
ncu -i /tmp/prof_general.ncu-rep --page details > gpurun_out/general_details.txt 2>> gpurun_out/export.err
ls -la gpurun_out/general_* | cut -c1-120
