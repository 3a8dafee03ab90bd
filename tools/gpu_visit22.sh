#!/bin/bash
# Round 2, visit 22 (1 GPU): the per-layer table (cuDNN bf16 / TF32 against the tcgen05 kernels) on the final build.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 600 python tools/conv_layer_table.py --out gpurun_out/conv_layer_table.txt > gpurun_out/conv_layer_table.log 2>&1; echo "conv table rc=$?"; tail -22 gpurun_out/conv_layer_table.log | cut -c1-230
