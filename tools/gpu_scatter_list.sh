#!/bin/bash
# ncu launch list of the config-5 scatter microbench (cold-cache, serialised per-launch times)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/scatter_launches.csv python bench.py --scatter-only --steps 1 > gpurun_out/scatter_ncu.log 2>&1
echo "rc=$?"
