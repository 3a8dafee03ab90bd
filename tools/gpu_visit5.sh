#!/bin/bash
# Round 2, visit 5 (1 GPU): full suite at HEAD (hard voxelisation, chamfer / seflowLoss, 3-D voxels, feed fix), smoke, ncu --set
# full probe of the 1x1 / stride-2 data-gradient launches where cuDNN wins, warm --set full capture of the main tensor-core
# kernels (tensor-pipe utilisation), scatter sweep with a global warm-up.
mkdir -p gpurun_out
T="timeout --kill-after=15 -s TERM"
$T 1500 python -m pytest tests -m gpu -q -rf --durations=6 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "^E  |^FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-250 | tail -40
$T 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
$T 200 python tools/igemm_1x1_probe.py > gpurun_out/igemm_probe.log 2>&1; cat gpurun_out/igemm_probe.log
$T 600 ncu --set full --import-source on --clock-control none -k regex:k_conv_igemm --launch-skip 4 --launch-count 12 \
  -o gpurun_out/igemm_probe -f python tools/igemm_1x1_probe.py > gpurun_out/igemm_probe_ncu.log 2>&1; echo "ncu probe rc=$?"
DFB_PROFILE=1 $T 900 ncu --set full --cache-control none --clock-control none -k regex:'k_conv_igemm_halo|k_conv_wgrad_x|k_conv_wgrad_halo|k_gru_fused' \
  --launch-skip 120 --launch-count 40 -o gpurun_out/main_warm_full -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-scatter --no-e2e > gpurun_out/ncu_main_warm.log 2>&1; echo "ncu main rc=$?"
$T 600 python tools/scatter_sweep.py --out gpurun_out/scatter_sweep.txt > gpurun_out/scatter_sweep.log 2>&1; echo "sweep rc=$?"; tail -8 gpurun_out/scatter_sweep.log
ls -la gpurun_out/*.ncu-rep
