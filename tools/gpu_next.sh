#!/bin/bash
# First GPU visit of the next round: validate the experimental cross-shift weight-gradient kernel (k_conv_wgrad_x,
# DFB_WGRAD_X=1, written without GPU time at the end of round 1) and A/B it against the default.  Outputs: gpurun_out/.
mkdir -p gpurun_out
T="timeout -s KILL"
# 1. the skipped parity test, forced to run
$T 300 python - <<'PY' > gpurun_out/wgrad_x_test.log 2>&1
import re, subprocess, sys
src = open("tests/test_gpu_tensorcore.py").read()
src = re.sub(r'@pytest\.mark\.skip\(reason="k_conv_wgrad_x.*?\)\n', "", src, flags=re.S)
open("/tmp/test_wgrad_x.py", "w").write(src)
sys.exit(subprocess.call([sys.executable, "-m", "pytest", "/tmp/test_wgrad_x.py", "-q", "-k", "cross_shift",
                          "--rootdir", ".", "-c", "/dev/null", "-p", "no:cacheprovider"], env={**__import__("os").environ, "PYTHONPATH": ".:tests"}))
PY
echo "wgrad_x test rc=$?"; tail -5 gpurun_out/wgrad_x_test.log
# 2. A/B bench
for x in 0 1; do
  DFB_WGRAD_X=$x $T 300 python bench.py --no-scatter --no-cpu-baseline --no-e2e > gpurun_out/bench_wx$x.log 2> gpurun_out/bench_wx$x.err
  python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bench_wx$x.log') if l.startswith('{')][-1])
k = [r for r in d['roofline']['kernels'] if 'wgrad_halo<64>' in r['kernel']]
print('DFB_WGRAD_X=$x', round(d['value'], 1), 'pairs/s', round(d['ms_per_step'], 3), 'ms', k)
PY
done
