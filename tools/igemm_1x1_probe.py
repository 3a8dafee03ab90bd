"""Probe for an ncu --set full capture: the 1x1 / stride-2 data gradients on which cuDNN beats k_conv_igemm
(profiles/r02_conv_layer_table.txt): dec2.u1 (gy [16,128,128,128] -> gx 256 ch), dec3.u1 (gy [16,256,256,64] -> 128 ch),
dec2.u3 two-source (gy [16,256,256,128] -> 64 + 64 ch), enc3.0 stride-2 (gy [16,64,64,256] -> [16,128,128,128]).
Prints CUDA-event times; run under `ncu -k regex:k_conv_igemm`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deflow_b200 import conv as tc  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B = 16
cases = [("dec2.u1 dgrad", 128, 128, [256], 1, 1), ("dec3.u1 dgrad", 256, 64, [128], 1, 1), ("dec2.u3 dgrad2", 256, 128, [64, 64], 1, 1),
         ("enc3.0 dgrad s2", 128, 256, [128], 3, 2)]
for name, H, cout, cins, k, stride in cases:
    cin = sum(cins)
    Ho = (H + 2 * (k // 2) - k) // stride + 1
    w = torch.randn(cout, cin, k, k, device=dev) * 0.05
    _, wd = tc.pack_weights(w, True, False)
    gy = torch.randn(B, Ho, Ho, cout, device=dev).to(torch.bfloat16)
    ts = []
    for it in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if len(cins) == 2:
            tc.conv2d_dgrad_two(gy, wd, H, H, cins[0], cins[1], cin, k)
        else:
            tc.conv2d_dgrad(gy, wd, H, H, cin, cin, 0, k, stride, colsum=(it % 2 == 1))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, ["%.3f" % t for t in ts], flush=True)
