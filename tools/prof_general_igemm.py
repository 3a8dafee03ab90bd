"""Single launches of the general implicit-GEMM kernel (and one row-pair launch) at benchmark shapes, for an
`ncu --set full --import-source on` capture (tools/gpu_visit16.sh): which instruction the epilogue / producer / MMA warps
wait on when a tile has one or two K steps."""
import sys

import torch

sys.path.insert(0, ".")
from deflow_b200 import conv  # noqa: E402

DEV = "cuda"
torch.manual_seed(0)


def bf(*s):
    return torch.randn(*s, device=DEV).to(torch.bfloat16)


B = 16
cases = {
    # dec3.u3 forward: 1x1, 32 -> 64 at 512^2  (k_conv_igemm<64,32>, one K step per tile)
    "u3": lambda: conv.conv2d_forward([x32], wf_u3, b64, 64, 1, 1),
    # dec3.u1 data gradient: 1x1, 64 -> 128 at 256^2 (k_conv_igemm<128,64>, one K step per tile)
    "u1d": lambda: conv.conv2d_dgrad(gy_u1, wd_u1, 256, 256, 128, 128, 0, 1, 1),
    # enc1.0 data gradient: 3x3 stride 2, 64 -> 32, 256^2 -> 512^2 (k_conv_igemm<32,64>, parity-merged)
    "e1d": lambda: conv.conv2d_dgrad(gy_e1, wd_e1, 512, 512, 32, 32, 0, 3, 2),
    # enc1.1 forward: 3x3 64 -> 64 at 256^2 (row-pair kernel)
    "pair": lambda: conv.conv2d_forward([x64], wf_p, b64, 64, 3, 1),
    # enc2.1 forward: 3x3 128 -> 128 at 128^2 (k_conv_igemm_halo<128>); enc3.1: 3x3 256 -> 256 at 64^2 (halo<256>)
    "h128": lambda: conv.conv2d_forward([x128], wf_h128, b128, 128, 3, 1),
    "h256": lambda: conv.conv2d_forward([x256], wf_h256, b256, 256, 3, 1),
    # weight gradients of the same layers and of enc1.1 (k_conv_wgrad_halo<128>, <256>, k_conv_wgrad_x)
    "wh128": lambda: conv.conv2d_wgrad([x128], gy128, 3, 1),
    "wh256": lambda: conv.conv2d_wgrad([x256], gy256, 3, 1),
    "wx": lambda: conv.conv2d_wgrad([x64], gy64, 3, 1),
}
x32 = bf(B, 512, 512, 32)
b64 = torch.randn(64, device=DEV)
wf_u3, _ = conv.pack_weights(torch.randn(64, 32, 1, 1, device=DEV) / 6)
gy_u1 = bf(B, 256, 256, 64)
_, wd_u1 = conv.pack_weights(torch.randn(64, 128, 1, 1, device=DEV) / 11)
gy_e1 = bf(B, 256, 256, 64)
_, wd_e1 = conv.pack_weights(torch.randn(64, 32, 3, 3, device=DEV) / 17)
x64 = bf(B, 256, 256, 64)
wf_p, _ = conv.pack_weights(torch.randn(64, 64, 3, 3, device=DEV) / 24)
x128, gy128, b128 = bf(B, 128, 128, 128), bf(B, 128, 128, 128), torch.randn(128, device=DEV)
x256, gy256, b256 = bf(B, 64, 64, 256), bf(B, 64, 64, 256), torch.randn(256, device=DEV)
gy64 = bf(B, 256, 256, 64)
wf_h128, _ = conv.pack_weights(torch.randn(128, 128, 3, 3, device=DEV) / 34)
wf_h256, _ = conv.pack_weights(torch.randn(256, 256, 3, 3, device=DEV) / 48)

which = sys.argv[1:] or list(cases)
for rep in range(3):
    for k in which:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        cases[k]()
        ev1.record()
        torch.cuda.synchronize()
        if rep == 2:
            print(k, round(ev0.elapsed_time(ev1) * 1000, 1), "us")
