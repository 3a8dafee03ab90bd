"""Per-tap error of the weight-gradient kernels (diagnostic)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deflow_b200 import conv

for (n, H, W, cin, cout) in ((2, 32, 32, 64, 64), (1, 32, 32, 128, 128), (1, 16, 16, 256, 256)):
    torch.manual_seed(0)
    x = torch.randn(n, H, W, cin, device="cuda").to(torch.bfloat16)
    gy = torch.randn(n, H, W, cout, device="cuda").to(torch.bfloat16)
    w = torch.zeros(cout, cin, 3, 3, device="cuda", requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, padding=1)
    y.backward(gy.float().permute(0, 3, 1, 2))
    gw = conv.conv2d_wgrad([x], gy, 3, 1)
    torch.cuda.synchronize()
    ref = w.grad
    print(f"cin {cin} cout {cout}: per-tap max err / ref absmax")
    for t in range(9):
        e = (gw[:, :, t // 3, t % 3] - ref[:, :, t // 3, t % 3]).abs().max()
        # also compare against every other reference tap to spot mix-ups
        best = min(range(9), key=lambda u: float((gw[:, :, t // 3, t % 3] - ref[:, :, u // 3, u % 3]).abs().max()))
        print(f"   tap {t}: err {float(e):9.4f} (ref {float(ref[:, :, t // 3, t % 3].abs().max()):7.2f})  closest ref tap {best}", flush=True)
