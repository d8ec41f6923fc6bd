"""Weight-gradient kernel at the training shapes (B=16): timing, and a driver for ncu --set full."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
def case(B, H, W, Cin, Cout, kh=3, kw=3, n=10):
    dy = torch.randn(B, H, W, Cout, device=dev).to(torch.bfloat16)
    x = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16)
    dw = torch.zeros(Cout, kh * kw * Cin, device=dev)
    f = lambda: L.conv_wgrad(dy, x, dw, kh, kw)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    fl = 2.0 * B * H * W * Cout * Cin * kh * kw
    print("wgrad B%d %dx%d %d->%d %dx%d: %.1f us  %.0f TFLOP/s" % (B, H, W, Cin, Cout, kh, kw, us, fl / us / 1e6), flush=True)
case(16, 64, 64, 128, 128)
case(16, 64, 64, 64, 64)
case(16, 64, 64, 32, 128)
case(16, 256, 256, 32, 32)
case(16, 256, 256, 32, 128)
case(16, 512, 512, 32, 32, 9, 1)
