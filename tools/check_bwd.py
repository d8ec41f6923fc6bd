"""GPU bring-up checks of the backward kernels against torch autograd (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from depth_aware_endoscopy_sr_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
allok = True


def report(name, got, ref, tol):
    global allok
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(scale, 1e-6)
    allok &= ok
    print("%-52s max|err|=%.4g (ref max %.3g) %s" % (name, err, scale, "PASS" if ok else "FAIL"), flush=True)


def bf(x):
    return x.to(torch.bfloat16).float()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def wgrad_case(B, H, W, Cin, Cout, kh=3, kw=3, time=False):
    x = torch.randn(B, Cin, H, W, device=dev)
    dy = torch.randn(B, Cout, H, W, device=dev)
    xa, dya = nhwc(x).to(torch.bfloat16), nhwc(dy).to(torch.bfloat16)
    dw = torch.zeros(Cout, kh * kw * Cin, device=dev)
    L.conv_wgrad(dya, xa, dw, kh, kw)
    torch.cuda.synchronize()
    w = torch.zeros(Cout, Cin, kh, kw, device=dev, requires_grad=True)
    F.conv2d(bf(x), w, None, padding=(kh // 2, kw // 2)).backward(bf(dy))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin)      # [o][(t,u),i]
    report("wgrad B%d %dx%d %d->%d k%dx%d" % (B, H, W, Cin, Cout, kh, kw), dw, ref, 2e-3)
    if time:
        ms = timeit(lambda: L.conv_wgrad(dya, xa, dw, kh, kw))
        print("     %.3f ms  %.0f TFLOP/s" % (ms, 2.0 * B * H * W * Cout * Cin * kh * kw / ms / 1e9), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["wgrad"]
    if "wgrad" in which:
        wgrad_case(2, 16, 16, 64, 64)
        wgrad_case(2, 64, 64, 64, 64)
        wgrad_case(2, 64, 64, 128, 128)
        wgrad_case(1, 32, 48, 32, 32)
        wgrad_case(1, 24, 40, 64, 32)
        wgrad_case(1, 24, 40, 32, 64)
        wgrad_case(1, 31, 31, 256, 256)
        wgrad_case(1, 31, 31, 128, 256)
        wgrad_case(2, 64, 64, 64, 256)
        wgrad_case(1, 135, 240, 64, 64)
        wgrad_case(1, 256, 256, 32, 128)
        wgrad_case(1, 64, 64, 32, 32, kh=9, kw=1)
        wgrad_case(1, 128, 256, 32, 32, kh=9, kw=1)
        wgrad_case(16, 64, 64, 64, 64, time=True)
        wgrad_case(16, 64, 64, 128, 128, time=True)
        wgrad_case(16, 256, 256, 32, 32, time=True)
    print("ALL PASS" if allok else "SOME FAILED")
