"""Run the memory-bound SEAN input kernels (actv, K-DYN apply) at the bench shape (B=64, 64x64) for ncu / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
lib = L.load(); s = L.stream_ptr()
lq, depth, masks = [t.to(dev) for t in synthetic_inputs(B, 64, 64, seed=0)]
K, H, W, nf2 = 10, 64, 64, 128
labels = torch.empty(B, H, W, device=dev, dtype=torch.uint8); flag = torch.zeros(1, device=dev, dtype=torch.int32)
L.check(lib.dasr_mask_labels(L.ptr(masks), L.ptr(labels), L.ptr(flag), B, K, H, W, s))
w = torch.randn(nf2, 9, device=dev); b = torch.randn(nf2, device=dev)
actv = torch.empty(B, H, W, nf2, device=dev, dtype=torch.bfloat16)
table = torch.randn(B, K, 9, nf2, device=dev).to(torch.bfloat16)
gbs = torch.empty(B, H, W, nf2, device=dev, dtype=torch.bfloat16)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
fa = lambda: L.check(lib.dasr_actv_fwd(L.ptr(depth), L.ptr(w), L.ptr(b), L.ptr(actv), B, H, W, nf2, 0, s))
fd = lambda: L.check(lib.dasr_dynconv_fwd(L.ptr(table), L.ptr(labels), L.ptr(masks), L.ptr(flag), L.ptr(gbs), B, K, H, W, nf2, s))
mb = actv.numel() * 2 / 1e6
ua, ud = t(fa), t(fd)
print("B=%d actv %.1f us (%.0f GB/s)   dynconv %.1f us (%.0f GB/s)   [%.1f MB stored each]" % (B, ua, mb / ua * 1e3, ud, mb / ud * 1e3, mb))
