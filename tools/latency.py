"""Per-frame latency of the public call at batch 1 (BASELINE configs[0] shape 64x64 and configs[4] shape 135x240):
host issue time vs device time."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
for (B, h, w) in ((1, 64, 64), (1, 135, 240), (4, 135, 240)):
    inp = [t.cuda() for t in synthetic_inputs(B, h, w, scale=8, seed=1)]
    with torch.no_grad():
        for _ in range(5):
            net(*inp)
        torch.cuda.synchronize()
        n = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); e0.record()
        for _ in range(n):
            net(*inp)
        e1.record(); th = time.time() - t0
        torch.cuda.synchronize()
    print("B=%d %dx%d: host issue %.3f ms/call, device %.3f ms/call -> %.1f frames/s" % (
        B, h, w, th / n * 1e3, e0.elapsed_time(e1) / n, B * n / (e0.elapsed_time(e1) * 1e-3)), flush=True)
