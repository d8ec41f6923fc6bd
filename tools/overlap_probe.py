"""Forward step time (B=64, x8) with / without the side-stream actv prefetch; and B=16."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
eng = net.engine()
for B in (64, 16):
    inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1)]
    ref = None
    for ov in (False, True, False, True):
        eng.actv_overlap = ov
        with torch.no_grad():
            for _ in range(3):
                out = net(*inp)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n):
                out = net(*inp)
            e1.record()
            torch.cuda.synchronize()
        if ref is None:
            ref = out.clone()
        print("B=%d overlap=%d: %.3f ms/step  (%.0f frames/s)  max|diff vs first|=%.3g" % (
            B, ov, e0.elapsed_time(e1) / n, B * n / e0.elapsed_time(e1) * 1e3, (out - ref).abs().max().item()), flush=True)
