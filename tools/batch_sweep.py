"""Forward time of the x8 64x64 network against the batch size (run under gpurun).

Question: the trunk kernels run 16 tiles of 256 pixels per image on 148 SMs (74 CTA pairs); is a 64-frame batch better
served as ONE call (6.92 -> 7 waves, every tensor streamed through HBM) or as sub-batches whose working set stays in
the 126 MB L2 (37 frames = exactly 4 waves, 27 frames = 2.92 -> 3 waves)?
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

torch.manual_seed(0)
net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16, nf=64, depthRangeNum=10).cuda().eval()
res = {}
sizes = [int(x) for x in (sys.argv[1:] or "64 37 27 32 18 19 9 10 46".split())]
for B in sizes:
    lq, depth, masks = [t.cuda() for t in synthetic_inputs(B, 64, 64, seed=3)]
    with torch.no_grad():
        for _ in range(5):
            net(lq, depth, masks)
        torch.cuda.synchronize()
        ts = []
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            n = 20
            for _ in range(n):
                net(lq, depth, masks)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / n)
    ms = sorted(ts)[1]
    res[B] = ms
    print("B=%3d: %.3f ms/forward  %.1f us/frame  -> %.0f frames/s   (runs %s)" % (
        B, ms, ms / B * 1e3, B / ms * 1e3, " ".join("%.3f" % t for t in ts)), flush=True)
if 64 in res:
    for combo in ((37, 27), (32, 32), (18, 19, 27), (46, 18)):
        if all(c in res for c in combo):
            print("64 frames as %s: %.3f ms against %.3f ms in one call" % (
                "+".join(map(str, combo)), sum(res[c] for c in combo), res[64]))
