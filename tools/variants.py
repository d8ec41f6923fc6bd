"""Throughput of the x4 / x2 variants (BASELINE configs[3]) and the 1080p frame (configs[4]): forward and training."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

def tm(fn, n):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for (scale, which, latent, B, h, w, Bt) in ((8, range(14), 256, 64, 64, 64, 16), (4, range(14), 256, 16, 128, 128, 8),
                                            (2, range(16), 32, 4, 256, 256, 2), (8, range(14), 256, 8, 135, 240, 0)):
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(which), scale=scale, nb=16, depth_latent_ch=latent).cuda().eval()
    inp = [t.cuda() for t in synthetic_inputs(B, h, w, scale=scale, seed=1)]
    with torch.no_grad():
        ms = tm(lambda: net(*inp), 10)
    line = "x%d  B=%d %dx%d -> %dx%d: forward %.2f ms = %.0f frames/s (%.1f Mpixel HR/s)" % (
        scale, B, h, w, h * scale, w * scale, ms, B / ms * 1e3, B * h * w * scale * scale / ms / 1e3)
    if Bt:
        net.train()
        step = dasr.TrainStep(net, graph=True)
        tin = [t.cuda() for t in synthetic_inputs(Bt, h, w, scale=scale, seed=2, with_gt=True)]
        ms_t = tm(lambda: step(*tin), 6)
        line += "   | train B=%d: %.2f ms/step = %.0f images/s" % (Bt, ms_t, Bt / ms_t * 1e3)
    print(line, flush=True)
    del net
    torch.cuda.empty_cache()
