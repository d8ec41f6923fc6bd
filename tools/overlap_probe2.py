"""Per-family kernel durations (CUDA events on the main stream) with and without the side-stream actv prefetch."""
import os, sys, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
eng = net.engine()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1)]
for ov in (True, "force", True, "force"):
    eng.actv_overlap = ov
    with torch.no_grad():
        for _ in range(3):
            net(*inp)
        torch.cuda.synchronize()
        eng.profile = []
        for _ in range(5):
            net(*inp)
        torch.cuda.synchronize()
        recs, eng.profile = eng.profile, None
    t = collections.defaultdict(float); c = collections.Counter()
    for r in recs:
        t[r["family"]] += r["e0"].elapsed_time(r["e1"]); c[r["family"]] += 1
    print("overlap=%s" % (ov == "force"))
    for f in sorted(t, key=lambda f: -t[f])[:5]:
        print("   %-28s %4d launches  %.1f us each" % (f, c[f] // 5, t[f] / c[f] * 1e3))
