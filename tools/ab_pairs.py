"""A/B of the CTA-pair kernels inside ONE process (same box, same thermal state): device time per B=64 forward with
dasr_set_sean_pair(0 / 1), alternating."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200 import _lib as L
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = 30
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
sets = [[t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=s)] for s in (1, 2, 3)]
lib = L.load()
MODES = [int(m) for m in os.environ.get("MODES", "0,7,6,4,2").split(",")]
res = {m: [] for m in MODES}
with torch.no_grad():
    for rep in range(4):
        for on in MODES:
            L.check(lib.dasr_set_sean_pair(on))
            for i in range(4):
                net(*sets[i % 3])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(K):
                net(*sets[i % 3])
            e1.record()
            torch.cuda.synchronize()
            res[on].append(e0.elapsed_time(e1) / K)
for on in MODES:
    print("pair mask %d: ms/step %s  median %.3f" % (on, ["%.3f" % v for v in res[on]], sorted(res[on])[len(res[on]) // 2]))
