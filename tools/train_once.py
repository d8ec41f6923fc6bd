"""Run a few x8 training steps (B=16, 64x64 LR); the last one sits between cudaProfilerStart/Stop so that
`ncu --profile-from-start off` lists exactly one step.  Also prints host wall time vs device time per step."""
import sys, os, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().train()
step = dasr.TrainStep(net, graph=(len(sys.argv) > 2 and sys.argv[2] == 'graph'))
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1, with_gt=True)]
for _ in range(3):
    step(*inp)
torch.cuda.synchronize()
for phase in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    n = 5
    for _ in range(n):
        step(*inp)
    e1.record()
    t_host = time.time() - t0
    torch.cuda.synchronize()
    print("steps: host-issue %.2f ms/step, device %.2f ms/step" % (t_host / n * 1e3, e0.elapsed_time(e1) / n))
torch.cuda.cudart().cudaProfilerStart()
step(*inp)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
