"""Where does the captured x8 training step (B=16) spend its time?  (run under gpurun)

Measures the CUDA-graph step (a) as shipped, (b) with every weight-gradient launch removed (the tool replaces
Tape.wgrad by a no-op: the numbers are wrong, the main chain is unchanged), (c) with the whole backward on one stream,
and (d) the forward alone -- to tell whether the step is bound by the main chain (latency) or by SM time (throughput).
"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200 import autograd as AG
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

B = int(os.environ.get("B", "16"))
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1, with_gt=True)]


def make():
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().train()


def timed(fn, n=30, warm=6):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n)
    return sorted(ts)[1]


def step_time(label, **eng_attrs):
    net = make()
    eng = net.engine()
    for k, v in eng_attrs.items():
        setattr(eng, k, v)
    step = dasr.TrainStep(net, graph=True)
    ms = timed(lambda: step(*inp))
    print("%-58s %.3f ms/step  (%d launches)" % (label, ms, step.launches_per_step), flush=True)
    return ms


step_time("captured step as shipped")
step_time("whole backward on one stream (wgrad_overlap off)", wgrad_overlap=False)
step_time("only weight gradients on side streams (leaf_overlap off)", leaf_overlap=False)
orig = AG.Tape.wgrad
AG.Tape.wgrad = lambda self, *a, **k: None
step_time("NO weight-gradient launches (diagnostic, wrong numbers)")
step_time("NO weight-gradient launches, one stream", wgrad_overlap=False)
AG.Tape.wgrad = orig
net = make().eval()
with torch.no_grad():
    ms = timed(lambda: net(*inp[:3]))
print("%-58s %.3f ms" % ("inference forward, B=%d" % B, ms))
