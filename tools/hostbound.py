"""Is the B=64 forward host-bound?  Host time to ISSUE a step vs device time per step, kernel by kernel and replayed
from a CUDA graph (Engine.graph_max_pixels raised)."""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = 20
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1)]
eng = net.engine()


def run(tag):
    with torch.no_grad():
        for _ in range(4):
            net(*inp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(K):
            net(*inp)
        e1.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_all = time.perf_counter() - t0
    print("%-28s host issue %.3f ms/step   device (events) %.3f ms/step   wall %.3f ms/step" % (
        tag, t_issue / K * 1e3, e0.elapsed_time(e1) / K, t_all / K * 1e3), flush=True)


run("kernel by kernel")
eng.graph_max_pixels = 10 ** 9
run("CUDA graph replay")
eng.actv_overlap = False
eng._graphs.clear()
run("graph, actv on main stream")
eng.graph_max_pixels = 0
run("eager, actv on main stream")
