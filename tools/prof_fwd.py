"""One forward of the bench workload (x8, B=64, 64x64 LR) between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum  (launch list of one step)
   ncu --profile-from-start off --set full -k regex:conv_halo ...   (full capture of the dominant kernel)"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1)]
with torch.no_grad():
    for _ in range(3):
        net(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    net(*inp)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
