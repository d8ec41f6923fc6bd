"""Kernel-by-kernel (no CUDA graph) training steps: per-step host / device time, allocator state and Python GC
activity, to see where an erratic step time comes from."""
import sys, os, time, warnings, gc
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

B = 16
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().train()
step = dasr.TrainStep(net, graph=False)
inp = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1, with_gt=True)]
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
if mode == "nogc":
    gc.disable()
gc.callbacks.append(lambda phase, info: phase == "stop" and print("   gc gen%d collected %d" % (info["generation"], info["collected"])))
for i in range(24):
    t0 = time.time()
    step(*inp)
    t1 = time.time()
    torch.cuda.synchronize()
    t2 = time.time()
    st = torch.cuda.memory_stats()
    print("step %2d issue %.1f ms  total %.1f ms  reserved %.2f GB allocated %.2f GB  cudaMalloc calls %d  retries %d" % (
        i, (t1 - t0) * 1e3, (t2 - t0) * 1e3, st["reserved_bytes.all.current"] / 1e9, st["allocated_bytes.all.current"] / 1e9,
        st["segment.all.allocated"], st["num_alloc_retries"]), flush=True)
