import os, sys, time, warnings
sys.path.insert(0, "/root/repo")
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
B=64
host = [t.pin_memory() for t in synthetic_inputs(B, 64, 64, scale=8, seed=1)]
dev_in = [tuple(torch.empty_like(t, device="cuda") for t in host) for _ in range(2)]
out_host = [torch.empty(B, 3, 512, 512).pin_memory() for _ in range(2)]
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
cur = torch.cuda.current_stream()
ev_in = [torch.cuda.Event() for _ in range(2)]; ev_done = [torch.cuda.Event() for _ in range(2)]
def step(i):
    slot = i % 2
    with torch.cuda.stream(s_in):
        s_in.wait_event(ev_done[slot])
        for d, h in zip(dev_in[slot], host): d.copy_(h, non_blocking=True)
        ev_in[slot].record(s_in)
    cur.wait_event(ev_in[slot])
    sr = net(*dev_in[slot])
    ev_done[slot].record(cur)
    with torch.cuda.stream(s_out):
        s_out.wait_event(ev_done[slot])
        out_host[slot].copy_(sr, non_blocking=True)
        sr.record_stream(s_out)
with torch.no_grad():
    for mode in ("graph", "eager"):
        net.engine().use_graphs = (mode == "graph")
        for i in range(5): step(i)
        cur.wait_stream(s_in); cur.wait_stream(s_out); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0=time.time(); e0.record()
        for i in range(40): step(i)
        cur.wait_stream(s_in); cur.wait_stream(s_out); e1.record(); th=time.time()-t0; torch.cuda.synchronize()
        print(mode, "e2e %.0f frames/s  (%.2f ms/step, host %.2f ms/step)  mem reserved %.1f GB" % (B*40/(e0.elapsed_time(e1)*1e-3), e0.elapsed_time(e1)/40, th/40*1e3, torch.cuda.memory_reserved()/1e9), flush=True)
