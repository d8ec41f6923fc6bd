"""Why does the end-to-end leg of bench.py sometimes come out 30-50 % low?  Repeats the frame-level e2e loop (pinned host
buffers, copy streams, record_stream) and prints, per repetition, frames/s, the slowest host-side step, the number of
cudaMalloc calls of the caching allocator inside the loop and the same loop with the outputs HELD per slot instead of
record_stream (deterministic reuse on the main stream).  (run under gpurun)"""
import os, sys, time, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
B, K = 64, 50
dev = torch.device("cuda:0")
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
host_sets = [tuple(t.pin_memory() for t in synthetic_inputs(B, 64, 64, seed=s)[:2]) for s in (1, 2, 3)]
out_u8 = [torch.empty(B, 512, 512, 3, dtype=torch.uint8).pin_memory() for _ in range(2)]
dev_in = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0]) for _ in range(2)]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
cur = torch.cuda.current_stream()


def loop(hold):
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    held = [None, None]
    worst = [0.0]

    def step(i):
        t0 = time.perf_counter()
        slot = i % 2
        with torch.cuda.stream(s_in):
            s_in.wait_event(ev_done[slot])
            for d, h in zip(dev_in[slot], host_sets[i % 3]):
                d.copy_(h, non_blocking=True)
            ev_in[slot].record(s_in)
        cur.wait_event(ev_in[slot])
        if hold:
            cur.wait_event(ev_out[slot])      # the read-back of the frames this slot held has finished
            held[slot] = None
        with torch.no_grad():
            img = net.infer_frames(dev_in[slot][0], dev_in[slot][1])
        ev_done[slot].record(cur)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_done[slot])
            out_u8[slot].copy_(img, non_blocking=True)
            if hold:
                ev_out[slot].record(s_out)
            else:
                img.record_stream(s_out)
        if hold:
            held[slot] = img
        worst[0] = max(worst[0], time.perf_counter() - t0)

    for i in range(3):
        step(i)
    cur.wait_stream(s_in); cur.wait_stream(s_out); torch.cuda.synchronize()
    worst[0] = 0.0
    n0 = torch.cuda.memory_stats()["num_device_alloc"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        step(i)
    cur.wait_stream(s_in); cur.wait_stream(s_out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return B * K / ms * 1e3, worst[0] * 1e3, torch.cuda.memory_stats()["num_device_alloc"] - n0


for rep in range(6):
    for hold in (False, True):
        fps, worst, mallocs = loop(hold)
        print("rep %d %-14s %7.0f frames/s   slowest host step %.2f ms   cudaMalloc calls in the loop %d" % (
            rep, "hold per slot" if hold else "record_stream", fps, worst, mallocs), flush=True)
