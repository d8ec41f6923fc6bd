"""Does a 64-frame batch run faster as TWO 32-frame forwards on two streams (CUDA-graph replays side by side)?
Every kernel is persistent with one CTA per SM, so a kernel boundary costs the drain of the last tiles plus the next
kernel's prologue and pipeline ramp; two independent chains can fill each other's boundaries.  (run under gpurun)"""
import copy, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

B = int(os.environ.get("B", "64"))
PARTS = int(os.environ.get("PARTS", "2"))
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    mk = lambda: dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16)
    net = mk().cuda().eval()
    nets = [net]
    for _ in range(PARTS - 1):            # one engine (packed weights, recorded graphs, static buffers) per stream
        n = mk()
        n.load_state_dict(net.state_dict())
        nets.append(n.cuda().eval())
lq, depth, masks = [t.cuda() for t in synthetic_inputs(B, 64, 64, seed=3)]


def timed(fn, n=20):
    for _ in range(5):
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
    return ts


with torch.no_grad():
    ref = net(lq, depth, masks)
    t1 = timed(lambda: net(lq, depth, masks))
    print("one call, B=%d:            %s ms" % (B, " ".join("%.3f" % t for t in t1)), flush=True)
    step = B // PARTS
    chunks = [tuple(t[i * step:(i + 1) * step].contiguous() for t in (lq, depth, masks)) for i in range(PARTS)]
    streams = [torch.cuda.Stream() for _ in range(PARTS)]
    outs = [None] * PARTS

    def split():
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        for i, st in enumerate(streams):
            st.wait_event(ev)
            with torch.cuda.stream(st):
                outs[i] = nets[i](*chunks[i])
        for st in streams:
            main.wait_stream(st)

    t2 = timed(split)
    print("%d x B=%d on %d streams:   %s ms" % (PARTS, step, PARTS, " ".join("%.3f" % t for t in t2)), flush=True)
    torch.cuda.synchronize()
    got = torch.cat(outs, 0)
    print("max |split - one call| = %.3g" % (got - ref).abs().max().item())
    t3 = timed(lambda: [nets[0](*chunks[i]) for i in range(PARTS)])
    print("%d x B=%d back to back:    %s ms" % (PARTS, step, " ".join("%.3f" % t for t in t3)), flush=True)
