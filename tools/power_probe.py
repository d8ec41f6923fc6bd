"""Is the B=64 forward step power-bound?  (run under gpurun)

Runs the forward in a loop for a few seconds per setting while `nvidia-smi -lms 50` samples SM clock, board power and
the throttle reasons, and prints ms/step over consecutive 0.5 s windows next to the clock / power of the same window.
Settings: single-CTA kernels everywhere (pair mask 0), the default mask, all pair kernels (mask 7).
"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import depth_aware_endoscopy_sr_b200 as dasr
from depth_aware_endoscopy_sr_b200 import _lib as L
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
from bench import ClockSampler

B = int(os.environ.get("B", "64"))
SECONDS = float(os.environ.get("SECONDS", "4"))
torch.manual_seed(0)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().eval()
sets = [[t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=s)] for s in (1, 2, 3)]
lib = L.load()
N = 40          # steps per window


def window():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        net(*sets[i % 3])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N


with torch.no_grad():
    for mask in [int(m) for m in os.environ.get("MODES", "0,5,7").split(",")]:
        L.check(lib.dasr_set_sean_pair(mask))
        for i in range(4):
            net(*sets[i % 3])
        torch.cuda.synchronize()
        time.sleep(3.0)                 # let the board cool down to its idle state
        cs = ClockSampler(0)
        cs.start()
        time.sleep(0.3)
        t0 = time.time()
        print("== pair mask %d" % mask)
        while time.time() - t0 < SECONDS:
            m0 = cs.mark()
            ms = window()
            rows = cs.rows[m0:]
            clk = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
            pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
            cap = any(r[6].lower().startswith("active") for r in rows)
            print("  t=%4.1f s  %.3f ms/step   SM %s MHz   %s W   sw_power_cap %s" % (
                time.time() - t0, ms, "%.0f" % (sum(clk) / len(clk)) if clk else "?",
                "%.0f" % (sum(pw) / len(pw)) if pw else "?", "active" if cap else "-"), flush=True)
        cs.stop()
