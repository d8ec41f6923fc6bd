"""In-kernel stall accounting of the SEAN convolution in its NETWORK form (fused finalize from the statistics
partials, K-DYN extension, residual), single-CTA kernel vs CTA pairs.  Needs the -DDASR_PROFILE build:
   python -c "import __graft_entry__ as g; print(g.build_profile())"
   DASR_LIB_PATH=depth_aware_endoscopy_sr_b200/libdasr_b200_prof.so [DASR_DBG=1|2|3] python tools/prof_sean.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.load()
NAMES = ["mma:wait acc_empty", "mma:wait a_full", "mma:wait b_full", "mma:issue", "epi:wait acc_full", "epi:work"]
B, H, nf = int(os.environ.get("B", "64")), 64, 64
actv = torch.randn(B, H, H, 2 * nf, device=dev).relu().to(torch.bfloat16)
w = torch.randn(2 * nf, 2 * nf, 3, 3, device=dev) / (2 * nf * 9) ** 0.5
wp = torch.zeros(2 * nf, 9 * 2 * nf, device=dev, dtype=torch.bfloat16)
bp = torch.zeros(2 * nf, device=dev)
L.pack_weights([L.pack_desc(w, wp, bias=torch.zeros(2 * nf, device=dev), dst_bias=bp)], torch.zeros(4096, device=dev))
y = torch.randn(B, H, H, nf, device=dev).to(torch.bfloat16)
resid = torch.randn(B, H, H, nf, device=dev).to(torch.bfloat16)
nslots = L.conv_stats_slots(B, H, H, nf, nf)
stats = torch.zeros(B, nslots, nf, 2, device=dev)
stats[..., 1] = H * H / nslots
lab = torch.randint(0, 10, (B, H, H), device=dev)
mask16 = torch.nn.functional.one_hot(lab, 16).to(torch.bfloat16).contiguous()
wdyn = (torch.randn(B * 2 * nf, 9 * 16, device=dev) * 0.1).to(torch.bfloat16)
out = torch.empty(B, H, H, nf, device=dev, dtype=torch.bfloat16)
dbg = int(os.environ.get("DASR_DBG", "0"))
if hasattr(lib, "dasr_prof_set"):
    lib.dasr_prof_set(dbg)
print("#### DASR_DBG=%d (bit 0: no epilogue global loads, bit 1: no epilogue global stores)" % dbg)
for pair in (0, 1):
    L.check(lib.dasr_set_sean_pair(pair))
    f = lambda: L.conv_fwd(actv, wp, bp, out, Cout=2 * nf, ks=3, epi=L.EPI_SEAN, act=L.ACT_RELU, y=y, stats=stats,
                           dyn_x=mask16, dyn_w=wdyn, resid=resid)
    for _ in range(3):
        f()
    buf = (ctypes.c_ulonglong * 16)()
    if hasattr(lib, "dasr_prof_read"):
        lib.dasr_prof_read(buf, 1)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if hasattr(lib, "dasr_prof_read"):
        lib.dasr_prof_read(buf, 1)
    ms = sorted(ts)[len(ts) // 2]
    fl = 2.0 * B * H * H * (2 * nf) * (2 * nf) * 9
    n = 5 * (74 if pair else 148)
    print("== sean 128->128 B%d@64 %s  %.1f us  %.0f TFLOP/s (main GEMM only)" % (B, "CTA pairs " if pair else "single CTA", ms * 1e3, fl / ms / 1e9))
    print("   per issuing CTA, k cycles: " + "  ".join("%s=%d" % (NAMES[i], buf[i] / n / 1000) for i in range(6)), flush=True)
    print("   whole kernel per CTA: %.1f k cycles in %.1f us -> SM clock during the kernel %.0f MHz" % (
        buf[8] / (5 * 148) / 1e3, buf[9] / (5 * 148) / 1e3, 1e3 * buf[8] / max(buf[9], 1)))
