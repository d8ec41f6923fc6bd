"""In-kernel stall accounting of conv_halo_kernel (needs the -DDASR_PROFILE build):
   DASR_LIB_PATH=depth_aware_endoscopy_sr_b200/libdasr_b200_prof.so python tools/prof_stalls.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.load()
NAMES = ["mma:wait acc_empty", "mma:wait a_full", "mma:wait b_full", "mma:issue", "epi:wait acc_full", "epi:work",
         "aprod:wait a_empty", "bprod:wait b_empty"]

def pack(w, bias=None, shuffle_r=0):
    O, I, ks = w.shape[0], w.shape[1], w.shape[2]
    dst = torch.zeros(O, ks * ks * I, device=dev, dtype=torch.bfloat16)
    db = torch.zeros(O, device=dev)
    L.pack_weights([L.pack_desc(w, dst, bias=bias, dst_bias=db, shuffle_r=shuffle_r)], torch.zeros(4096, device=dev))
    return dst, db

def run(name, B, H, Cin, Cout, ks, epi, **kw):
    x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    wp, bp = pack(w, bias=torch.zeros(Cout, device=dev), shuffle_r=2 if epi == L.EPI_SHUFFLE2 else 0)
    if epi == L.EPI_SHUFFLE2:
        out = torch.empty(B, 2 * H, 2 * H, Cout // 4, device=dev, dtype=torch.bfloat16)
    elif epi == L.EPI_SEAN:
        out = torch.empty(B, H, H, Cout // 2, device=dev, dtype=torch.bfloat16)
    else:
        out = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
    extra = {}
    if epi == L.EPI_STATS:
        extra["stats"] = torch.zeros(B, L.conv_stats_slots(B, H, H, Cin, Cout), Cout, 2, device=dev)
    if epi == L.EPI_SEAN:
        extra["y"] = torch.randn(B, H, H, Cout // 2, device=dev).to(torch.bfloat16)
        extra["norm"] = torch.rand(B, Cout // 2, 2, device=dev)
        extra["gb_s"] = torch.randn(B, H, H, Cout, device=dev).to(torch.bfloat16)
        extra["resid"] = torch.randn(B, H, H, Cout // 2, device=dev).to(torch.bfloat16)
    f = lambda: L.conv_fwd(x, wp, bp, out, Cout=Cout, ks=ks, epi=epi, **extra, **kw)
    for _ in range(3):
        f()
    buf = (ctypes.c_ulonglong * 16)()
    lib.dasr_prof_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); f(); e1.record()
    torch.cuda.synchronize()
    lib.dasr_prof_read(buf, 1)
    ms = e0.elapsed_time(e1)
    fl = 2.0 * B * H * H * Cout * Cin * ks * ks
    print("== %-34s %.3f ms  %.0f TFLOP/s   per-CTA cycles (avg over 148):" % (name, ms, fl / ms / 1e9))
    print("   " + "  ".join("%s=%dk" % (NAMES[i], buf[i] / 148 / 1000) for i in range(8)), flush=True)
    print("   whole kernel per CTA: %.1f k cycles in %.1f us -> SM clock during the kernel %.0f MHz" % (
        buf[8] / 148 / 1e3, buf[9] / 148 / 1e3, 1e3 * buf[8] / max(buf[9], 1)), flush=True)

DBG = int(os.environ.get("DASR_DBG", "0"))
lib.dasr_prof_set(DBG)
print("#### ablation knob DASR_DBG=%d (bit 0: no epilogue global loads, bit 1: no epilogue global stores)" % DBG)
run("trunk 64->64 stats B64@64", 64, 64, 64, 64, 3, L.EPI_STATS)
run("sean 128->128 B64@64", 64, 64, 128, 128, 3, L.EPI_SEAN, act=L.ACT_RELU)
run("classic 32->32 B64@256", 64, 256, 32, 32, 3, L.EPI_STORE, act=L.ACT_RELU)
run("up 32->128 shuffle B64@256", 64, 256, 32, 128, 3, L.EPI_SHUFFLE2, act=L.ACT_LRELU)
run("up1 64->256 shuffle B64@64", 64, 64, 64, 256, 3, L.EPI_SHUFFLE2, act=L.ACT_LRELU)
run("head 64->32 B64@128", 64, 128, 64, 32, 3, L.EPI_STORE, act=L.ACT_LRELU)
