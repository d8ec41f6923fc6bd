"""Epilogue ablation of the trunk convolution on CTA pairs (stats_pair_kernel, B = 64, 64x64): in-kernel cycle accounting
with parts of the statistics epilogue removed.  Needs the -DDASR_PROFILE build:
   DASR_LIB_PATH=depth_aware_endoscopy_sr_b200/libdasr_b200_prof.so python tools/prof_stats.py
knob bits: 2 no global stores, 4 no column-sum butterflies, 8 no per-tile slot combine, 16 no block barriers"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.load()
B, H, C = 64, 64, 64
x = torch.randn(B, H, H, C, device=dev).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device=dev) / (C * 9) ** 0.5
wp = torch.zeros(C, 9 * C, device=dev, dtype=torch.bfloat16)
bp = torch.zeros(C, device=dev)
L.pack_weights([L.pack_desc(w, wp, bias=torch.zeros(C, device=dev), dst_bias=bp)], torch.zeros(4096, device=dev))
out = torch.empty(B, H, H, C, device=dev, dtype=torch.bfloat16)
stats = torch.zeros(B, L.conv_stats_slots(B, H, H, C, C), C, 2, device=dev)
f = lambda: L.conv_fwd(x, wp, bp, out, Cout=C, ks=3, epi=L.EPI_STATS, stats=stats)
NAMES = ["mma:wait acc_empty", "mma:wait a_full", "mma:wait b_full", "mma:issue", "epi:wait acc_full", "epi:work"]
for pair in (0, 7):
    L.check(lib.dasr_set_sean_pair(pair))
    for dbg in ((0, 2) if pair == 0 else (0, 2, 4, 6, 8, 24, 30)):
        lib.dasr_prof_set(dbg)
        for _ in range(3):
            f()
        buf = (ctypes.c_ulonglong * 16)()
        lib.dasr_prof_read(buf, 1)
        n = 5
        for _ in range(n):
            f()
        lib.dasr_prof_read(buf, 1)
        div = n * (74 if pair else 148)
        print("%s knob %2d: kernel %.1f k cycles / %.1f us per CTA;  %s" % (
            "pairs " if pair else "single", dbg, buf[8] / (n * 148) / 1e3, buf[9] / (n * 148) / 1e3,
            "  ".join("%s=%d" % (NAMES[i], buf[i] / div / 1000) for i in range(6))), flush=True)
