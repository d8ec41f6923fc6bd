"""Conv-kernel-only driver for ncu (--set full): one launch each of the four dominant shapes after a warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from depth_aware_endoscopy_sr_b200 import _lib as L
dev = torch.device("cuda:0")
torch.manual_seed(0)

def pack(w, rows_pad=None, bias=None, shuffle_r=0):
    O, I, ks = w.shape[0], w.shape[1], w.shape[2]
    dst = torch.zeros(rows_pad or O, ks * ks * I, device=dev, dtype=torch.bfloat16)
    db = torch.zeros(rows_pad or O, device=dev)
    L.pack_weights([L.pack_desc(w, dst, bias=bias, dst_bias=db, shuffle_r=shuffle_r)], torch.zeros(4096, device=dev))
    return dst, db

cases = []
def case(B, H, Cin, Cout, ks, epi, **kw):
    x = torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    wp, bp = pack(w, rows_pad=16 if Cout == 3 else None, bias=torch.zeros(Cout, device=dev), shuffle_r=2 if epi == L.EPI_SHUFFLE2 else 0)
    if epi == L.EPI_NCHW_F32:
        out = torch.empty(B, 3, H, H, device=dev)
    elif epi == L.EPI_SHUFFLE2:
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
    cases.append(lambda: L.conv_fwd(x, wp, bp, out, Cout=Cout, ks=ks, epi=epi, **extra, **kw))

case(64, 64, 64, 64, 3, L.EPI_STATS)
case(64, 64, 128, 128, 3, L.EPI_SEAN, inner_relu=1)
case(16, 256, 32, 32, 3, L.EPI_STORE, act=L.ACT_RELU)
case(16, 256, 32, 128, 3, L.EPI_SHUFFLE2, act=L.ACT_LRELU)
case(4, 512, 32, 3, 9, L.EPI_NCHW_F32, clamp01=1)
for _ in range(2):
    for c in cases:
        c()
torch.cuda.synchronize()
evs = []
for c in cases:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c(); e1.record(); evs.append((e0, e1))
torch.cuda.synchronize()
print([round(a.elapsed_time(b), 4) for a, b in evs])
