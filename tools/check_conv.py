"""Ad-hoc GPU bring-up check of the implicit-GEMM conv kernel against torch (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False          # references must be true fp32 (cuDNN defaults to TF32 convs)
torch.backends.cuda.matmul.allow_tf32 = False
import torch.nn.functional as F
from depth_aware_endoscopy_sr_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
L.check(L.load().dasr_check_device())


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def pack(w, rows_pad=None, mode=L.PACK_CONV, g=None, shuffle_r=0, bias=None):
    O = w.shape[0] if mode == L.PACK_CONV else w.shape[1]
    I = w.shape[1] if mode == L.PACK_CONV else w.shape[0]
    ks = w.shape[2]
    rows = rows_pad or O
    dst = torch.zeros(rows, ks * ks * I, device=dev, dtype=torch.bfloat16)
    dbias = torch.zeros(rows, device=dev, dtype=torch.float32)
    scratch = torch.zeros(4096, device=dev, dtype=torch.float32)
    d = L.pack_desc(w, dst, g=g, mode=mode, shuffle_r=shuffle_r, bias=bias, dst_bias=dbias if bias is not None else None)
    L.pack_weights([d], scratch)
    return dst, dbias


def report(name, got, ref, tol):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(scale, 1.0)
    print("%-44s max|err|=%.4g (ref max %.3g) %s" % (name, err, scale, "PASS" if ok else "FAIL"), flush=True)
    return ok


def bf(x):
    return x.to(torch.bfloat16).float()


allok = True


def run_plain(B, H, W, Cin, Cout, ks=3, act=L.ACT_NONE, name=""):
    global allok
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    b = torch.randn(Cout, device=dev)
    wp, bp = pack(w, bias=b)
    xa = nhwc(x).to(torch.bfloat16)
    out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.bfloat16)
    L.conv_fwd(xa, wp, bp, out, Cout=Cout, ks=ks, act=act)
    torch.cuda.synchronize()
    ref = F.conv2d(bf(x), bf(w), b, padding=ks // 2)
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    if act == L.ACT_LRELU:
        ref = F.leaky_relu(ref, 0.2)
    allok &= report("plain %s B%d %dx%d %d->%d k%d" % (name, B, H, W, Cin, Cout, ks), nchw(out), ref, 1.5e-2)


run_plain(2, 64, 64, 64, 64)
run_plain(1, 64, 64, 128, 128, act=L.ACT_RELU)
run_plain(2, 128, 128, 32, 32, act=L.ACT_LRELU)
run_plain(1, 135, 240, 64, 64, name="odd")
run_plain(1, 16, 16, 64, 128, name="small")
run_plain(1, 31, 31, 256, 256, name="enc5-like")
run_plain(1, 1, 640, 256, 1152, ks=1, name="table-gemm")
run_plain(3, 64, 64, 64, 256, name="up1")

# --- stats epilogue
B, H, W, Cn = 2, 64, 64, 64
x = torch.randn(B, Cn, H, W, device=dev)
w = torch.randn(Cn, Cn, 3, 3, device=dev) / 24
b = torch.randn(Cn, device=dev)
wp, bp = pack(w, bias=b)
y = torch.empty(B, H, W, Cn, device=dev, dtype=torch.bfloat16)
nslots = L.conv_stats_slots(B, H, W, Cn, Cn)
stats = torch.full((B, nslots, Cn, 2), float("nan"), device=dev)
L.conv_fwd(nhwc(x).to(torch.bfloat16), wp, bp, y, Cout=Cn, ks=3, epi=L.EPI_STATS, stats=stats)
torch.cuda.synchronize()
ref = F.conv2d(bf(x), bf(w), b, padding=1)
allok &= report("stats: y", nchw(y), ref, 1.5e-2)
allok &= report("stats: sum", stats[..., 0].sum(1), ref.sum(dim=(2, 3)), 2e-3)
allok &= report("stats: sumsq", stats[..., 1].sum(1), (ref * ref).sum(dim=(2, 3)), 2e-3)
norm = torch.empty(B, Cn, 2, device=dev)
L.check(L.load().dasr_instats_finalize(stats.data_ptr(), norm.data_ptr(), None, B, Cn, H * W, nslots, L.stream_ptr()))
mu = ref.mean(dim=(2, 3))
var = ref.var(dim=(2, 3), unbiased=False)
sc = (var + 1e-5).rsqrt() * (var / (var + 1e-5) + 1e-5).rsqrt()
allok &= report("instats: mean", norm[..., 0], mu, 2e-3)
allok &= report("instats: scale", norm[..., 1], sc, 2e-3)

# --- SEAN epilogue: acc=[gamma_o|beta_o], out = relu(x + n*(1+g)+b)
actv = torch.randn(B, 128, H, W, device=dev)
wgb = torch.randn(128, 128, 3, 3, device=dev) / 34
bgb = torch.randn(128, device=dev) * 0.1
wp2, bp2 = pack(wgb, bias=bgb)
gbs = (torch.randn(B, H, W, 128, device=dev) * 0.3).to(torch.bfloat16)
resid = torch.randn(B, H, W, Cn, device=dev).to(torch.bfloat16)
resid32 = torch.randn(B, H, W, Cn, device=dev)
for inner, use_res in ((1, 0), (0, 1), (0, 2)):
    out = torch.empty(B, H, W, Cn, device=dev, dtype=torch.bfloat16)
    out32 = torch.empty(B, H, W, Cn, device=dev) if use_res == 2 else None
    L.conv_fwd(nhwc(actv).to(torch.bfloat16), wp2, bp2, out, Cout=128, ks=3, epi=L.EPI_SEAN,
               act=L.ACT_RELU if use_res else L.ACT_NONE, inner_relu=inner, y=y, norm=norm, gb_s=gbs,
               resid=resid if use_res == 1 else None, resid_f32=resid32 if use_res == 2 else None, out_aux_f32=out32)
    torch.cuda.synchronize()
    gb = F.conv2d(bf(actv), bf(wgb), bgb, padding=1) + nchw(gbs).float()
    n = (nchw(y).float() - norm[..., 0][:, :, None, None]) * norm[..., 1][:, :, None, None]
    r = n * (1 + gb[:, :64]) + gb[:, 64:]
    if inner:
        r = F.relu(r)
    if use_res == 1:
        r = F.relu(r + nchw(resid).float())
    if use_res == 2:
        r = F.relu(r + nchw(resid32))
        allok &= report("sean epilogue fp32 aux out", nchw(out32), r, 2e-3)
    allok &= report("sean epilogue inner=%d resid=%d" % (inner, use_res), nchw(out), r, 2e-2)

# --- pixel shuffle epilogue
x = torch.randn(2, 32, 64, 64, device=dev)
w = torch.randn(128, 32, 3, 3, device=dev) / 17
b = torch.randn(128, device=dev)
wp, bp = pack(w, bias=b, shuffle_r=2)
out = torch.empty(2, 128, 128, 32, device=dev, dtype=torch.bfloat16)
L.conv_fwd(nhwc(x).to(torch.bfloat16), wp, bp, out, Cout=128, ks=3, epi=L.EPI_SHUFFLE2, act=L.ACT_LRELU)
torch.cuda.synchronize()
ref = F.leaky_relu(F.pixel_shuffle(F.conv2d(bf(x), bf(w), b, padding=1), 2), 0.2)
allok &= report("shuffle2 32->128", nchw(out), ref, 1.5e-2)
x = torch.randn(1, 64, 64, 64, device=dev)
w = torch.randn(256, 64, 3, 3, device=dev) / 24
b = torch.randn(256, device=dev)
wp, bp = pack(w, bias=b, shuffle_r=2)
out = torch.empty(1, 128, 128, 64, device=dev, dtype=torch.bfloat16)
L.conv_fwd(nhwc(x).to(torch.bfloat16), wp, bp, out, Cout=256, ks=3, epi=L.EPI_SHUFFLE2, act=L.ACT_LRELU)
torch.cuda.synchronize()
ref = F.leaky_relu(F.pixel_shuffle(F.conv2d(bf(x), bf(w), b, padding=1), 2), 0.2)
allok &= report("shuffle2 64->256", nchw(out), ref, 1.5e-2)

# --- 9x9 output conv, NCHW fp32 + clamp (K-OUT9)
def pack9(w, b):
    wq = torch.zeros(9 * 32, 32, device=dev, dtype=torch.bfloat16)
    bq = torch.zeros(3, device=dev)
    L.pack_weights([L.pack_desc(w, wq, bias=b, dst_bias=bq, mode=L.PACK_ROWTAPS)], torch.zeros(64, device=dev))
    return wq, bq

for (bb, hh, ww, cl) in ((1, 96, 96, 1), (2, 64, 200, 0), (1, 37, 61, 1), (3, 128, 128, 0)):
    x = torch.rand(bb, 32, hh, ww, device=dev)
    w = torch.randn(3, 32, 9, 9, device=dev) / 51
    b = torch.randn(3, device=dev) * 0.1 + 0.3
    wq, bq = pack9(w, b)
    out = torch.full((bb, 3, hh, ww), float("nan"), device=dev)
    L.check(L.load().dasr_conv_out9(nhwc(x).to(torch.bfloat16).data_ptr(), wq.data_ptr(), bq.data_ptr(), out.data_ptr(),
                                    bb, hh, ww, 3, cl, L.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv2d(bf(x), bf(w), b, padding=4)
    if cl:
        ref = ref.clamp(0, 1)
    allok &= report("conv_out9 B%d %dx%d clamp=%d" % (bb, hh, ww, cl), out, ref, 2e-3)

# --- stride 2 via subsample
for hh in (64, 31):
    x = torch.randn(2, 64, hh, hh, device=dev)
    w = torch.randn(128, 64, 3, 3, device=dev) / 24
    b = torch.randn(128, device=dev)
    wp, bp = pack(w, bias=b)
    ho = (hh + 1) // 2
    out = torch.empty(2, ho, ho, 128, device=dev, dtype=torch.bfloat16)
    L.conv_fwd(nhwc(x).to(torch.bfloat16), wp, bp, out, Cout=128, ks=3, subsample=2, act=L.ACT_LRELU)
    torch.cuda.synchronize()
    ref = F.leaky_relu(F.conv2d(bf(x), bf(w), b, stride=2, padding=1), 0.2)
    allok &= report("stride2 %d" % hh, nchw(out), ref, 1.5e-2)

# --- transposed conv via zero insertion + CONVT pack + weight norm
x = torch.randn(2, 128, 16, 16, device=dev)
v = torch.randn(128, 256, 3, 3, device=dev) / 34
g = torch.rand(128, 1, 1, 1, device=dev) + 0.5
b = torch.randn(256, device=dev)
wn = v * (g / v.reshape(128, -1).norm(dim=1).reshape(128, 1, 1, 1))
wp, bp = pack(v, mode=L.PACK_CONVT, g=g, bias=b)
xz = torch.empty(2, 31, 31, 128, device=dev, dtype=torch.bfloat16)
L.check(L.load().dasr_zero_insert2(nhwc(x).to(torch.bfloat16).data_ptr(), xz.data_ptr(), 2, 16, 16, 128, L.stream_ptr()))
out = torch.empty(2, 31, 31, 256, device=dev, dtype=torch.bfloat16)
L.conv_fwd(xz, wp, bp, out, Cout=256, ks=3)
torch.cuda.synchronize()
ref = F.conv_transpose2d(bf(x), bf(wn), b, stride=2, padding=1)
allok &= report("convT 128->256 (wn)", nchw(out), ref, 2e-2)

# --- timing of the two dominant shapes at B=64
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for (Cin, Cout, hh, name) in ((64, 64, 64, "trunk 64->64"), (128, 128, 64, "gamma/beta 128->128"), (32, 32, 256, "classic16 32->32@256"),
                              (32, 128, 256, "up3 32->128@256")):
    Bt = 64
    x = torch.randn(Bt, hh, hh, Cin, device=dev).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 3, 3, device=dev) / 30
    b = torch.randn(Cout, device=dev)
    wp, bp = pack(w, bias=b)
    out = torch.empty(Bt, hh, hh, Cout, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: L.conv_fwd(x, wp, bp, out, Cout=Cout, ks=3))
    fl = 2.0 * Bt * hh * hh * Cout * Cin * 9
    print("time %-26s %.3f ms  %.1f TFLOP/s" % (name, ms, fl / ms / 1e9), flush=True)
    xt = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    wt = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    ms2 = timeit(lambda: F.conv2d(xt, wt, None, padding=1))
    print("     cudnn bf16 channels_last            %.3f ms  %.1f TFLOP/s" % (ms2, fl / ms2 / 1e9), flush=True)

x = torch.rand(64, 512, 512, 32, device=dev).to(torch.bfloat16)
w = torch.randn(3, 32, 9, 9, device=dev) / 51
b = torch.zeros(3, device=dev)
wq, bq = pack9(w, b)
out = torch.empty(64, 3, 512, 512, device=dev)
ms = timeit(lambda: L.check(L.load().dasr_conv_out9(x.data_ptr(), wq.data_ptr(), bq.data_ptr(), out.data_ptr(), 64, 512, 512, 3, 1, L.stream_ptr())), n=5)
print("time conv_out9 B64@512: %.3f ms (%.1f real TFLOP/s, %.0f GB/s algorithmic)" % (ms, 2.0 * 64 * 512 * 512 * 3 * 32 * 81 / ms / 1e9, (x.numel() * 2 + out.numel() * 4) / ms / 1e6))
xt = x.permute(0, 3, 1, 2)          # NHWC storage viewed NCHW = channels_last
wt = w.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
torch.backends.cudnn.benchmark = True
ms2 = timeit(lambda: F.conv2d(xt, wt, None, padding=4), n=5)
print("     cudnn bf16 channels_last 9x9 32->3   %.3f ms  %.1f TFLOP/s" % (ms2, 2.0 * 64 * 512 * 512 * 3 * 32 * 81 / ms2 / 1e9), flush=True)

print("ALL PASS" if allok else "SOME FAILED")
