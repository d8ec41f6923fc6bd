"""Kernel-level bring-up of the fp32-split (dasr_set_planes(3)) mode against fp64 torch (run under gpurun):
plain / strided / shuffle convolutions, the weight gradient, conv_out9, then a whole forward + backward."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from depth_aware_endoscopy_sr_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
L.set_planes(int(os.environ.get("PLANES", "3")))
print("planes", L.planes())


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel(got, ref):
    return ((got.double() - ref.double()).norm() / ref.double().norm()).item()


def pack(w, bias=None, mode=L.PACK_CONV, shuffle_r=0):
    O, I, ks = w.shape[0], w.shape[1], w.shape[2]
    rows = O
    kdim = ks * ks * I
    if mode == L.PACK_ROWTAPS:
        dst = L.act_zeros(9 * 32, 32, device=dev)
    elif mode == L.PACK_DGRAD:
        dst = L.act_zeros(I, ks * ks * O, device=dev)
    else:
        dst = L.act_zeros(rows, kdim, device=dev)
    dbias = torch.zeros(max(rows, 32), device=dev, dtype=torch.float32)
    scratch = torch.zeros(4096, device=dev, dtype=torch.float32)
    d = L.pack_desc(w, dst, mode=mode, shuffle_r=shuffle_r, bias=bias, dst_bias=dbias if bias is not None else None)
    L.pack_weights([d], scratch)
    return dst, dbias


def conv_case(B, H, W, Cin, Cout, ks=3, act=L.ACT_NONE, subsample=1, resid=False):
    x = torch.randn(B, Cin, H, W, device=dev)
    w = torch.randn(Cout, Cin, ks, ks, device=dev) / (Cin * ks * ks) ** 0.5
    b = torch.randn(Cout, device=dev)
    wp, bp = pack(w, bias=b)
    xa = L.act_from(nhwc(x))
    Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if subsample == 2 else (H, W)
    out = L.act_empty(B, Ho, Wo, Cout, device=dev)
    r = torch.randn(B, Cout, Ho, Wo, device=dev) if resid else None
    ra = L.act_from(nhwc(r)) if resid else None
    L.conv_fwd(xa, wp, bp, out, Cout=Cout, ks=ks, act=act, subsample=subsample, resid=ra)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=ks // 2, stride=subsample)
    if resid:
        ref = ref + r.double()
    if act == L.ACT_RELU:
        ref = F.relu(ref)
    print("conv B%d %dx%d %d->%d k%d sub%d resid%d: rel err %.3g" % (B, H, W, Cin, Cout, ks, subsample, resid,
                                                                  rel(nchw(L.act_value(out)), ref)), flush=True)


def wgrad_case(B, H, W, Cin, Cout, kh=3, kw=3):
    x = torch.randn(B, Cin, H, W, device=dev)
    dy = torch.randn(B, Cout, H, W, device=dev)
    dw = torch.zeros(Cout, kh * kw * Cin, device=dev)
    db = torch.zeros(Cout, device=dev)
    L.conv_wgrad(L.act_from(nhwc(dy)), L.act_from(nhwc(x)), dw, kh, kw, db=db)
    torch.cuda.synchronize()
    xd = x.double().requires_grad_(False)
    wref = torch.zeros(Cout, Cin, kh, kw, device=dev, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xd, wref, padding=(kh // 2, kw // 2))
    (y * dy.double()).sum().backward()
    ref = wref.grad.permute(0, 2, 3, 1).reshape(Cout, -1)
    print("wgrad B%d %dx%d %d->%d %dx%d: dw rel err %.3g  db rel err %.3g" % (
        B, H, W, Cin, Cout, kh, kw, rel(dw, ref), rel(db, dy.double().sum((0, 2, 3)))), flush=True)


def out9_case(B, H, W):
    x = torch.randn(B, 32, H, W, device=dev)
    w = torch.randn(3, 32, 9, 9, device=dev) / 50.0
    b = torch.randn(3, device=dev)
    wq, bq = pack(w, bias=b, mode=L.PACK_ROWTAPS)
    out = torch.empty(B, 3, H, W, device=dev)
    L.check(L.load().dasr_conv_out9(L.ptr(L.act_from(nhwc(x))), L.ptr(wq), L.ptr(bq), L.ptr(out), B, H, W, 3, 0,
                                    L.stream_ptr()))
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=4)
    print("out9 B%d %dx%d: rel err %.3g" % (B, H, W, rel(out, ref)), flush=True)


conv_case(2, 16, 16, 64, 64)
conv_case(1, 24, 40, 32, 32, act=L.ACT_RELU, resid=True)
conv_case(1, 64, 64, 128, 128)
conv_case(1, 31, 31, 256, 256, subsample=2)
conv_case(1, 64, 64, 32, 64, subsample=2)
conv_case(1, 128, 128, 32, 32)
wgrad_case(2, 16, 16, 64, 64)
wgrad_case(1, 24, 40, 32, 32)
wgrad_case(1, 64, 64, 128, 128)
wgrad_case(1, 64, 64, 32, 32, 9, 1)
out9_case(1, 64, 64)

# ---- whole network
from common import case_tensors, load_golden, oracle  # noqa: E402
import numpy as np  # noqa: E402
import warnings  # noqa: E402
import depth_aware_endoscopy_sr_b200 as dasr  # noqa: E402

name = os.environ.get("CASE", "x8_b2_16")
z, meta = load_golden(name)
sd, (lq, depth, masks, gt) = case_tensors(meta)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"],
                        nb=16, nf=64, depthRangeNum=10)
net.load_state_dict(sd, strict=True)
net = net.cuda().eval()
cap = {}
with torch.no_grad():
    pre = net.engine().infer(lq.cuda(), depth.cuda(), masks.cuda(), clamp=False, cap=cap)
    ocap = {}
    ref = oracle.depthnet_forward(sd, lq, depth, masks, scale=meta["scale"], which=meta["which"], cap=ocap)
torch.cuda.synchronize()
st = meta["stride"]
print("depthVec rel", rel(cap["depthVec"].cpu(), ocap["depthVec"]))
for k in ("fea_bef", "block1.out", "block13.out"):
    ok = {"fea_bef": "fea_bef", "block1.out": "depth-residual1.out", "block13.out": "depth-residual13.out"}[k]
    if k in cap and ok in ocap:
        print(k, "rel", rel(nchw(L.act_value(cap[k])).cpu(), ocap[ok]))
print("pre_clamp max abs err", np.abs(pre.cpu().numpy()[:, :, ::st, ::st] - z["pre_clamp"]).max(), "range",
      np.abs(z["pre_clamp"]).max())
