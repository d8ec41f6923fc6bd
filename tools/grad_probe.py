"""Where does a whole-network gradient deviate?  d(loss)/d(block output) of every depth-guided block, CUDA (optionally
in the fp32-split mode, PLANES=3) against the fp64 oracle.  Run under gpurun: CASE=x4_b1_24 PLANES=3 python tools/grad_probe.py"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

from common import case_tensors, load_golden, oracle
import depth_aware_endoscopy_sr_b200 as dasr
import depth_aware_endoscopy_sr_b200.loss as bl
from depth_aware_endoscopy_sr_b200 import _lib as L

L.set_planes(int(os.environ.get("PLANES", "3")))
name = os.environ.get("CASE", "x4_b1_24")
z, meta = load_golden(name)
sd, (lq, depth, masks, gt) = case_tensors(meta)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


# oracle fp64 with retained activation gradients
sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
wd = torch.ones(10, dtype=torch.float64, requires_grad=True)
cap = {}
sr = oracle.depthnet_forward(sdr, lq.double(), depth.double(), masks.double(), scale=meta["scale"], which=meta["which"], cap=cap)
for k, v in cap.items():
    if v.requires_grad:
        v.retain_grad()
total, *_ = oracle.training_loss(sr, gt.double(), masks.double(), wd)
total.backward()

with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"], nb=16, nf=64, depthRangeNum=10)
net.load_state_dict(sd, strict=True)
net = net.cuda().train()
eng = net.engine()
eng.debug = {}
wdc = torch.ones(10, device="cuda", requires_grad=True)
src = net(lq.cuda(), depth.cuda(), masks.cuda())
t, *_ = bl.training_loss(src, gt.cuda(), masks.cuda(), wdc)
t.backward()
torch.cuda.synchronize()
print("loss", t.item(), total.item(), "sr max abs", (src.detach().cpu().double() - sr.detach()).abs().max().item())
dbg = eng.debug


def nchw(x):
    return x.permute(0, 3, 1, 2).cpu()


print("feat_add1 dout rel", rel(nchw(dbg["feat_add1.dout"]), cap["feat_add1"].grad))
for i in sorted(meta["which"], reverse=True):
    p = "depth-residual%d" % (i + 1)
    if p + ".out" not in cap or p + ".norm2.dout" not in dbg:
        continue
    print("%-20s out rel %.3g   d/d(out) rel %.3g   a rel %.3g   d/d(a) rel %.3g" % (
        p, rel(nchw(dbg[p + ".norm2.out"]), cap[p + ".out"]), rel(nchw(dbg[p + ".norm2.dout"]), cap[p + ".out"].grad),
        rel(nchw(dbg[p + ".norm1.out"]), cap[p + ".a"]), rel(nchw(dbg[p + ".norm1.dout"]), cap[p + ".a"].grad)))
worst = []
for k, prm in net.named_parameters():
    if prm.grad is None or sdr[k].grad is None or sdr[k].grad.norm() < 1e-12:
        continue
    worst.append((rel(prm.grad.cpu(), sdr[k].grad), k))
worst.sort(reverse=True)
print("worst params", worst[:6])
import numpy as np
print("median", np.median([w[0] for w in worst]))
