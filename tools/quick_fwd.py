"""GPU bring-up: forward parity against the goldens / oracle + a first timing (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import CASES, case_tensors, load_golden, oracle, psnr
import depth_aware_endoscopy_sr_b200 as dasr

def build(meta, sd):
    net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"], nb=16, nf=64, depthRangeNum=10)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval()

def nchw(t): return t.float().permute(0, 3, 1, 2).contiguous().cpu()

for name in CASES:
    z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    net = build(meta, sd)
    cap = {}
    ocap = {}
    with torch.no_grad():
        sr = net.engine().infer(lq.cuda(), depth.cuda(), masks.cuda(), cap=cap)
        pre = net.engine().infer(lq.cuda(), depth.cuda(), masks.cuda(), clamp=False)
        ref = oracle.depthnet_forward(sd, lq, depth, masks, scale=meta["scale"], which=meta["which"], cap=ocap)
    torch.cuda.synchronize()
    st = meta["stride"]
    print("== %s" % name)
    print("  depthVec   err %.4g (max %.3g)" % (np.abs(cap["depthVec"].cpu().numpy() - z["depthVec"]).max(), np.abs(z["depthVec"]).max()))
    print("  fea_bef    err %.4g (max %.3g)" % ((nchw(cap["fea_bef"]) - ocap["fea_bef"]).abs().max(), ocap["fea_bef"].abs().max()))
    for i in (1, 2, 7, 13):
        k = "depth-residual%d.out" % i
        print("  block%-2d out err %.4g (max %.3g)" % (i, (nchw(cap["block%d.out" % i]) - ocap[k]).abs().max(), ocap[k].abs().max()))
    print("  feat_up3   err %.4g (max %.3g)" % ((nchw(cap["feat_up3"]) - ocap["feat_up3"]).abs().max(), ocap["feat_up3"].abs().max()))
    print("  pre_clamp  err %.4g   sr err vs golden %.4g  vs oracle %.4g  PSNR delta %.5f" % (
        (pre.cpu() - ocap["pre_clamp"]).abs().max(), np.abs(sr.cpu().numpy()[:, :, ::st, ::st] - z["sr"]).max(),
        (sr.cpu() - ref).abs().max(), abs(psnr(sr.cpu(), gt) - psnr(ref, gt))), flush=True)

# timing, B=64 x8 64x64
z, meta = load_golden("x8_b1_64")
sd, _ = case_tensors(meta)
net = build(meta, sd)
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
for B in (1, 16, 64):
    lq, depth, masks = [t.cuda() for t in synthetic_inputs(B, 64, 64, seed=3)]
    with torch.no_grad():
        for _ in range(3): net(lq, depth, masks)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time(); e0.record()
        n = 10
        for _ in range(n): net(lq, depth, masks)
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("B=%d: %.3f ms/forward (wall %.3f)  -> %.1f frames/s" % (B, ms, (time.time() - t0) / n * 1e3, B / ms * 1e3), flush=True)
