"""GPU bring-up: full-network gradients of the CUDA path against the fp32 CPU oracle (run under gpurun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from common import case_tensors, load_golden, oracle
import depth_aware_endoscopy_sr_b200 as dasr

def run(name, verbose=True):
    z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    # oracle (fp32 CPU autograd)
    sdr = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    wdyn = torch.ones(10, requires_grad=True)
    if os.environ.get("BF16REF"):
        with oracle.bf16_operands():
            sr_ref = oracle.depthnet_forward(sdr, lq, depth, masks, scale=meta["scale"], which=meta["which"])
            total_ref, *_ = oracle.training_loss(sr_ref, gt, masks, wdyn)
            total_ref.backward()
    else:
        sr_ref = oracle.depthnet_forward(sdr, lq, depth, masks, scale=meta["scale"], which=meta["which"])
        total_ref, *_ = oracle.training_loss(sr_ref, gt, masks, wdyn)
        total_ref.backward()
    # CUDA path
    net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"], nb=16, nf=64, depthRangeNum=10)
    net.load_state_dict(sd, strict=True)
    net = net.cuda().train()
    wd = torch.ones(10, device="cuda", requires_grad=True)
    sr = net(lq.cuda(), depth.cuda(), masks.cuda())
    total, *_ = oracle.training_loss(sr, gt.cuda(), masks.cuda(), wd)
    total.backward()
    torch.cuda.synchronize()
    print("== %s: loss %.6f (oracle %.6f)  max|sr-ref| %.4g" % (name, total.item(), total_ref.item(), (sr.detach().cpu() - sr_ref.detach()).abs().max()))
    worst = []
    for k, p in net.named_parameters():
        gr = sdr[k].grad
        if gr is None:
            assert p.grad is None, k
            continue
        if p.grad is None:
            print("   MISSING grad for", k); worst.append((9.9, k)); continue
        g = p.grad.detach().cpu()
        rel = ((g - gr).norm() / (gr.norm() + 1e-20)).item()
        if gr.norm() < 1e-9:      # exactly-zero gradients (conv bias in front of IN)
            rel = g.abs().max().item()
        worst.append((rel, k))
    worst.sort(reverse=True)
    if verbose:
        for rel, k in worst[:25]:
            print("   rel-L2 err %.4f  %s  (|ref| %.3g)" % (rel, k, sdr[k].grad.norm() if sdr[k].grad is not None else 0))
    if os.environ.get("ALLP"):
        order = {k: i for i, (k, _) in enumerate(net.named_parameters())}
        for rel, k in sorted(worst, key=lambda t: order[t[1]]):
            if ("residual" not in k) or any(("residual%d." % i) in k for i in (1, 7, 13, 15, 16)):
                print("   %-55s rel %.4f  |ref| %.3g" % (k, rel, sdr[k].grad.norm()))
    rels = np.array([w[0] for w in worst])
    print("   params: %d   median rel err %.4f   90%% %.4f   max %.4f" % (len(rels), np.median(rels), np.quantile(rels, 0.9), rels.max()), flush=True)
    return net, (lq, depth, masks, gt)

if __name__ == "__main__":
    names = sys.argv[1:] or ["x8_b2_32_init", "x8_b2_16"]
    for n in names:
        net, inp = run(n, verbose=not os.environ.get("ALLP"))
    if os.environ.get("ALLP"):
        sys.exit(0)
    # timing of a training step (fwd + bwd), B=16
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(16, 64, 64, seed=3, with_gt=True)]
    wd = torch.ones(10, device="cuda", requires_grad=True)
    def step():
        for p in net.parameters(): p.grad = None
        sr = net(lq, depth, masks)
        total, *_ = oracle.training_loss(sr, gt, masks, wd)
        total.backward()
    for _ in range(2): step()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(5): step()
    torch.cuda.synchronize()
    print("train step B=16 (fwd+loss+bwd): %.1f ms" % ((time.time() - t0) / 5 * 1e3))
