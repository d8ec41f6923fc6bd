"""GPU bring-up checks of the backward kernels against torch autograd (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
torch.backends.cudnn.allow_tf32 = False          # references must be true fp32 (cuDNN defaults to TF32 convs)
torch.backends.cuda.matmul.allow_tf32 = False
import torch.nn.functional as F
from depth_aware_endoscopy_sr_b200 import _lib as L

dev = torch.device("cuda:0")
torch.manual_seed(0)
allok = True
failures = []


def report(name, got, ref, tol):
    global allok
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= tol * max(scale, 1e-6)
    allok &= ok
    if not ok:
        failures.append(name)
    print("%-52s max|err|=%.4g (ref max %.3g) %s" % (name, err, scale, "PASS" if ok else "FAIL"), flush=True)


def bf(x):
    return x.to(torch.bfloat16).float()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def wgrad_case(B, H, W, Cin, Cout, kh=3, kw=3, time=False):
    x = torch.randn(B, Cin, H, W, device=dev)
    dy = torch.randn(B, Cout, H, W, device=dev)
    xa, dya = nhwc(x).to(torch.bfloat16), nhwc(dy).to(torch.bfloat16)
    dw = torch.zeros(Cout, kh * kw * Cin, device=dev)
    L.conv_wgrad(dya, xa, dw, kh, kw)
    # the same with the fused bias gradient (one more MMA per K step against a block of ones)
    dw2 = torch.zeros(Cout, kh * kw * Cin, device=dev)
    fused_bias = kh * kw * min(Cin, 128) // (1 if kw == 1 else kh) + 16 <= 512
    db = torch.zeros(Cout, device=dev)
    if fused_bias:
        L.conv_wgrad(dya, xa, dw2, kh, kw, db=db)
    torch.cuda.synchronize()
    w = torch.zeros(Cout, Cin, kh, kw, device=dev, requires_grad=True)
    F.conv2d(bf(x), w, None, padding=(kh // 2, kw // 2)).backward(bf(dy))
    ref = w.grad.permute(0, 2, 3, 1).reshape(Cout, kh * kw * Cin)      # [o][(t,u),i]
    report("wgrad B%d %dx%d %d->%d k%dx%d" % (B, H, W, Cin, Cout, kh, kw), dw, ref, 2e-3)
    if fused_bias:
        report("   + fused bias: dw", dw2, ref, 2e-3)
        report("   + fused bias: db", db, bf(dy).sum(dim=(0, 2, 3)), 2e-3)
    if time:
        ms = timeit(lambda: L.conv_wgrad(dya, xa, dw, kh, kw))
        print("     %.3f ms  %.0f TFLOP/s" % (ms, 2.0 * B * H * W * Cout * Cin * kh * kw / ms / 1e9), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["wgrad"]
    if "wgrad" in which:
        wgrad_case(2, 16, 16, 64, 64)
        wgrad_case(2, 64, 64, 64, 64)
        wgrad_case(2, 64, 64, 128, 128)
        wgrad_case(1, 32, 48, 32, 32)
        wgrad_case(1, 24, 40, 64, 32)
        wgrad_case(1, 24, 40, 32, 64)
        wgrad_case(1, 31, 31, 256, 256)
        wgrad_case(1, 31, 31, 128, 256)
        wgrad_case(2, 64, 64, 64, 256)
        wgrad_case(1, 135, 240, 64, 64)
        wgrad_case(1, 256, 256, 32, 128)
        wgrad_case(1, 64, 64, 32, 32, kh=9, kw=1)
        wgrad_case(1, 128, 256, 32, 32, kh=9, kw=1)
        wgrad_case(16, 64, 64, 64, 64, time=True)
        wgrad_case(16, 64, 64, 128, 128, time=True)
        wgrad_case(16, 256, 256, 32, 32, time=True)
    print("ALL PASS" if allok else "SOME FAILED")


def sean_bwd_case(B=2, H=16, W=16, nf=64):
    """sean_bwd1 / bwd2 against autograd of IN(IN(y)) * (1 + gamma) + beta -> relu."""
    lib = L.load()
    s = L.stream_ptr()
    HW = H * W
    y = (torch.randn(B, H, W, nf, device=dev) * 3 + 1.5).to(torch.bfloat16)
    gamma = torch.randn(B, H, W, nf, device=dev).to(torch.bfloat16)
    beta = torch.randn(B, H, W, nf, device=dev)
    dout = torch.randn(B, H, W, nf, device=dev).to(torch.bfloat16)
    # forward statistics exactly like the forward kernels (sum / sumsq -> finalize)
    yf = y.float()
    stats = torch.stack([yf.sum(dim=(1, 2)), (yf * yf).sum(dim=(1, 2))], dim=-1).reshape(B, 1, nf, 2).contiguous()
    norm = torch.empty(B, nf, 2, device=dev)
    normk = torch.empty(B, nf, device=dev)
    L.check(lib.dasr_instats_finalize(L.ptr(stats), L.ptr(norm), L.ptr(normk), B, nf, HW, 1, s))
    # torch reference
    yr = yf.clone().requires_grad_(True)
    gr = gamma.float().clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    def inorm(t):
        mu = t.mean(dim=(1, 2), keepdim=True)
        var = t.var(dim=(1, 2), unbiased=False, keepdim=True)
        return (t - mu) / torch.sqrt(var + 1e-5)
    n = inorm(inorm(yr))
    out = torch.relu(n * (1 + gr) + br)
    (out * dout.float()).sum().backward()
    act_out = out.detach().to(torch.bfloat16)
    slots = lib.dasr_sean_bwd_slots(HW)
    dgb = torch.empty(B, H, W, 2 * nf, device=dev, dtype=torch.bfloat16)
    dn = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    dskip = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    part = torch.empty(B, slots, nf, 4, device=dev)
    L.check(lib.dasr_sean_bwd1(L.ptr(dout), L.ptr(act_out), L.ptr(y), L.ptr(norm), L.ptr(gamma), L.ptr(dgb), L.ptr(dn),
                               L.ptr(dskip), L.ptr(part), B, HW, nf, s))
    dy = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    dbias = torch.zeros(2 * nf, device=dev)
    L.check(lib.dasr_sean_bwd2(L.ptr(dn), L.ptr(y), L.ptr(norm), L.ptr(normk), L.ptr(part), L.ptr(dy), L.ptr(dbias), B,
                               HW, nf, s))
    torch.cuda.synchronize()
    report("sean_bwd dbias gamma", dbias[:nf], gr.grad.sum(dim=(0, 1, 2)), 1e-2)
    report("sean_bwd dbias beta", dbias[nf:], br.grad.sum(dim=(0, 1, 2)), 1e-2)
    report("sean_bwd dgamma nf%d %dx%d" % (nf, H, W), dgb[..., :nf], gr.grad, 1e-2)
    report("sean_bwd dbeta", dgb[..., nf:], br.grad, 1e-2)
    report("sean_bwd dskip", dskip, br.grad, 1e-2)
    report("sean_bwd dy", dy, yr.grad, 1.5e-2)


def dyn_bwd_case(B=2, K=10, H=16, W=16, nf2=128, L_=256):
    lib = L.load()
    s = L.stream_ptr()
    lab = torch.randint(0, K, (B, H, W), device=dev)
    masks = F.one_hot(lab, K).permute(0, 3, 1, 2).float().contiguous()
    labels = torch.empty(B, H, W, device=dev, dtype=torch.uint8)
    flag = torch.zeros(1, device=dev, dtype=torch.int32)
    L.check(lib.dasr_mask_labels(L.ptr(masks), L.ptr(labels), L.ptr(flag), B, K, H, W, s))
    dgb = torch.randn(B, H, W, nf2, device=dev).to(torch.bfloat16)
    # reference: gb[b] = conv3x3(mask[b], T[b]) with T [nf2][K][3][3]  -> dT via autograd
    T = torch.zeros(B, nf2, K, 3, 3, device=dev, requires_grad=True)
    outs = torch.cat([F.conv2d(masks[b:b + 1], T[b], None, padding=1) for b in range(B)], 0)   # [B,nf2,H,W]
    (outs * dgb.float().permute(0, 3, 1, 2)).sum().backward()
    ref = T.grad.permute(0, 2, 3, 4, 1).reshape(B * K, 9 * nf2).contiguous()       # [b][k][tap][c]
    for use_labels in (True, False):
        dT = torch.zeros(B * K, 9 * nf2, device=dev)
        L.check(lib.dasr_dynconv_bwd(L.ptr(dgb), L.ptr(labels) if use_labels else None, L.ptr(masks), L.ptr(flag) if use_labels else None, L.ptr(dT), B, K, H, W, nf2, s))
        torch.cuda.synchronize()
        report("dynconv_bwd labels=%d" % use_labels, dT, ref, 1e-4)
    # tensor-core path (one-hot channels of the aux tensor) + the flag protocol between the two kernels
    depth = torch.rand(B, 1, H, W, device=dev) * 9.99 + 0.01
    aux = torch.empty(B, H, W, L.AUX_CH, device=dev, dtype=torch.bfloat16)
    L.check(lib.dasr_build_aux(L.ptr(labels), L.ptr(depth), L.ptr(aux), B, K, H, W, s))
    dT = torch.zeros(B * K, 9 * nf2, device=dev)
    L.check(lib.dasr_dynconv_bwd_tc(L.ptr(dgb), L.ptr(aux), L.ptr(flag), L.ptr(dT), B, K, H, W, nf2, s))
    L.check(lib.dasr_dynconv_bwd(L.ptr(dgb), None, L.ptr(masks), L.ptr(flag), L.ptr(dT), B, K, H, W, nf2, s))
    torch.cuda.synchronize()
    report("dynconv_bwd_tc (flag 0: tc runs, fallback idle)", dT, ref, 1e-4)
    one = torch.ones(1, device=dev, dtype=torch.int32)
    dT = torch.zeros(B * K, 9 * nf2, device=dev)
    L.check(lib.dasr_dynconv_bwd_tc(L.ptr(dgb), L.ptr(aux), L.ptr(one), L.ptr(dT), B, K, H, W, nf2, s))
    torch.cuda.synchronize()
    report("dynconv_bwd_tc (flag 1: idle)", dT, torch.zeros_like(dT), 1e-9)
    L.check(lib.dasr_dynconv_bwd(L.ptr(dgb), None, L.ptr(masks), L.ptr(one), L.ptr(dT), B, K, H, W, nf2, s))
    torch.cuda.synchronize()
    report("dynconv_bwd fallback (flag 1)", dT, ref, 1e-4)
    # mlp_mask backward on the tensor cores against autograd of conv3x3(depth, 1 -> nf2)
    dA = torch.randn(B, H, W, nf2, device=dev).to(torch.bfloat16)
    wm = torch.zeros(nf2, 1, 3, 3, device=dev, requires_grad=True)
    bm = torch.zeros(nf2, device=dev, requires_grad=True)
    (F.conv2d(depth, wm, bm, padding=1) * dA.float().permute(0, 3, 1, 2)).sum().backward()
    scr = torch.zeros(nf2, 9 * L.AUX_CH, device=dev)
    gW = torch.zeros(nf2, 9, device=dev); gb = torch.zeros(nf2, device=dev)
    L.check(lib.dasr_actv_bwd_tc(L.ptr(dA), L.ptr(aux), L.ptr(scr), L.ptr(gW), L.ptr(gb), B, H, W, nf2, s))
    torch.cuda.synchronize()
    report("actv_bwd_tc dW", gW, wm.grad.reshape(nf2, 9), 2e-4)
    report("actv_bwd_tc db", gb, bm.grad, 2e-4)
    # table backward
    stp = torch.randn(B * K, L_, device=dev).to(torch.bfloat16)
    Ws = (torch.randn(9 * nf2, L_, device=dev) / 16).to(torch.bfloat16)
    dWs = torch.empty(9 * nf2, L_, device=dev)
    dstp = torch.empty(B * K, L_, device=dev)
    L.check(lib.dasr_table_bwd(L.ptr(ref), L.ptr(stp), L.ptr(Ws), L.ptr(dWs), L.ptr(dstp), B * K, 9 * nf2, L_, s))
    torch.cuda.synchronize()
    report("table_bwd dWs", dWs, ref.t() @ stp.float(), 1e-4)
    report("table_bwd dstp", dstp, ref @ Ws.float(), 1e-4)
    # style mix backward
    vec = torch.randn(B, K, L_, device=dev)
    A = torch.randn(K, K, device=dev, requires_grad=True)
    a = torch.randn(K, device=dev, requires_grad=True)
    vr = vec.clone().requires_grad_(True)
    stp_r = torch.einsum("ji,bic->bjc", A, vr) + a[None, :, None]
    dst = torch.randn(B, K, L_, device=dev)
    (stp_r * dst).sum().backward()
    dA = torch.zeros(K, K, device=dev); da = torch.zeros(K, device=dev); dvec = torch.zeros(B, K, L_, device=dev)
    A_d = A.detach().contiguous()
    L.check(lib.dasr_style_mix_bwd(L.ptr(dst), L.ptr(vec), L.ptr(A_d), L.ptr(dA), L.ptr(da), L.ptr(dvec), B, K, L_, s))
    torch.cuda.synchronize()
    report("style_mix_bwd dA", dA, A.grad, 1e-4)
    report("style_mix_bwd da", da, a.grad, 1e-4)
    report("style_mix_bwd dvec", dvec, vr.grad, 1e-4)


def misc_bwd_case():
    lib = L.load()
    s = L.stream_ptr()
    # actv backward
    B, H, W, C = 2, 16, 24, 128
    depth = torch.rand(B, 1, H, W, device=dev)
    dA = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    w = torch.zeros(C, 1, 3, 3, device=dev, requires_grad=True)
    b = torch.zeros(C, device=dev, requires_grad=True)
    (F.conv2d(depth, w, b, padding=1) * dA.float().permute(0, 3, 1, 2)).sum().backward()
    dW = torch.zeros(C, 9, device=dev); db = torch.zeros(C, device=dev)
    L.check(lib.dasr_actv_bwd(L.ptr(dA), L.ptr(depth), L.ptr(dW), L.ptr(db), B, H, W, C, s))
    torch.cuda.synchronize()
    report("actv_bwd dW", dW, w.grad.reshape(C, 9), 1e-4)
    report("actv_bwd db", db, b.grad, 1e-4)
    # colsum
    x = torch.randn(3, 20, 20, 64, device=dev).to(torch.bfloat16)
    out = torch.zeros(64, device=dev)
    L.check(lib.dasr_colsum(L.ptr(x), L.ptr(out), 3 * 400, 64, s))
    torch.cuda.synchronize()
    report("colsum", out, x.float().sum(dim=(0, 1, 2)), 1e-4)
    # unshuffle + lrelu grad
    B, H, W, Cq = 2, 8, 12, 32
    conv = torch.randn(B, 4 * Cq, H, W, device=dev, requires_grad=True)
    ps = F.leaky_relu(F.pixel_shuffle(conv, 2), 0.2)
    dps = torch.randn_like(ps)
    ps.backward(dps)
    dconv = torch.empty(B, H, W, 4 * Cq, device=dev, dtype=torch.bfloat16)
    dps_a, ps_a = nhwc(dps).to(torch.bfloat16), nhwc(ps.detach()).to(torch.bfloat16)   # keep the operands alive
    L.check(lib.dasr_unshuffle_actgrad(L.ptr(dps_a), L.ptr(ps_a), L.ptr(dconv), B, H, W, Cq, 0.2, 2, s))
    torch.cuda.synchronize()
    # our channel order is s*Cq + c (packed / permuted), torch's is c*4 + s
    ref = conv.grad.reshape(B, Cq, 4, H, W).permute(0, 3, 4, 2, 1).reshape(B, H, W, 4 * Cq)
    report("unshuffle_actgrad", dconv, ref, 1e-2)
    # out9 prep
    B, H, W = 2, 20, 28
    dout = torch.randn(B, 3, H, W, device=dev)
    sr = torch.rand(B, 3, H, W, device=dev); sr[sr < 0.2] = 0; sr[sr > 0.8] = 1
    ap = torch.empty(B, H, W, 32, device=dev, dtype=torch.bfloat16)
    dbias = torch.zeros(3, device=dev)
    L.check(lib.dasr_out9_bwd_prep(L.ptr(dout), L.ptr(sr), L.ptr(ap), L.ptr(dbias), B, H, W, s))
    torch.cuda.synchronize()
    g = dout * ((sr > 0) & (sr < 1)).float()
    ref = torch.zeros(B, H, W, 32, device=dev)
    for u in range(9):
        sh = u - 4      # A'[.., w, u*3+co] = g[.., w - sh]
        src = torch.zeros_like(g)
        if sh >= 0:
            src[..., sh:] = g[..., :W - sh] if sh > 0 else g
        else:
            src[..., :W + sh] = g[..., -sh:]
        ref[..., u * 3:u * 3 + 3] = src.permute(0, 2, 3, 1)
    report("out9_bwd_prep A'", ap, ref, 1e-2)
    report("out9_bwd_prep dbias", dbias, g.sum(dim=(0, 2, 3)), 1e-4)


if __name__ == "__main__" and "elem" in sys.argv[1:]:
    allok = True
    sean_bwd_case()
    sean_bwd_case(1, 24, 40, 32)
    dyn_bwd_case()
    misc_bwd_case()
    print("ALL PASS" if allok else "SOME FAILED")
