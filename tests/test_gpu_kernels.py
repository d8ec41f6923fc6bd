"""B200 kernel-level tests (through the C ABI) of the fused SEAN convolution: the [gamma_o;beta_o] GEMM with the
K-DYN extension (mask image x per-image dynamic filters), the double-InstanceNorm finalize in its prologue and the
modulate / residual epilogue, against plain fp32 torch of the same op on bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


@pytest.mark.parametrize("nf", [64, 32])      # 32: a depth-guided block behind upscale1 / upscale2 (which_ResBlk_depth 14, 15)
@pytest.mark.parametrize("shape", [(2, 64, 64), (3, 24, 40), (1, 135, 240)])
def test_sean_conv_with_kdyn_extension_and_fused_finalize(shape, nf):
    from depth_aware_endoscopy_sr_b200 import _lib as L
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, H, W = shape
    K = 10
    dev = torch.device("cuda:0")
    lib = L.load()
    s = L.stream_ptr()
    g = torch.Generator(device="cpu").manual_seed(B * 100 + H)
    rnd = lambda *sh: torch.randn(*sh, generator=g).to(dev)

    def pack(w, bias):
        O, I, ks = w.shape[0], w.shape[1], w.shape[2]
        dst = torch.zeros(O, ks * ks * I, device=dev, dtype=torch.bfloat16)
        db = torch.zeros(O, device=dev)
        L.pack_weights([L.pack_desc(w, dst, bias=bias, dst_bias=db)], torch.zeros(4096, device=dev))
        return dst, db

    # producing convolution with statistics partials
    x = rnd(B, nf, H, W)
    w1, b1 = rnd(nf, nf, 3, 3) / 24, rnd(nf)
    wp1, bp1 = pack(w1, b1)
    y = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    nslots = L.conv_stats_slots(B, H, W, nf, nf)
    stats = torch.full((B, nslots, nf, 2), float("nan"), device=dev)
    L.conv_fwd(_nhwc(x).to(torch.bfloat16), wp1, bp1, y, Cout=nf, ks=3, epi=L.EPI_STATS, stats=stats)
    # masks (one-hot with a few pixels in no mask and a few fractional/overlapping ones), table, dynamic weights
    lab = torch.randint(0, K, (B, H, W), generator=g)
    masks = F.one_hot(lab, K).permute(0, 3, 1, 2).float()
    masks[:, :, 0, :3] = 0.0
    masks[:, 2, 1, :5] = 0.5
    masks = masks.to(dev).contiguous()
    table = (rnd(B, K, 9, 2 * nf) * 0.2).to(torch.bfloat16)
    mask16 = torch.empty(B, H, W, 16, device=dev, dtype=torch.bfloat16)
    L.check(lib.dasr_build_mask16(L.ptr(masks), L.ptr(mask16), B, K, H, W, s))
    wdyn = torch.empty(B * 2 * nf, 9 * 16, device=dev, dtype=torch.bfloat16)
    L.check(lib.dasr_table_to_dynweights(L.ptr(table), L.ptr(wdyn), B, K, 2 * nf, 0, s))
    # the SEAN convolution itself
    actv = rnd(B, 2 * nf, H, W)
    wgb, bgb = rnd(2 * nf, 2 * nf, 3, 3) / 34, rnd(2 * nf) * 0.1
    wp2, bp2 = pack(wgb, bgb)
    resid32 = rnd(B, H, W, nf)
    out = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    out32 = torch.empty(B, H, W, nf, device=dev)
    norm = torch.empty(B, nf, 2, device=dev)
    normk = torch.empty(B, nf, device=dev)
    gamma = torch.empty(B, H, W, nf, device=dev, dtype=torch.bfloat16)
    L.conv_fwd(_nhwc(actv).to(torch.bfloat16), wp2, bp2, out, Cout=2 * nf, ks=3, epi=L.EPI_SEAN, act=L.ACT_RELU, y=y,
               stats=stats, norm_out=norm, normk_out=normk, dyn_x=mask16, dyn_w=wdyn, resid_f32=resid32,
               out_aux_f32=out32, gamma_out=gamma)
    torch.cuda.synchronize()
    # reference
    yf = _nchw(y).float()
    mu = yf.mean(dim=(2, 3))
    var = yf.var(dim=(2, 3), unbiased=False)
    sc = (var + 1e-5).rsqrt() * (var / (var + 1e-5) + 1e-5).rsqrt()
    assert (norm[..., 0] - mu).abs().max() <= 2e-3 * mu.abs().max() + 1e-5
    assert (norm[..., 1] - sc).abs().max() <= 2e-3 * sc.abs().max()
    a, r = var + 1e-5, var / (var + 1e-5) + 1e-5
    assert (normk - (1 / a + 1e-5 / (a * a * r))).abs().max() <= 5e-3 * (1 / a).abs().max()
    # dynamic conv: per-image weights Tw[b][c][k][t][u] = table[b][k][t*3+u][c]
    Tw = table.float().view(B, K, 3, 3, 2 * nf).permute(0, 4, 1, 2, 3).contiguous()
    gbs = torch.cat([F.conv2d(_bf(masks[b:b + 1]), Tw[b], None, padding=1) for b in range(B)], 0)
    gb = F.conv2d(_bf(actv), _bf(wgb), bgb, padding=1) + gbs
    n = (yf - norm[..., 0][:, :, None, None]) * norm[..., 1][:, :, None, None]
    ref = F.relu(n * (1 + gb[:, :nf]) + gb[:, nf:] + _nchw(resid32))
    scale = ref.abs().max().item()
    assert (_nchw(out32) - ref).abs().max().item() <= 3e-3 * scale
    assert (_nchw(out).float() - ref).abs().max().item() <= 1.5e-2 * scale
    assert (_nchw(gamma).float() - gb[:, :nf]).abs().max().item() <= 1.5e-2 * gb.abs().max().item()


@pytest.mark.parametrize("B", [10, 11])
def test_sean_conv_on_cta_pairs_equals_single_cta_kernel(B):
    """The [gamma_o; beta_o] convolution on CTA pairs (tcgen05.mma.cta_group::2, two images per pair, K-DYN in two
    passes against a zero patch) against the single-CTA kernel: whole network, even and odd batch (the odd image's
    partner CTA runs on zero-filled TMA boxes), inference and training (saved gamma / norm coefficients, gradients)."""
    import warnings
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200 import _lib as L
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda()
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(B, 64, 64, scale=8, seed=B, with_gt=True)]
    lib = L.load()
    out = {}
    try:
        for on in (0, 1):
            L.check(lib.dasr_set_sean_pair(on))
            net.eval()
            with torch.no_grad():
                sr = net(lq, depth, masks).clone()
            net.train()
            net.zero_grad(set_to_none=True)
            (net(lq, depth, masks) - gt).abs().mean().backward()
            torch.cuda.synchronize()
            out[on] = (sr, {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None})
    finally:
        L.check(lib.dasr_set_sean_pair(-1))
    d = (out[0][0] - out[1][0]).abs().max().item()
    print("B=%d: max|sr(pairs) - sr(single)| = %.3g (%s)" % (B, d, "bit-identical" if d == 0 else "differs"))
    assert d <= 1e-5
    for k, g0 in out[0][1].items():
        g1 = out[1][1][k]
        assert (g0 - g1).abs().max().item() <= 1e-3 * g0.abs().max().item() + 1e-9, k
