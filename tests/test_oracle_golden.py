"""The CPU oracle (oracle/depthnet_oracle.py) against the golden vectors recorded from the REAL reference
(tests/golden/make_golden.py).  fp32 round-off only: the restatement reorders a few sums."""
import numpy as np
import pytest
import torch

from common import CASES, HR_CASES, INIT_CASES, case_tensors, load_golden, oracle


@pytest.mark.parametrize("name", INIT_CASES + CASES + HR_CASES)
def test_oracle_forward_matches_reference(name):
    z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    chk = np.array([sum(v.double().sum().item() for v in sd.values()),
                    sum(v.double().abs().sum().item() for v in sd.values())])
    np.testing.assert_allclose(chk, z["sd_checksum"], rtol=1e-12)      # identical weights as the reference had
    cap = {}
    with torch.no_grad():
        sr = oracle.depthnet_forward(sd, lq, depth, masks, scale=meta["scale"], which=meta["which"], cap=cap)
    st = meta["stride"]
    np.testing.assert_allclose(cap["depthVec"].numpy(), z["depthVec"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(cap["fea_bef"].numpy()[:, ::4], z["fea_bef"], atol=2e-5, rtol=1e-5)
    np.testing.assert_allclose(cap["depth-residual1.out"].numpy()[:, ::4], z["dgb1_out"], atol=2e-4, rtol=1e-4)
    np.testing.assert_allclose(cap["depth-residual13.out"].numpy()[:, ::4], z["dgb13_out"], atol=2e-3, rtol=1e-4)
    for i in (15, 16):      # depth-guided blocks behind upscale1 / upscale2 (HR_CASES)
        if "dgb%d_out" % i in z.files:
            np.testing.assert_allclose(cap["depth-residual%d.out" % i].numpy()[:, ::4], z["dgb%d_out" % i], atol=2e-3,
                                       rtol=1e-4)
    np.testing.assert_allclose(cap["pre_clamp"].numpy()[:, :, ::st, ::st], z["pre_clamp"], atol=1e-4, rtol=1e-4)
    # the tolerance north_star states for fp32: max-abs <= 1e-4 on [0,1] pixels
    assert np.abs(sr.numpy()[:, :, ::st, ::st] - z["sr"]).max() <= 1e-4


@pytest.mark.parametrize("name", ["x8_b2_16", "x4_b1_24", "x2_b1_32", "x3_b1_24"] + HR_CASES)
def test_oracle_loss_and_gradients_match_reference(name):
    z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    # fp64 on both sides (the golden gradients were recorded from the reference modules in .double())
    sd = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    wdyn = torch.ones(10, dtype=torch.float64, requires_grad=True)
    sr = oracle.depthnet_forward(sd, lq.double(), depth.double(), masks.double(), scale=meta["scale"],
                                 which=meta["which"])
    total, l_pix, l_dyn, raw = oracle.training_loss(sr, gt.double(), masks.double(), wdyn)
    got = np.array([total.item(), l_pix.item(), l_dyn.item()] + [r.item() for r in raw])
    np.testing.assert_allclose(got, z["loss"], rtol=1e-9, atol=1e-12)
    total.backward()
    np.testing.assert_allclose(wdyn.grad.numpy(), z["dyn_weight_grad"], rtol=1e-7, atol=1e-12)
    names = [str(n) for n in z["grad_names"]]
    sig = z["grad_sig"]
    unused = 0
    for i, k in enumerate(names):
        g = sd[k].grad
        if np.isnan(sig[i]).all():          # the reference left .grad = None (parameter never used)
            assert g is None or float(g.abs().max()) == 0.0, k
            unused += 1
            continue
        gen = torch.Generator().manual_seed(sum(map(ord, k)))
        proj = torch.randn(g.numel(), generator=gen, dtype=torch.float64)
        g64 = g.flatten()
        got_sig = np.array([g64.sum().item(), g64.abs().sum().item(), g64.norm().item(), (g64 * proj).sum().item()])
        np.testing.assert_allclose(got_sig, sig[i], rtol=1e-6, atol=1e-9 * sig[i][1] + 1e-12, err_msg=k)
    assert unused > 0   # depth-residual14.* (SURVEY.md headline fact 5)
    for key in z.files:
        if key.startswith("grad:"):
            ref = z[key]
            np.testing.assert_allclose(sd[key[5:]].grad.numpy(), ref, atol=1e-9 * np.abs(ref).max() + 1e-12, rtol=1e-6)
            # (conv biases in front of an InstanceNorm have an exactly-zero gradient: only round-off is left)


def test_dynamic_conv_restatement_is_exact():
    """Style branch == per-image dynamic 3x3 convolution of the mask, in fp64."""
    torch.manual_seed(0)
    B, K, L, nf, H, W = 2, 10, 32, 8, 9, 11
    stp = torch.randn(B, K, L, dtype=torch.float64)
    w = torch.randn(nf, L, 3, 3, dtype=torch.float64)
    b = torch.randn(nf, dtype=torch.float64)
    mask = (torch.rand(B, K, H, W) > 0.7).double() * torch.rand(B, K, H, W).double()
    style_map = torch.einsum("bkc,bkhw->bchw", stp, mask)
    ref = torch.nn.functional.conv2d(style_map, w, b, padding=1)
    got = oracle.dynconv_apply(oracle.style_table(w, stp), b, mask)
    assert (ref - got).abs().max().item() < 1e-11


@pytest.mark.parametrize("name", ["x8_b2_16", "x4_b1_24"])
def test_forced_activation_pattern_is_the_same_function(name):
    """oracle.activation_pattern: recording the pattern of a run and forcing it back reproduces the run exactly (so
    a forced evaluation IS the reference function on that smooth piece), every nonlinearity has a key, and forcing
    a pattern with a few flipped units changes the gradient discontinuously (what the GPU gradient tests avoid)."""
    _z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)

    def run(force=None, record=None):
        sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
        wd = torch.ones(10, dtype=torch.float64, requires_grad=True)
        with oracle.activation_pattern(force=force, record=record):
            sr = oracle.depthnet_forward(sdr, lq.double(), depth.double(), masks.double(), scale=meta["scale"],
                                         which=meta["which"])
            total, *_ = oracle.training_loss(sr, gt.double(), masks.double(), wd)
        total.backward()
        return total.item(), {k: v.grad for k, v in sdr.items() if v.grad is not None}

    rec = {}
    t0, g0 = run(record=rec)
    n_dgb = len([i for i in meta["which"] if i < 13 or meta["scale"] <= 3 or i >= 13])
    assert "clamp" in rec and "l1.sign" in rec and "head.2" in rec and "depth-residual1.norm1.actv" in rec
    assert sum(k.endswith(".actv") for k in rec) == 2 * sum(1 for k in rec if k.endswith(".norm2.out")) > 0 and n_dgb
    t1, g1 = run(force=rec)
    assert t1 == t0
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k
    flipped = {k: v.clone() for k, v in rec.items()}
    m = flipped["depth-residual13.norm2.out"] if "depth-residual13.norm2.out" in flipped else flipped["head.2"]
    m.view(-1)[::97] ^= True
    _t2, g2 = run(force=flipped)
    k = "head.0.weight_v"
    assert ((g2[k] - g0[k]).norm() / g0[k].norm()).item() > 1e-3
