"""B200 parity tests in the fp32-split ("precise") storage mode -- dasr_set_planes(3), include/dasr.h.

The SAME kernels and the same host schedule as the product run here with every activation / packed weight stored as
three bf16 planes (hi + mid + lo = the fp32 value) and every tensor-core GEMM accumulating the six cross terms of its
operand planes: fp32-class arithmetic.  DepthNet's gradient is ill-conditioned (bf16 operand rounding alone moves the
REFERENCE's own early-layer gradients by 30-40 %, tests/test_gpu_backward.py), so a tolerance-class statement about
the backward algorithm -- K-DYN, SEAN / double-InstanceNorm backward, weight-norm, pooling, every index map -- needs
this mode.  Bars (north_star): SR max-abs <= 1e-4 against the reference's fp32 output; every parameter gradient of the
full-depth x8 / x4 / x2 / x3 networks within 1e-3 (relative L2) of the fp64 gradients of the reference
(tests/golden/*.npz: ``grad_sig`` / ``grad:*`` recorded from the real reference in .double(); the fp64 oracle, itself
pinned to those vectors by tests/test_oracle_golden.py, provides the full tensors).
"""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from common import (CASES, HR_CASES, INIT_CASES, case_tensors, cuda_net as _net, cuda_train_grads as _train_grads, load_golden,
                    oracle, oracle_grads64 as _oracle_grads64, skip_grad_param as _skip_param)

pytestmark = pytest.mark.gpu

# Relative L2 error per parameter gradient.  north_star asks for "the same relative tolerance" as the output (1e-4 in
# fp32): every tensor-valued parameter meets it (measured max 4.2e-5).  The 52 one-element SEAN blend scalars
# (alpha_gamma / alpha_beta) are differences of two full reductions, d(alpha) = <dW_s, W_s> - <dW_o, W_o> + bias terms,
# that cancel to ~1e-2 of their summands: their relative error is the summands' error times that factor (measured max
# 6.6e-4; the reference's own fp32 value deviates from its fp64 value by up to 9e-2 on them, golden ``grad_dev32``).
GRAD_TOL = 1e-4
SCALAR_TOL = 1e-3


@pytest.fixture()
def precise():
    from depth_aware_endoscopy_sr_b200 import _lib as L
    L.set_planes(3)
    try:
        yield L
    finally:
        L.set_planes(1)


def test_plane_split_is_exact_and_mode_switches_back(precise):
    L = precise
    assert L.planes() == 3
    x = torch.randn(2, 8, 8, 64, device="cuda") * 3.0
    a = L.act_from(x)
    assert torch.equal(L.act_value(a), x)
    b = L.act_from(torch.zeros_like(x))
    out = L.act_empty(2, 8, 8, 64, device="cuda")
    # x + 0 through the add kernel: plane sums in, plane split out -- must reproduce the fp32 values exactly
    L.check(L.load().dasr_add(L.ptr(a), None, L.ptr(b), L.ptr(out), a.numel(), L.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(L.act_value(out), x)


@pytest.mark.parametrize("name", INIT_CASES + CASES + HR_CASES)
def test_precise_forward_matches_reference_fp32(name, precise):
    """north_star's fp32 bar: max-abs <= 1e-4 on [0,1] pixels against the golden recorded from the real reference."""
    z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    net = _net(meta, sd).eval()
    st = meta["stride"]
    with torch.no_grad():
        sr = net(lq.cuda(), depth.cuda(), masks.cuda())
        pre = net.engine().infer(lq.cuda(), depth.cuda(), masks.cuda(), clamp=False)
    torch.cuda.synchronize()
    err = np.abs(sr.cpu().numpy()[:, :, ::st, ::st] - z["sr"]).max()
    ref_pre = z["pre_clamp"]
    err_pre = np.abs(pre.cpu().numpy()[:, :, ::st, ::st] - ref_pre).max()
    print("%s precise: max|sr - ref| %.3g   max|pre_clamp - ref| %.3g (range %.3g)" % (name, err, err_pre,
                                                                                       np.abs(ref_pre).max()))
    assert err <= 1e-4
    assert err_pre <= 1e-4 * max(1.0, np.abs(ref_pre).max())


@pytest.mark.parametrize("name", ["x8_b2_16", "x4_b1_24", "x2_b1_32", "x3_b1_24", "x8_b2_32_init", "x8_b1_24x40"] + HR_CASES)
def test_precise_gradients_match_reference_fp64(name, precise):
    """Every parameter gradient of the full-depth network against fp64: relative L2 error <= 1e-4 per tensor, <= 1e-3
    for the one-element blend scalars.

    DepthNet is piecewise smooth; two correct implementations whose forwards differ by eps sit on different pieces
    wherever a ReLU / clamp / L1-sign argument is within eps of its switching point, and their gradients then differ
    by ~sqrt(flipped fraction) (the reference's OWN fp32 gradients deviate from its fp64 gradients by 3e-4..2e-3 in
    the median and up to 20 % on single parameters, golden ``grad_dev32``).  So the tolerance-class comparison is made
    on the SAME piece: the fp64 oracle is evaluated with the activation pattern of the CUDA run
    (oracle.activation_pattern -- tests/test_oracle_golden.py shows that a forced evaluation is bit-identical to the
    free one on its own pattern, and the free oracle matches the reference's golden gradients to 1e-6).
    The free comparison (different pieces) is reported next to it and bounded loosely."""
    z, meta = load_golden(name)
    sd, inputs = case_tensors(meta)
    g, gw, loss, pattern = _train_grads(meta, sd, inputs)
    np.testing.assert_allclose(loss, z["loss"], rtol=5e-5, atol=1e-7)
    rec = {}
    gfree, _ = _oracle_grads64(meta, sd, inputs, record=rec)
    missing = sorted(set(rec) - set(pattern))
    assert not missing, "activation sites without a CUDA pattern: %s" % missing
    flips = sum(int((rec[k] != pattern[k]).sum()) for k in rec)
    units = sum(rec[k].numel() for k in rec)
    gref, gw_ref = _oracle_grads64(meta, sd, inputs, pattern=pattern)
    np.testing.assert_allclose(gw.numpy(), gw_ref.numpy(), rtol=1e-3, atol=1e-7)
    names = [str(n) for n in z["grad_names"]]
    sig = z["grad_sig"]
    dev32 = dict(zip(names, z["grad_dev32"])) if "grad_dev32" in z.files else {}
    rows = []
    for i, k in enumerate(names):
        if np.isnan(sig[i]).all():
            assert g[k] is None, "%s: the reference leaves this gradient None" % k
            continue
        assert g[k] is not None, "missing gradient for " + k
        if _skip_param(k, gref[k]):
            assert g[k].abs().max().item() <= 1e-6, k
            continue
        e_same = ((g[k] - gref[k]).norm() / gref[k].norm()).item()          # same smooth piece: the parity number
        e_free = ((g[k] - gfree[k]).norm() / gfree[k].norm()).item()        # different pieces (flips included)
        e_l2 = abs(g[k].norm().item() - sig[i][2]) / sig[i][2]              # L2 norm vs the REAL reference's golden
        rows.append((k, g[k].numel(), e_same, e_free, e_l2, float(dev32.get(k, np.nan))))
    e_same = np.array([r[2] for r in rows])
    e_free = np.array([r[3] for r in rows])
    worst = max(rows, key=lambda r: r[2])
    print("%s precise gradients, %d parameters; %d of %d activation units (%.1e) on a different piece than the fp64 "
          "oracle\n   same piece : rel-L2 median %.2e  max %.2e (%s)\n   free       : rel-L2 median %.2e  max %.2e\n"
          "   reference fp32 vs fp64 (golden): median %.2e  max %.2e" % (
              name, len(rows), flips, units, flips / units, np.median(e_same), e_same.max(), worst[0], np.median(e_free),
              e_free.max(), np.nanmedian([r[5] for r in rows]), np.nanmax([r[5] for r in rows])))
    out_dir = os.environ.get("DASR_PARITY_OUT")
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "grad_parity_precise_%s.json" % name), "w") as fh:
            json.dump(dict(case=name, flipped_units=flips, units=units,
                           params=[dict(param=k, numel=n, rel_l2_same_piece=a, rel_l2_free=b, l2_norm_vs_golden=c,
                                        reference_fp32_vs_fp64=d) for k, n, a, b, c, d in rows]), fh, indent=0)
    for k, n, a, b, c, d in rows:
        assert a <= (SCALAR_TOL if n == 1 else GRAD_TOL), (k, "same piece", a)
    assert np.median(e_same) <= 2e-5
    # different pieces: bounded by what flipping that many units can do (and sanity: the L2 norms of the real
    # reference's golden gradients are reproduced)
    assert np.median(e_free) <= max(5e-3, 30 * np.sqrt(flips / units)), (np.median(e_free), flips, units)
    assert np.median([r[4] for r in rows]) <= max(5e-3, 30 * np.sqrt(flips / units))


def test_precise_training_step_graph_and_eager_agree(precise):
    """The captured training step (side streams, CUDA graph) in precise mode equals the eager step."""
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=[0, 1], scale=8, nb=5).cuda().train()
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(2, 16, 16, scale=8, seed=4, with_gt=True)]
    eng = net.engine()

    def run(eager_overlap):
        eng.overlap_eager = eager_overlap
        net.zero_grad(set_to_none=True)
        sr = net(lq, depth, masks)
        (sr - gt).abs().mean().backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}

    g0, g1 = run(False), run(True)
    for k in g0:
        assert (g0[k] - g1[k]).abs().max().item() <= 1e-5 * g0[k].abs().max().item() + 1e-12, k


@pytest.mark.parametrize("name", ["x4_b1_24", "x2_b1_32", "x8_b2_16"])
def test_precise_train_steps_follow_the_oracle_trajectory(name, precise):
    """optimize_parameters for three steps (F_model_depthCond.py:158-192: forward, L1 + dynamic depth-mask loss,
    backward, Adam(1e-3, (0.9, 0.99)) over netG + the 10 loss weights): TrainStep on the CUDA path against the same
    steps taken by the fp64 ORACLE with torch.optim.Adam on the CPU -- nothing of this repository on the reference
    side.  Adam's first steps are ~lr * sign(g): a discontinuous function of the gradient, so an element whose tiny
    gradient differs in sign moves by 2 lr and the trajectories separate step by step (measured: 2-10 % of the
    elements after three steps).  Bars: the first loss (same parameters) to 1e-5, the second (one Adam step later) to
    1e-3, the third to 2e-2; >= 98 % of all parameter elements within 0.05 lr of the oracle's after the first step."""
    import depth_aware_endoscopy_sr_b200 as dasr
    _z, meta = load_golden(name)
    sd, (lq, depth, masks, gt) = case_tensors(meta)
    steps, lr = 3, 1e-3
    # ---- oracle (fp64, CPU)
    prm = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    wd = torch.ones(10, dtype=torch.float64, requires_grad=True)
    opt = torch.optim.Adam(list(prm.values()) + [wd], lr=lr, betas=(0.9, 0.99))
    ref_losses, ref_p1 = [], None
    for _ in range(steps):
        opt.zero_grad(set_to_none=True)
        sr = oracle.depthnet_forward(prm, lq.double(), depth.double(), masks.double(), scale=meta["scale"],
                                     which=meta["which"])
        total, *_ = oracle.training_loss(sr, gt.double(), masks.double(), wd)
        total.backward()
        opt.step()
        ref_losses.append(total.item())
        if ref_p1 is None:
            ref_p1 = {k: v.detach().clone() for k, v in prm.items() if v.grad is not None}
    # ---- CUDA
    net = _net(meta, sd).train()
    step = dasr.TrainStep(net, num_masks=10, lr=lr, betas=(0.9, 0.99))
    ins = [t.cuda() for t in (lq, depth, masks, gt)]
    losses, p1 = [], None
    for _ in range(steps):
        losses.append(step(*ins)[0].item())
        if p1 is None:
            p1 = {k: p.detach().double().cpu().clone() for k, p in net.named_parameters()}
    print("%s losses: cuda %s  oracle %s" % (name, losses, ref_losses))
    for got, ref, tol in zip(losses, ref_losses, (1e-5, 1e-3, 2e-2)):
        assert abs(got - ref) <= tol * abs(ref), (losses, ref_losses)
    # after ONE Adam step (identical starting point): every element moved by ~lr * sign(g); elements whose tiny
    # gradients differ in sign are 2 lr apart.  Later steps are compared through the losses only (a chaotic map).
    a = torch.cat([p1[k].reshape(-1) for k in ref_p1])
    b = torch.cat([ref_p1[k].reshape(-1) for k in ref_p1])
    agree = ((a - b).abs() <= 0.05 * lr).double().mean().item()
    print("   elements within 0.05 lr of the oracle's after the first Adam step: %.4f" % agree)
    assert agree >= 0.98, agree
    # the 10 loss weights move by ~lr per step; after three steps of a separating trajectory they agree to a fraction of it
    assert (step.dynamic_loss.trainable_weight.detach().double().cpu() - wd.detach()).abs().max().item() <= 0.5 * lr


@pytest.mark.parametrize("scale,B,h,w", [(8, 2, 24, 40), (4, 1, 40, 24)])
def test_precise_hr_blocks_on_non_square_frames_match_the_oracle(scale, B, h, w, precise):
    """Depth-guided blocks above LR resolution (which_ResBlk_depth = 0..15) on non-square frames with several images:
    the nearest-resized depth map / masks, the 32-channel SEAN instances and their style-table group against the fp32
    oracle (pinned to the reference for this configuration by the goldens x8_b1_16_hr / x4_b2_16_hr): <= 1e-4."""
    from depth_aware_endoscopy_sr_b200.synthetic import fill_state_dict, synthetic_inputs
    meta = dict(scale=scale, latent=64, which=tuple(range(16)))
    sd = fill_state_dict(oracle.state_layout(scale=scale, nb=16, which=meta["which"], latent=64, K=10), seed=11)
    lq, depth, masks = synthetic_inputs(B, h, w, scale=scale, seed=12)
    net = _net(meta, sd).eval()
    with torch.no_grad():
        ref = oracle.depthnet_forward(sd, lq, depth, masks, scale=scale, which=meta["which"])
        pre_ref = {}
        oracle.depthnet_forward(sd, lq, depth, masks, scale=scale, which=meta["which"], cap=pre_ref)
        sr = net(lq.cuda(), depth.cuda(), masks.cuda()).cpu()
        pre = net.engine().infer(lq.cuda(), depth.cuda(), masks.cuda(), clamp=False).cpu()
    err = (sr - ref).abs().max().item()
    err_pre = (pre - pre_ref["pre_clamp"]).abs().max().item()
    rng = pre_ref["pre_clamp"].abs().max().item()
    print("x%d B=%d %dx%d which=0..15 precise: max|sr - oracle| %.3g  max|pre_clamp - oracle| %.3g (range %.3g)" % (
        scale, B, h, w, err, err_pre, rng))
    assert err <= 1e-4
    assert err_pre <= 1e-4 * max(1.0, rng)
