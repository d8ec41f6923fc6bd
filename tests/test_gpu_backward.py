"""B200 tests of the backward pass (through the C ABI).

 1. every backward kernel against torch autograd of the op it differentiates, on identical (bf16-rounded) inputs;
 2. whole-network gradients at full depth in the product arithmetic (bf16), PER PARAMETER, against the fp64 oracle
    on the same smooth piece of the network (oracle.activation_pattern: DepthNet is piecewise smooth, and comparing
    across ReLU / clamp / L1-sign flips measures the flips, not the implementation);
 3. the fp32-class statement (every gradient within 1e-4 of fp64, x8 / x4 / x2 / x3) lives in test_gpu_precise.py.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from common import ROOT, case_tensors, load_golden, oracle

pytestmark = pytest.mark.gpu


def _tool(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def kb():
    return _tool("check_bwd")


@pytest.mark.parametrize("shape", [(2, 16, 16, 64, 64, 3, 3), (2, 64, 64, 128, 128, 3, 3), (1, 24, 40, 64, 32, 3, 3),
                                   (1, 24, 40, 32, 64, 3, 3), (1, 31, 31, 128, 256, 3, 3), (1, 135, 240, 64, 64, 3, 3),
                                   (1, 64, 64, 32, 32, 9, 1), (2, 64, 64, 32, 128, 3, 3), (1, 24, 40, 32, 32, 3, 3),
                                   (1, 135, 240, 32, 32, 3, 3), (1, 40, 300, 32, 32, 9, 1), (3, 64, 64, 64, 128, 3, 3), (1, 24, 24, 64, 288, 3, 3)])
def test_conv_wgrad_kernel(kb, shape):
    B, H, W, Cin, Cout, kh, kw = shape
    kb.failures.clear()
    kb.wgrad_case(B, H, W, Cin, Cout, kh, kw)
    assert not kb.failures, kb.failures


def test_sean_instance_norm_backward_kernels(kb):
    kb.failures.clear()
    kb.sean_bwd_case()
    kb.sean_bwd_case(1, 24, 40, 32)
    assert not kb.failures, kb.failures


def test_dynamic_conv_and_style_backward_kernels(kb):
    kb.failures.clear()
    kb.dyn_bwd_case()
    kb.dyn_bwd_case(B=3, H=24, W=40)
    kb.dyn_bwd_case(B=4, H=64, W=64)
    assert not kb.failures, kb.failures


def test_small_backward_kernels(kb):
    kb.failures.clear()
    kb.misc_bwd_case()
    assert not kb.failures, kb.failures


# ------------------------------------------------------------------------------------------------ whole network
# (the tolerance-class statement about the backward ALGORITHM is tests/test_gpu_precise.py: every parameter gradient
# of the full-depth x8 / x4 / x2 / x3 networks within 1e-4 of fp64 in the fp32-split mode of the same kernels.)
# per-parameter bounds of the bf16 product arithmetic, relative to what bf16 rounding does to the REFERENCE itself on
# the same smooth piece (oracle.bf16_operands: conv operands rounded to bf16, everything else fp64):
#   tensors:  err <= 3.5 * err_emulated + 0.03      (measured: ratio median 1.0, max 2.9 over 5 cases x ~440 tensors)
#   one-element blend scalars (differences of two cancelling full reductions: their relative error is unbounded as the
#   sum approaches zero): ABSOLUTE error <= 0.75 * rms of the reference's blend-scalar gradients of the case
BF16_TENSOR_FACTOR, BF16_TENSOR_FLOOR, BF16_SCALAR_ABS = 3.5, 0.03, 0.75


@pytest.mark.parametrize("name", ["x8_b2_32_init", "x8_b2_16", "x4_b1_24", "x2_b1_32", "x3_b1_24", "x8_b1_16_hr", "x4_b2_16_hr"])
def test_full_depth_bf16_gradients_per_parameter(name):
    """Product arithmetic (bf16 operands and activations, fp32 accumulate): every parameter gradient of the full-depth
    network against the fp64 oracle evaluated ON THE SAME SMOOTH PIECE (the activation pattern of the CUDA run,
    oracle.activation_pattern) -- what is left is the effect of bf16 rounding itself, bounded PER PARAMETER:
    see the bounds above.  The free comparison (different pieces: bf16 flips ~1e-3 of the ReLU / clamp units, and each flip
    moves the gradient discontinuously) is reported for reference; so is the same measurement for the oracle with
    bf16-rounded conv operands (what torch.autocast does to the reference)."""
    import json
    from common import cuda_train_grads, oracle_grads64, skip_grad_param
    z, meta = load_golden(name)
    sd, inputs = case_tensors(meta)
    g, _gw, loss, pattern = cuda_train_grads(meta, sd, inputs)
    np.testing.assert_allclose(loss[:3], z["loss"][:3], rtol=2e-2)
    rec = {}
    gfree, _ = oracle_grads64(meta, sd, inputs, record=rec)
    assert not (set(rec) - set(pattern))
    flips = sum(int((rec[k] != pattern[k]).sum()) for k in rec)
    units = sum(rec[k].numel() for k in rec)
    gref, _ = oracle_grads64(meta, sd, inputs, pattern=pattern)
    with oracle.bf16_operands():
        gemu, _ = oracle_grads64(meta, sd, inputs, pattern=pattern)
    rows = []
    for k, r in gref.items():
        if r is None:
            assert g[k] is None, "%s: the reference leaves this gradient None" % k
            continue
        assert g[k] is not None, "missing gradient for " + k
        if skip_grad_param(k, r):
            continue
        rows.append((k, r.numel(), ((g[k] - r).norm() / r.norm()).item(), ((g[k] - gfree[k]).norm() / gfree[k].norm()).item(),
                     ((gemu[k] - r).norm() / r.norm()).item()))
    sc_keys = [k for k, n, *_ in rows if n == 1]
    sc_rms = float(np.sqrt(np.mean([gref[k].item() ** 2 for k in sc_keys])))
    sc_abs = max(abs(g[k].item() - gref[k].item()) for k in sc_keys) / sc_rms
    ratio = max(a / (BF16_TENSOR_FACTOR * c + BF16_TENSOR_FLOOR) for k, n, a, b, c in rows if n > 1)
    ten = np.array([r[2] for r in rows if r[1] > 1])
    sca = np.array([r[2] for r in rows if r[1] == 1])
    emu = np.array([r[4] for r in rows if r[1] > 1])
    free = np.array([r[3] for r in rows if r[1] > 1])
    wt = max((r for r in rows if r[1] > 1), key=lambda r: r[2])
    print("%s bf16 gradients: %d of %d units (%.1e) on a different piece\n   same piece, tensors: median %.3g max %.3g (%s);"
          " scalars: median %.3g max %.3g\n   bf16-operand ORACLE on the same piece, tensors: median %.3g max %.3g\n"
          "   free comparison, tensors: median %.3g max %.3g" % (
              name, flips, units, flips / units, np.median(ten), ten.max(), wt[0], np.median(sca), sca.max(),
              np.median(emu), emu.max(), np.median(free), free.max()))
    print("   worst tensor err / bound %.2f ; worst blend-scalar abs err / rms %.3f (bound %.2f)" % (ratio, sc_abs, BF16_SCALAR_ABS))
    out_dir = os.environ.get("DASR_PARITY_OUT")
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "grad_parity_bf16_%s.json" % name), "w") as fh:
            json.dump(dict(case=name, flipped_units=flips, units=units,
                           params=[dict(param=k, numel=n, rel_l2_same_piece=a, rel_l2_free=b, bf16_operand_oracle_same_piece=c)
                                   for k, n, a, b, c in rows]), fh, indent=0)
    for k, n, a, b, c in rows:
        if n > 1:
            assert a <= BF16_TENSOR_FACTOR * c + BF16_TENSOR_FLOOR, (k, a, c)
        else:
            assert abs(g[k].item() - gref[k].item()) <= BF16_SCALAR_ABS * sc_rms, (k, g[k].item(), gref[k].item(), sc_rms)
    assert g["depth-residual14.conv1.0.weight"] is None       # never used by the reference either


def test_side_stream_backward_equals_the_one_stream_backward():
    """The weight gradients and the other leaf chains of the backward run on side streams inside a captured training
    step (Engine.wgrad_overlap / leaf_overlap); forced here kernel by kernel, with every combination of stream count
    and split-K divisor, the gradients must equal the one-stream backward up to the order of the fp32 atomics."""
    import warnings
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=[0, 1], scale=8, nb=5).cuda().train()
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(2, 32, 32, scale=8, seed=9, with_gt=True)]
    eng = net.engine()

    def run(**kw):
        for k, v in kw.items():
            setattr(eng, k, v)
        eng._wg_streams = {}
        net.zero_grad(set_to_none=True)
        sr = net(lq, depth, masks)
        (sr - gt).abs().mean().backward()
        torch.cuda.synchronize()
        return sr.detach().clone(), {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}

    sr0, g0 = run(overlap_eager=False)
    for streams, div, leaf in ((1, 1, False), (3, 2, True), (6, 2, True), (2, 3, True)):
        sr1, g1 = run(overlap_eager=True, wgrad_overlap=True, wgrad_streams=streams, wgrad_ksplit_div=div, leaf_overlap=leaf)
        assert torch.equal(sr0, sr1)
        assert g0.keys() == g1.keys()
        for k in g0:
            err = (g0[k] - g1[k]).abs().max().item()
            assert err <= 1e-4 * g0[k].abs().max().item() + 1e-9, (streams, div, leaf, k, err)


def test_backward_twice_raises_and_weights_repack_after_update():
    from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs
    import depth_aware_endoscopy_sr_b200 as dasr
    import warnings
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(range(14)), scale=8, nb=16).cuda().train()
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(2, 16, 16, scale=8, seed=1, with_gt=True)]
    opt = torch.optim.Adam(net.parameters(), lr=1e-3, betas=(0.9, 0.99))
    losses = []
    for _ in range(4):
        opt.zero_grad()
        sr = net(lq, depth, masks)
        loss = (sr - gt).abs().mean()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0], losses          # the optimiser sees real gradients and the engine repacks
    sr = net(lq, depth, masks)
    loss = (sr - gt).abs().mean()
    loss.backward()
    with pytest.raises(RuntimeError):
        loss.backward()
