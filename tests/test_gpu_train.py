"""B200 tests of the training-step kernels behind the generator (through the C ABI): K-LOSS forward / backward
against the CPU oracle of F_model_depthCond.py:163-190 + mask_loss.py:64-90, K-ADAM against torch.optim.Adam, and the
whole optimize_parameters step (TrainStep) against the same step assembled from torch ops."""
import warnings

import numpy as np
import pytest
import torch

from common import oracle
from depth_aware_endoscopy_sr_b200.synthetic import synthetic_inputs

pytestmark = pytest.mark.gpu


def _loss_inputs(B=2, h=16, w=24, scale=8, seed=5):
    lq, depth, masks, gt = synthetic_inputs(B, h, w, scale=scale, seed=seed, with_gt=True)
    g = torch.Generator().manual_seed(seed)
    # SR around HR with a few |d| > 1 outliers so both SmoothL1 branches and the L1 sign are exercised
    sr = gt + 0.3 * torch.randn(gt.shape, generator=g)
    sr.view(-1)[::97] += 2.5
    sr.view(-1)[::101] -= 1.7
    return sr, gt, masks


@pytest.mark.parametrize("shape", [(2, 16, 24, 8), (1, 24, 40, 4), (3, 32, 32, 2), (1, 135, 240, 8)])
def test_loss_forward_backward_match_oracle(shape):
    import depth_aware_endoscopy_sr_b200.loss as bl
    B, h, w, scale = shape
    sr, gt, masks = _loss_inputs(B, h, w, scale)
    wdyn = torch.linspace(-0.5, 0.7, 10)
    # oracle (fp64: the checker)
    sr_r = sr.double().requires_grad_(True)
    wd_r = wdyn.double().requires_grad_(True)
    total, l_pix, l_dyn, raw = oracle.training_loss(sr_r, gt.double(), masks.double(), wd_r)
    total.backward()
    # CUDA
    sr_c = sr.cuda().requires_grad_(True)
    wd_c = wdyn.cuda().requires_grad_(True)
    t, lp, ld, lk, sw = bl.training_loss(sr_c, gt.cuda(), masks.cuda(), wd_c)
    t.backward()
    torch.cuda.synchronize()
    np.testing.assert_allclose(t.item(), total.item(), rtol=2e-6)
    np.testing.assert_allclose(lp.item(), l_pix.item(), rtol=2e-6)
    np.testing.assert_allclose(ld.item(), l_dyn.item(), rtol=2e-6)
    np.testing.assert_allclose(lk.cpu().numpy(), [r.item() for r in raw], rtol=2e-6)
    np.testing.assert_allclose(sw.cpu().numpy(), torch.softmax(wdyn.double(), 0).numpy(), rtol=1e-6)
    gs, gr = sr_c.grad.cpu().double(), sr_r.grad
    assert (gs - gr).abs().max().item() <= 1e-6 * gr.abs().max().item() + 1e-12
    np.testing.assert_allclose(wd_c.grad.cpu().numpy(), wd_r.grad.numpy(), rtol=1e-4, atol=1e-7)


def test_loss_is_bit_reproducible_and_general_masks_take_the_exact_path():
    import depth_aware_endoscopy_sr_b200.loss as bl
    sr, gt, masks = _loss_inputs(2, 16, 16, 8)
    a = bl.loss_vector(sr.cuda(), gt.cuda(), masks.cuda(), torch.ones(10).cuda())
    b = bl.loss_vector(sr.cuda(), gt.cuda(), masks.cuda(), torch.ones(10).cuda())
    assert torch.equal(a[:3], b[:3])
    # soft, overlapping masks (not one-hot): the reference's literal m*SR - m*HR arithmetic
    g = torch.Generator().manual_seed(1)
    soft = torch.rand(masks.shape, generator=g) * (torch.rand(masks.shape, generator=g) > 0.5)
    sr_r = sr.double().requires_grad_(True)
    wd = torch.zeros(10, dtype=torch.float64, requires_grad=True)
    total, *_ = oracle.training_loss(sr_r, gt.double(), soft.double(), wd)
    total.backward()
    sr_c = sr.cuda().requires_grad_(True)
    t, *_ = bl.training_loss(sr_c, gt.cuda(), soft.cuda(), torch.zeros(10, device="cuda"))
    t.backward()
    np.testing.assert_allclose(t.item(), total.item(), rtol=5e-6)
    assert (sr_c.grad.cpu().double() - sr_r.grad).abs().max().item() <= 2e-6 * sr_r.grad.abs().max().item()


def test_reference_criteria_modules_are_drop_in():
    """cri_pix / dynamic_loss used separately, the way optimize_parameters calls them (F_model_depthCond.py:163-189)."""
    import depth_aware_endoscopy_sr_b200 as dasr
    sr, gt, masks = _loss_inputs(2, 16, 16, 8)
    cri_pix = dasr.L1Loss()
    dyn = dasr.dynamic_weight_mask_loss({"dynamic_criterion": "smoothl1", "dynamic_weight": 10.0}, num_trainable_para=10).cuda()
    assert [k for k, _ in dyn.named_parameters()] == ["trainable_weight"]
    sr_c = sr.cuda().requires_grad_(True)
    l_pix = 1.0 * cri_pix(sr_c, gt.cuda())
    raw, weighted, l_dyn, sw = dyn(sr_c, gt.cuda(), masks.cuda())
    (l_pix + l_dyn).backward()
    sr_r = sr.double().requires_grad_(True)
    wd = torch.ones(10, dtype=torch.float64, requires_grad=True)
    total, lp, ld, rawr = oracle.training_loss(sr_r, gt.double(), masks.double(), wd)
    total.backward()
    np.testing.assert_allclose((l_pix + l_dyn).item(), total.item(), rtol=2e-6)
    assert len(raw) == 10 and len(weighted) == 10 and sw.shape == (10,)
    np.testing.assert_allclose([r.item() for r in raw], [r.item() for r in rawr], rtol=2e-6)
    assert (sr_c.grad.cpu().double() - sr_r.grad).abs().max().item() <= 1e-6 * sr_r.grad.abs().max().item()
    np.testing.assert_allclose(dyn.trainable_weight.grad.cpu().numpy(), wd.grad.numpy(), rtol=1e-4, atol=1e-7)
    with pytest.raises(NotImplementedError):
        dasr.dynamic_weight_mask_loss({"dynamic_criterion": "l2", "dynamic_weight": 1.0})


def test_fused_adam_matches_torch_adam():
    import depth_aware_endoscopy_sr_b200 as dasr
    torch.manual_seed(0)
    shapes = [(64, 64, 3, 3), (1,), (10, 10, 1, 1), (7,), (3, 32, 9, 9), (33,)]
    p_ref = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    p_new = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=1e-3, betas=(0.9, 0.99), weight_decay=0)
    o_new = dasr.FusedAdam(p_new, lr=1e-3, betas=(0.9, 0.99), weight_decay=0)
    for step in range(6):
        for a, b in zip(p_ref, p_new):
            g = torch.randn_like(a) * (10.0 ** (step - 3))
            a.grad = g.clone()
            b.grad = None if (step == 2 and a.numel() == 7) else g.clone()
            if b.grad is None:
                a.grad = None
        if step == 3:
            for grp in list(o_ref.param_groups) + list(o_new.param_groups):
                grp["lr"] = 3e-4         # what the reference's lr schedulers do through param_groups
        v0 = p_new[0]._version
        o_ref.step()
        o_new.step()
        assert p_new[0]._version > v0
    for a, b in zip(p_ref, p_new):
        assert (a - b).abs().max().item() <= 2e-6 * a.abs().max().item() + 1e-7
    # ---- checkpoint interchange (codes/models/base_model.py resume_training -> optimizer.load_state_dict): the state
    # dict has torch.optim.Adam's layout, so either optimiser resumes from the other's *.state file
    sd = o_new.state_dict()
    sd_ref = o_ref.state_dict()
    assert set(sd["state"].keys()) == set(sd_ref["state"].keys()) == set(range(len(shapes)))
    for i in range(len(shapes)):
        assert float(sd["state"][i]["step"]) == float(sd_ref["state"][i]["step"]) == (5 if shapes[i] == (7,) else 6)
        assert (sd["state"][i]["exp_avg"] - sd_ref["state"][i]["exp_avg"]).abs().max().item() <= 2e-6 * sd_ref["state"][i]["exp_avg"].abs().max().item() + 1e-9
    p_a = [p.detach().clone().requires_grad_(True) for p in p_ref]      # torch Adam resuming from FusedAdam's file
    p_b = [p.detach().clone().requires_grad_(True) for p in p_ref]      # FusedAdam resuming from torch Adam's file
    o_a = torch.optim.Adam(p_a, lr=3e-4, betas=(0.9, 0.99))
    o_b = dasr.FusedAdam(p_b, lr=3e-4, betas=(0.9, 0.99))
    o_a.load_state_dict(sd)
    o_b.load_state_dict(sd_ref)
    for a, b, r in zip(p_a, p_b, p_ref):
        g = torch.randn_like(a)
        a.grad, b.grad, r.grad = g.clone(), g.clone(), g.clone()
    o_a.step(), o_b.step(), o_ref.step()
    for a, b, r in zip(p_a, p_b, p_ref):
        assert (a - r).abs().max().item() <= 2e-6 * r.abs().max().item() + 1e-7
        assert (b - r).abs().max().item() <= 2e-6 * r.abs().max().item() + 1e-7


def test_fused_adam_weight_decay_skips_parameters_without_gradient():
    """train.weight_decay_G > 0: inside the engine's flat gradient buffer a never-used parameter is a zero slice; like
    torch.optim.Adam (which skips grad-None parameters) it must neither decay nor build moments."""
    import depth_aware_endoscopy_sr_b200 as dasr
    from depth_aware_endoscopy_sr_b200 import _lib as L
    torch.manual_seed(1)
    shapes = [(8, 4), (16,), (5, 3), (12,)]
    p_ref = [torch.randn(s, device="cuda").requires_grad_(True) for s in shapes]
    p_new = [p.detach().clone().requires_grad_(True) for p in p_ref]
    o_ref = torch.optim.Adam(p_ref, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.1)
    o_new = dasr.FusedAdam(p_new, lr=1e-2, betas=(0.9, 0.99), weight_decay=0.1)
    o_new.step()                                            # re-homes the parameters into the flat buffer (no grads yet)
    for step in range(3):
        offs, n = {}, 0
        for p in p_new:
            offs[id(p)] = n
            n += L.flat_pad(p.numel())
        flat = torch.zeros(n, device="cuda")
        for i, (a, b) in enumerate(zip(p_ref, p_new)):
            if i == 1:                                      # "never used": zero slice, grad None
                a.grad, b.grad = None, None
                continue
            g = torch.randn_like(a)
            a.grad = g.clone()
            flat[offs[id(b)]:offs[id(b)] + b.numel()] = g.reshape(-1)
            b.grad = flat[offs[id(b)]:offs[id(b)] + b.numel()].view(b.shape)
        L.register_flat_grad(flat, offs)
        o_ref.step()
        o_new.step()
    for a, b in zip(p_ref, p_new):
        assert (a - b).abs().max().item() <= 2e-6 * a.abs().max().item() + 1e-7
    assert torch.equal(p_new[1].detach(), p_ref[1].detach())                 # untouched


def _torch_step_reference(sd, meta, inputs, steps, lr):
    """The same optimize_parameters step assembled from OUR generator + torch criteria + torch.optim.Adam: isolates the
    K-LOSS / K-ADAM kernels and the step plumbing (the generator is shared, so this says nothing about ITS gradients --
    those are pinned to the fp64 oracle / the reference goldens in test_gpu_precise.py and test_gpu_backward.py,
    including three optimizer steps against the oracle: test_precise_train_steps_follow_the_oracle_trajectory)."""
    import depth_aware_endoscopy_sr_b200 as dasr
    lq, depth, masks, gt = [t.cuda() for t in inputs]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], nb=meta["nb"],
                            depth_latent_ch=meta.get("latent", 256)).cuda().train()
    net.load_state_dict(sd)
    wd = torch.ones(10, device="cuda", requires_grad=True)
    opt = torch.optim.Adam([p for p in net.parameters()] + [wd], lr=lr, betas=(0.9, 0.99))
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        sr = net(lq, depth, masks)
        total, *_ = oracle.training_loss(sr, gt, masks, wd)
        total.backward()
        opt.step()
        losses.append(total.item())
    return losses, net, wd


def test_train_step_matches_torch_assembled_step():
    import depth_aware_endoscopy_sr_b200 as dasr
    meta = dict(scale=8, which=(0, 1), nb=5)
    torch.manual_seed(11)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=[0, 1], scale=8, nb=5).cuda().train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    inputs = synthetic_inputs(2, 16, 16, scale=8, seed=4, with_gt=True)
    ref_losses, ref_net, ref_wd = _torch_step_reference(sd, meta, inputs, steps=4, lr=1e-3)
    step = dasr.TrainStep(net, num_masks=10, lr=1e-3, betas=(0.9, 0.99))
    lq, depth, masks, gt = [t.cuda() for t in inputs]
    losses = [step(lq, depth, masks, gt)[0].item() for _ in range(4)]
    # same kernels for the generator on both sides; the criteria / optimiser differ (CUDA kernels vs torch ops)
    np.testing.assert_allclose(losses, ref_losses, rtol=5e-3)
    assert losses[-1] < losses[0]
    a = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    b = torch.cat([p.detach().reshape(-1) for p in ref_net.parameters()])
    # Adam normalises every step to ~lr, so bf16-level differences in tiny gradients can flip individual updates;
    # the bulk of the 4-step trajectory must coincide
    close = ((a - b).abs() <= 3e-4).float().mean().item()
    assert close >= 0.90, close
    assert (step.dynamic_loss.trainable_weight.detach() - ref_wd.detach()).abs().max().item() <= 1e-3


def test_train_step_cuda_graph_replays_the_eager_trajectory():
    """graph=True records the step once and replays it; the trajectory must equal the kernel-by-kernel one
    (same kernels, same order; Adam's step-dependent scalars come from device memory)."""
    import depth_aware_endoscopy_sr_b200 as dasr
    inputs = [[t.cuda() for t in synthetic_inputs(2, 16, 16, scale=8, seed=s, with_gt=True)] for s in (4, 5)]
    traj = {}
    for graph in (False, True):
        torch.manual_seed(11)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net = dasr.DepthNet(which_ResBlk_depth=[0, 1], scale=8, nb=5).cuda().train()
        step = dasr.TrainStep(net, num_masks=10, lr=1e-3, betas=(0.9, 0.99), graph=graph)
        losses = [step(*inputs[i % 2])[0].item() for i in range(6)]
        traj[graph] = (losses, torch.cat([p.detach().reshape(-1) for p in net.parameters()]).clone())
    # the first steps are identical; later ones drift in the 5th digit (atomic accumulation order differs run to run)
    np.testing.assert_allclose(traj[True][0][:3], traj[False][0][:3], rtol=2e-4)
    np.testing.assert_allclose(traj[True][0], traj[False][0], rtol=2e-2)
    # fp32 atomics in the weight-gradient kernels make runs differ in the last bits; Adam amplifies sign flips of
    # ~zero gradients, so compare the bulk
    close = ((traj[True][1] - traj[False][1]).abs() <= 1e-4).float().mean().item()
    assert close >= 0.5, close


@pytest.mark.parametrize("scale,which,latent,unused", [(4, list(range(14)), 256, ("upscale1.",)),
                                                      (2, list(range(16)), 32, ("upscale1.", "upscale2.")),
                                                      (3, list(range(16)), 256, ("upscale1.", "upscale2."))])
def test_training_step_at_x4_and_x2(scale, which, latent, unused):
    """BASELINE configs[3]: the x4 / x2 (/ x3: PixelShuffle(3) tail) variants train through the same kernels (smaller
    upsampler, 64-channel blocks 15/16, 32-channel latent at x2); parameters the reference never touches get no
    gradient."""
    import depth_aware_endoscopy_sr_b200 as dasr
    torch.manual_seed(5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=which, scale=scale, nb=16, depth_latent_ch=latent).cuda().train()
    lq, depth, masks, gt = [t.cuda() for t in synthetic_inputs(2, 32, 32, scale=scale, seed=9, with_gt=True)]
    sr = net(lq, depth, masks)
    assert tuple(sr.shape) == (2, 3, 32 * scale, 32 * scale)
    (sr - gt).abs().mean().backward()
    for name, p in net.named_parameters():
        if name.startswith(unused) or name.startswith("depth-residual14."):
            assert p.grad is None, name
        else:
            assert p.grad is not None and torch.isfinite(p.grad).all(), name
    net.zero_grad(set_to_none=True)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    inputs = [t.cpu() for t in (lq, depth, masks, gt)]
    ref_losses, _, _ = _torch_step_reference(sd, dict(scale=scale, which=which, nb=16, latent=latent), inputs, steps=3, lr=1e-3)
    step = dasr.TrainStep(net, num_masks=10, lr=1e-3, betas=(0.9, 0.99))
    losses = [step(lq, depth, masks, gt)[0].item() for _ in range(3)]
    assert all(np.isfinite(losses))
    np.testing.assert_allclose(losses, ref_losses, rtol=1e-2)     # same trajectory as torch criteria + torch Adam
