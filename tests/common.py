"""Shared helpers of the test-suite (case table, golden loading, oracle driver)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from depth_aware_endoscopy_sr_b200.synthetic import fill_state_dict, synthetic_inputs  # noqa: E402
from oracle import depthnet_oracle as oracle  # noqa: E402  (the checker; tests may import it)

GOLDEN = os.path.join(ROOT, "tests", "golden")
INIT_CASES = ["x8_b1_64_init", "x8_b2_32_init"]   # the reference's own random init (strict tolerance)
CASES = ["x8_b2_16", "x8_b1_64", "x8_b1_24x40", "x4_b1_24", "x2_b1_32", "x3_b1_24"]
# depth-guided blocks ABOVE LR resolution (which_ResBlk_depth = 0..15: the 32-channel blocks behind upscale1 / upscale2
# are depth-guided too and resize depth map / masks, normalization.py:58-59)
HR_CASES = ["x8_b1_16_hr", "x4_b2_16_hr"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    scale, latent, B, h, w, seed, stride = [int(v) for v in z["meta"]]
    which = tuple(int(v) for v in z["which"])
    init = str(z["init"]) if "init" in z.files else "synthetic"
    return z, dict(scale=scale, latent=latent, B=B, h=h, w=w, seed=seed, stride=stride, which=which, init=init)


def case_tensors(meta, with_gt=True):
    if meta.get("init", "synthetic") == "default":
        # the reference's own init: our DepthNet consumes the RNG exactly like the reference's constructor
        # (checked against the reference in make_golden.py via sd_checksum)
        import warnings
        from depth_aware_endoscopy_sr_b200.arch import DepthNet
        torch.manual_seed(meta["seed"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            net = DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"],
                           nb=16, nf=64, depthRangeNum=10)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    else:
        layout = oracle.state_layout(scale=meta["scale"], nb=16, which=meta["which"], latent=meta["latent"], K=10)
        sd = fill_state_dict(layout, seed=meta["seed"])
    inp = synthetic_inputs(meta["B"], meta["h"], meta["w"], scale=meta["scale"], seed=meta["seed"], with_gt=with_gt)
    return sd, inp


def psnr(a, b):
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 99.0 if mse == 0 else 10.0 * np.log10(1.0 / mse)


# ---- inputs of the golden vectors for the steps either side of the generator (tests/golden/make_io_golden.py)
IO_DEPTH_SHAPES = [(3, 16, 24), (2, 64, 64), (1, 135, 240)]


def io_depth_input(B, h, w):
    """(depth in [0.01, 10) with several pixels at the exclusive upper edge, depth in [0, 1) for depthFixedRange)."""
    g = torch.Generator().manual_seed(B * 1000 + h)
    depth = 0.01 + 9.99 * torch.rand(B, 1, h, w, generator=g)
    depth[0, 0, 0, :4] = depth[0].max()
    d01 = torch.rand(B, 1, h, w, generator=g)
    return depth, d01


def io_frames_input():
    """(sr with values outside [0,1] and exact .5/255 ties, gt) for tensor2img / PSNR / SSIM."""
    g = torch.Generator().manual_seed(3)
    gt = torch.rand(2, 3, 72, 100, generator=g)
    sr = gt + 0.05 * torch.randn(gt.shape, generator=g)
    sr.view(-1)[:512] = (torch.arange(512) // 2).float() / 255.0 + (torch.arange(512) % 2) * 0.5 / 255.0
    sr.view(-1)[600:700] = 1.3
    sr.view(-1)[700:800] = -0.2
    return sr, gt


# ---- whole-network gradient helpers of the GPU tests (test_gpu_precise.py, test_gpu_backward.py)
def cuda_net(meta, sd, nb=16):
    import warnings
    import depth_aware_endoscopy_sr_b200 as dasr
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = dasr.DepthNet(which_ResBlk_depth=list(meta["which"]), scale=meta["scale"], depth_latent_ch=meta["latent"],
                            nb=nb, nf=64, depthRangeNum=10)
    net.load_state_dict(sd, strict=True)
    return net.cuda()


def cuda_train_grads(meta, sd, inputs, nb=16):
    """One training forward + backward of the CUDA path (generator + the K-LOSS kernels).  Returns (parameter
    gradients, d/d(loss weights), the loss vector, the activation pattern of the run: every ReLU / LeakyReLU mask, the
    clamp mask and the L1 sign -- oracle.activation_pattern keys)."""
    import depth_aware_endoscopy_sr_b200.loss as bl
    lq, depth, masks, gt = inputs
    net = cuda_net(meta, sd, nb).train()
    net.engine().debug = {}
    wd = torch.ones(10, device="cuda", requires_grad=True)
    sr = net(lq.cuda(), depth.cuda(), masks.cuda())
    total, l_pix, l_dyn, lk, _sw = bl.training_loss(sr, gt.cuda(), masks.cuda(), wd)
    total.backward()
    torch.cuda.synchronize()
    g = {k: (p.grad.detach().double().cpu() if p.grad is not None else None) for k, p in net.named_parameters()}
    loss = np.array([total.item(), l_pix.item(), l_dyn.item()] + [v.item() for v in lk])
    pattern = {k[5:]: v.cpu() for k, v in net.engine().debug.items() if k.startswith("mask:")}
    srd = sr.detach()
    pattern["clamp"] = ((srd > 0) & (srd < 1)).cpu()
    pattern["l1.sign"] = (srd > gt.cuda()).cpu()
    net.engine().debug = None
    return g, wd.grad.detach().double().cpu(), loss, pattern


def oracle_grads64(meta, sd, inputs, pattern=None, record=None, nb=16):
    """fp64 gradients of the oracle; ``pattern``: evaluate on the smooth piece given by that activation pattern."""
    lq, depth, masks, gt = inputs
    sdr = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    wdyn = torch.ones(10, dtype=torch.float64, requires_grad=True)
    with oracle.activation_pattern(force=pattern, record=record):
        sr = oracle.depthnet_forward(sdr, lq.double(), depth.double(), masks.double(), scale=meta["scale"], nb=nb,
                                     which=meta["which"])
        total, *_ = oracle.training_loss(sr, gt.double(), masks.double(), wdyn)
    total.backward()
    return {k: v.grad for k, v in sdr.items()}, wdyn.grad


def skip_grad_param(k, ref):
    # a conv bias in front of an InstanceNorm has an exactly-zero gradient (the reference holds fp round-off there)
    return ref is None or ref.norm() < 1e-12 or ".conv1.0.bias" in k or ".conv2.0.bias" in k
